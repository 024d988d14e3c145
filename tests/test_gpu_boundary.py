"""GPU: the boundary's other entry points against the oracle — 16-byte packed records (incl. the duration
overflow side array), concurrent producers (the reference calls this seam from 4*NumCPU goroutines,
aggregator/data.go:230-232), incremental endpoint-table patches under churn, NUMA-local pinned memory."""
import threading

import numpy as np
import pytest

import oracle_lib as ol
from alaz_b200 import abi, capi
from helpers import edges_equal, explain_diff

pytestmark = pytest.mark.gpu


def test_packed_records_bit_exact_including_duration_overflow():
    S, N = 800, 1_500_000
    t = ol.Topo(S, seed=31, mix=abi.MIX_ALL)
    ev = t.events(0, N)
    # durations beyond 2^32 ns exercise the overflow array; one beyond 2^40 lands in the top bucket
    big = np.arange(0, N, 997)
    ev["duration_ns"][big] = (np.uint64(1) << np.uint64(32)) + ev["duration_ns"][big] * np.uint64(7)
    ev["duration_ns"][5] = np.uint64(0xFFFF_FFFF_FFFF_FFFF)
    rec16, ovf = capi.pack_l7(ev)
    assert len(ovf) == len(big) + (0 if 5 in big else 1)
    o = ol.Oracle(); o.load_tables(t.pod_ip, t.svc_ip); o.process(ev, 4)
    h = capi.Handle(max_endpoints=4 * S, max_pairs=1 << 16, max_batch=200_000)   # several chunks per submit
    h.load_tables(t.pod_ip, t.svc_ip)
    h.submit_packed(rec16, ovf)
    got, exp = h.flush(), o.edges()
    assert edges_equal(got, exp), explain_diff(got, exp)
    st, ost = h.stats(), o.stats()
    for k in ("events_in", "rows_emitted", "not_request", "src_unresolved"):
        assert st[k] == ost[k], (k, st[k], ost[k])
    # second window through a NUMA-local pinned buffer (no staging memcpy), no overflow entries
    ev2 = t.events(N, 300_001)
    r2, o2 = capi.pack_l7(ev2)
    assert len(o2) == 0
    pin = capi.PinnedBuffer(len(r2), abi.L7_REC16, handle=h)
    pin.array[:] = r2
    h.submit_packed_ptr(pin.ptr, len(r2))
    o.reset_window(); o.process(ev2, 4)
    got, exp = h.flush(), o.edges()
    assert edges_equal(got, exp), explain_diff(got, exp)
    pin.free()
    h.close()


@pytest.mark.parametrize("packed", [False, True])
def test_eight_concurrent_producers(packed):
    S, N, T = 600, 2_400_000, 8
    t = ol.Topo(S, seed=77, mix=abi.MIX_ALL)
    ev = t.events(0, N)
    o = ol.Oracle(); o.load_tables(t.pod_ip, t.svc_ip); o.process(ev, 4)
    h = capi.Handle(max_endpoints=4 * S, max_pairs=1 << 16, max_batch=50_000)    # many slots in flight
    h.load_tables(t.pod_ip, t.svc_ip)
    parts = np.array_split(ev, T * 5)
    errs = []

    def work(k):
        try:
            for p in parts[k::T]:
                if packed:
                    r, v = capi.pack_l7(p)
                    h.submit_packed(r, v)
                else:
                    h.submit(p)
        except Exception as e:   # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(k,)) for k in range(T)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs
    got, exp = h.flush(), o.edges()
    assert edges_equal(got, exp), explain_diff(got, exp)
    assert h.stats()["events_in"] == N
    h.close()


def test_endpoint_table_patches_under_heavy_churn():
    """Every commit uploads only the changed slots (backward-shift deletion on the host mirror): after
    thousands of random ADD/UPDATE/DELETE the device table must resolve exactly like the reference's maps."""
    S = 400
    t = ol.Topo(S, seed=9, mix=abi.MIX_SURVEY)
    rng = np.random.default_rng(1)
    h = capi.Handle(max_endpoints=4 * S, max_pairs=1 << 15)
    o = ol.Oracle()
    h.load_tables(t.pod_ip, t.svc_ip); o.load_tables(t.pod_ip, t.svc_ip)
    ips = np.concatenate([t.pod_ip, t.svc_ip])
    for rnd in range(12):
        for _ in range(400):
            ip = int(ips[rng.integers(0, len(ips))])
            table = int(rng.integers(0, 2))
            if rng.random() < 0.5:
                h.erase(table, ip); o.erase(table, ip)
            else:
                i = int(rng.integers(0, 1 << 20))
                h.upsert(table, ip, i); o.upsert(table, ip, i)
        h.commit()
        ev = t.events(rnd * 150_000, 150_000)
        h.submit(ev); o.process(ev, 2)
    got, exp = h.flush(), o.edges()
    assert edges_equal(got, exp), explain_diff(got, exp)
    st, ost = h.stats(), o.stats()
    for k in ("events_in", "rows_emitted", "not_request", "src_unresolved"):
        assert st[k] == ost[k], (k, st[k], ost[k])
    h.close()


def test_first_window_and_shifting_traffic():
    """No history (first window) and a hot list that mispredicts (the traffic moves to another topology's pairs):
    results must not depend on what the per-CTA table happened to hold."""
    S = 1200
    ta, tb = ol.Topo(S, seed=101), ol.Topo(S, seed=202)
    h = capi.Handle(max_endpoints=8 * S, max_pairs=1 << 17)
    o = ol.Oracle()
    for x in (h, o):
        for k, v in enumerate(ta.pod_ip):
            x.upsert(abi.TABLE_POD, int(v), k)
        for k, v in enumerate(tb.pod_ip):
            x.upsert(abi.TABLE_POD, int(v), 100_000 + k)
        for k, v in enumerate(ta.svc_ip):
            x.upsert(abi.TABLE_SVC, int(v), k)
    h.commit()
    for w, topo in enumerate([ta, ta, tb, ta, tb]):
        ev = topo.events(w * 700_000, 700_000)
        h.submit(ev); o.process(ev, 4)
        got, exp = h.flush(), o.edges()
        assert edges_equal(got, exp), f"window {w}: " + explain_diff(got, exp)
        o.reset_window()
    h.close()


def test_host_header_keyed_outbound_nodes():
    """an HTTP destination that is neither service nor pod is keyed by the request's Host header
    (setFromToV2, aggregator/data.go:851-854). The caller (the Go adapter; here numpy) decides that from its own
    copy of the tables and marks the record ALZ_PROTO_F_HOSTKEY with the header's id in `daddr`."""
    S, N = 300, 300_000
    t = ol.Topo(S, seed=404, mix=abi.MIX_ALL)
    ev = t.events(0, N)
    rng = np.random.default_rng(7)
    names = [f"host{k}.example.org" for k in range(40)] + ["203.0.113.9", "8.8.8.8"]
    host_idx = np.where(rng.random(N) < 0.5, rng.integers(1, len(names) + 1, N), 0).astype(np.uint32)
    o = ol.Oracle(); o.load_tables(t.pod_ip, t.svc_ip)
    o.process_hosts(ev, host_idx, names)
    # what the adapter does with each event before it packs the record
    in_cluster = np.isin(ev["daddr"], np.concatenate([t.pod_ip, t.svc_ip]))
    use = (host_idx > 0) & (ev["protocol"] == abi.PROTO_HTTP) & ~in_cluster
    rec = ev.copy()
    is_ip = np.array([n[0].isdigit() for n in names])
    name_ip = np.array([abi.ip(n) if n[0].isdigit() else 0 for n in names], dtype=np.uint32)
    hid = host_idx.astype(np.int64) - 1
    as_ip = use & is_ip[np.maximum(hid, 0)]
    as_host = use & ~is_ip[np.maximum(hid, 0)]
    rec["daddr"][as_ip] = name_ip[hid[as_ip]]
    rec["daddr"][as_host] = hid[as_host]
    rec["protocol"][as_host] |= abi.PROTO_F_HOSTKEY
    assert as_host.sum() > 1000 and as_ip.sum() > 10
    for flags in (0, abi.CFG_EAGER_JOIN, abi.CFG_NO_SMEM_CACHE):
        h = capi.Handle(max_endpoints=4 * S, max_pairs=1 << 16, flags=flags)
        h.load_tables(t.pod_ip, t.svc_ip)
        h.submit(rec[: N // 2]); h.submit(rec[N // 2:])
        got, exp = h.flush(), o.edges()
        assert edges_equal(got, exp), explain_diff(got, exp)
        assert (got["to_type"] == abi.NODE_OUTBOUND_HOST).sum() > 50
        h.close()
    # the same through 16-byte packed records (the flag travels in the protocol byte)
    r16, ovf = capi.pack_l7(rec)
    h = capi.Handle(max_endpoints=4 * S, max_pairs=1 << 16)
    h.load_tables(t.pod_ip, t.svc_ip)
    h.submit_packed(r16, ovf)
    got = h.flush()
    assert edges_equal(got, o.edges()), explain_diff(got, o.edges())
    h.close()
