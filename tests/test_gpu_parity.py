"""GPU parity: the CUDA path, driven through the C ABI, against the CPU oracle —
bit-exact per-edge integers on the same inputs (hand-derived vectors, seeded
synthetic streams), plus size-independent properties at BASELINE size."""
import numpy as np
import pytest

import oracle_lib as ol
from alaz_b200 import abi, capi
from helpers import load_branches, edges_equal, explain_diff

pytestmark = pytest.mark.gpu

PLANS = [0, abi.CFG_EAGER_JOIN, abi.CFG_NO_SMEM_CACHE]


def _oracle_for(topo, events, nthreads=4):
    o = ol.Oracle()
    o.load_tables(topo.pod_ip, topo.svc_ip)
    o.process(events, nthreads)
    return o


@pytest.mark.parametrize("flags", PLANS)
def test_hand_derived_branch_vectors(flags):
    pods, svcs, recs, exp, exp_stats = load_branches()
    h = capi.Handle(max_endpoints=64, max_pairs=1024, flags=flags)
    for ip, i in pods.items():
        h.upsert(abi.TABLE_POD, ip, i)
    for ip, i in svcs.items():
        h.upsert(abi.TABLE_SVC, ip, i)
    h.commit()
    h.submit(recs)
    h.fold()
    st = h.stats()
    got = h.flush()
    assert edges_equal(got, exp), explain_diff(got, exp)
    for k, v in exp_stats.items():
        assert st[k] == v, (k, st[k], v)
    h.close()


@pytest.mark.parametrize("flags", PLANS)
@pytest.mark.parametrize("mix", [abi.MIX_SURVEY, abi.MIX_ALL])
def test_synthetic_stream_bit_exact_vs_oracle(flags, mix):
    S, N = 1000, 2_000_000
    t = ol.Topo(S, seed=0xA1A20001 + mix, mix=mix)
    ev = t.events(0, N)
    o = _oracle_for(t, ev)
    h = capi.Handle(max_endpoints=4 * S, max_pairs=1 << 16, max_batch=300_000, flags=flags)  # forces chunking
    h.load_tables(t.pod_ip, t.svc_ip)
    h.submit(ev)
    h.fold()
    st, ost = h.stats(), o.stats()
    got, exp = h.flush(), o.edges()
    assert edges_equal(got, exp), explain_diff(got, exp)
    for k in ("events_in", "rows_emitted", "not_request", "src_unresolved"):
        assert st[k] == ost[k], (k, st[k], ost[k])
    assert int(got["count"].sum()) == st["rows_emitted"]
    assert np.array_equal(got["hist"].sum(axis=1), got["count"])
    h.close()


def test_device_generator_equals_host_generator():
    t = capi.Topo(500, seed=77, mix=abi.MIX_ALL)
    h = capi.Handle()
    n = 1_000_003
    d = h.dev_alloc(n * 32)
    t.fill_device(h, 11, n, d)
    h.sync()
    dev = h.d2h(d, n, abi.L7_REC)
    assert dev.tobytes() == t.events(11, n).tobytes()
    h.dev_free(d)
    t.close()
    h.close()


@pytest.mark.parametrize("flags", PLANS)
def test_hbm_resident_submit_and_windows(flags):
    S, N = 2000, 3_000_000
    ot = ol.Topo(S, seed=5)
    t = capi.Topo(S, seed=5)
    h = capi.Handle(max_endpoints=4 * S, max_pairs=1 << 17, flags=flags)
    h.load_tables(t.pod_ip, t.svc_ip)
    d = h.dev_alloc(N * 32)
    t.fill_device(h, 0, N, d)
    # window 1: first 2M events in two submits; window 2: the rest
    h.submit_device(d, 1_000_000)
    h.submit_device(d + 1_000_000 * 32, 1_000_000)
    w1 = h.flush()
    h.submit_device(d + 2_000_000 * 32, N - 2_000_000)
    w2 = h.flush()
    w3 = h.flush()   # nothing submitted: empty window
    o = ol.Oracle()
    o.load_tables(ot.pod_ip, ot.svc_ip)
    o.process(ot.events(0, 2_000_000), 4)
    assert edges_equal(w1, o.edges()), explain_diff(w1, o.edges())
    o.reset_window()
    o.process(ot.events(2_000_000, N - 2_000_000), 4)
    assert edges_equal(w2, o.edges()), explain_diff(w2, o.edges())
    assert len(w3) == 0
    h.dev_free(d)
    t.close()
    h.close()


@pytest.mark.parametrize("flags", PLANS)
def test_table_changes_between_submits_resolve_like_the_reference(flags):
    # events are resolved by the tables in force when they were submitted
    # (persist.go:55-71 / :114-130 mutate the maps between events)
    S = 200
    t = ol.Topo(S, seed=3, mix=abi.MIX_ALL)
    a, b, c = t.events(0, 100_000), t.events(100_000, 100_000), t.events(200_000, 100_000)
    o = ol.Oracle()
    h = capi.Handle(max_endpoints=4 * S, max_pairs=1 << 15, flags=flags)
    o.load_tables(t.pod_ip, t.svc_ip)
    h.load_tables(t.pod_ip, t.svc_ip)
    o.process(a); h.submit(a)
    # DELETE some pods and services, UPDATE a pod's UID, ADD a service on a pod IP
    for k in range(0, 40):
        o.erase(abi.TABLE_POD, int(t.pod_ip[k])); h.erase(abi.TABLE_POD, int(t.pod_ip[k]))
    for k in range(0, 30):
        o.erase(abi.TABLE_SVC, int(t.svc_ip[k])); h.erase(abi.TABLE_SVC, int(t.svc_ip[k]))
    o.upsert(abi.TABLE_POD, int(t.pod_ip[50]), 9000); h.upsert(abi.TABLE_POD, int(t.pod_ip[50]), 9000)
    o.upsert(abi.TABLE_SVC, int(t.pod_ip[60]), 9001); h.upsert(abi.TABLE_SVC, int(t.pod_ip[60]), 9001)
    h.commit()
    o.process(b); h.submit(b)
    for k in range(0, 40):   # ADD them back under new ids
        o.upsert(abi.TABLE_POD, int(t.pod_ip[k]), 5000 + k); h.upsert(abi.TABLE_POD, int(t.pod_ip[k]), 5000 + k)
    h.commit()
    o.process(c); h.submit(c)
    got, exp = h.flush(), o.edges()
    assert edges_equal(got, exp), explain_diff(got, exp)
    st, ost = h.stats(), o.stats()
    for k in ("events_in", "rows_emitted", "not_request", "src_unresolved"):
        assert st[k] == ost[k], (k, st[k], ost[k])
    h.close()


def _to_raw(recs):
    """compact records -> 1096-B struct l7_event samples (ebpf/l7_req/l7.go:345-369)."""
    n = len(recs)
    raw = np.zeros((n, abi.BPF_L7_EVENT_SIZE), dtype=np.uint8)
    def put(off, arr, dt):
        b = np.ascontiguousarray(arr.astype(dt)).view(np.uint8).reshape(n, -1)
        raw[:, off:off + b.shape[1]] = b
    put(0, np.arange(n) % 97, "<u8")                       # fd
    put(8, recs["write_time_ns"], "<u8")
    put(16, np.arange(n) % 4001, "<u4")                    # pid
    put(20, recs["status"], "<u4")
    put(24, recs["duration_ns"], "<u8")
    raw[:, 32] = recs["protocol"]
    raw[:, 33] = recs["method_flags"] & abi.MF_METHOD_MASK
    raw[:, 36:36 + 16] = np.frombuffer(b"GET /user HTTP1.", dtype=np.uint8)   # payload is ignored
    raw[:, 1066] = (recs["method_flags"] & abi.MF_TLS) != 0
    put(1076, recs["saddr"], "<u4")
    put(1080, recs["sport"], "<u2")
    put(1084, recs["daddr"], "<u4")
    put(1088, recs["dport"], "<u2")
    return raw.reshape(-1)


def test_raw_perf_samples_ingest():
    S, N = 300, 200_000
    t = ol.Topo(S, seed=21, mix=abi.MIX_SURVEY)
    ev = t.events(0, N)
    raw = _to_raw(ev)
    assert ol.compact_raw(raw).tobytes() == ev.tobytes()
    o = _oracle_for(t, ev)
    h = capi.Handle(max_endpoints=4 * S, max_pairs=1 << 15, max_batch=50_000)
    h.load_tables(t.pod_ip, t.svc_ip)
    h.submit_raw(raw)
    got = h.flush()
    assert edges_equal(got, o.edges()), explain_diff(got, o.edges())
    h.close()


def test_edge_cases_empty_single_ragged_and_sentinel_pair():
    h = capi.Handle(max_endpoints=64, max_pairs=256)
    o = ol.Oracle()
    for hh in (h, o):
        hh.upsert(abi.TABLE_POD, 0xFFFFFFFF, 7)        # 255.255.255.255 as a pod: collides with the
        hh.upsert(abi.TABLE_POD, abi.ip("10.0.0.1"), 1)  # dictionary's empty marker
    h.commit()
    h.submit(np.zeros(0, dtype=abi.L7_REC))
    assert len(h.flush()) == 0
    recs = np.zeros(37, dtype=abi.L7_REC)               # ragged: not a multiple of a warp
    recs["protocol"] = abi.PROTO_HTTP
    recs["method_flags"] = 1
    recs["saddr"] = 0xFFFFFFFF
    recs["daddr"] = 0xFFFFFFFF
    recs["status"] = 500
    recs["duration_ns"] = np.arange(37, dtype=np.uint64) * 1_000_003
    recs[5]["saddr"] = abi.ip("10.0.0.1")
    recs[6]["daddr"] = abi.ip("10.0.0.1")
    h.submit(recs[:1]); h.submit(recs[1:])
    o.process(recs)
    got, exp = h.flush(), o.edges()
    assert edges_equal(got, exp), explain_diff(got, exp)
    assert len(got) == 3
    h.close()


def test_capacity_is_reported_not_silent_and_covers_one_window():
    """A full pair/edge table is reported by the flush of THAT window; the handle keeps working and the
    next, healthy window returns ALZ_OK with exact edges (r1 returned ALZ_E_CAPACITY for ever after)."""
    t = ol.Topo(500, seed=8)
    ev = t.events(0, 300_000)
    for max_pairs, max_edges in ((64, 64), (1 << 15, 64)):   # pair table too small; only the edge table too small
        h = capi.Handle(max_endpoints=4096, max_pairs=max_pairs, max_edges=max_edges)
        h.load_tables(t.pod_ip, t.svc_ip)
        h.submit(ev)
        with pytest.raises(capi.AlzError) as e:
            h.flush()
        assert e.value.status == abi.E_CAPACITY
        assert h.stats()["capacity_events"] > 0
        small = ev[ev["saddr"] == ev["saddr"][0]][:2000]      # a handful of pairs: fits
        o = ol.Oracle(); o.load_tables(t.pod_ip, t.svc_ip); o.process(small)
        h.submit(small)
        got = h.flush()                                       # ALZ_OK again
        assert edges_equal(got, o.edges()), explain_diff(got, o.edges())
        h.close()


def test_full_size_config2_bit_exact_and_properties():
    """BASELINE configs[1]: 10k services / 100M events on one GPU, at full size.
    (1) per-edge bit-exact against the oracle run with every host core on the very same 100M records;
    (2) size-independent properties: conservation of rows, count == sum(hist), total latency == an
    independent numpy pass over the stream; (3) window linearity (A then B in one window == whole)."""
    import os
    S, N, CH = 10_000, 100_000_000, 10_000_000
    t = capi.Topo(S, seed=0xA1A20000 + 1)
    h = capi.Handle(max_endpoints=4 * S, max_pairs=1 << 20)
    h.load_tables(t.pod_ip, t.svc_ip)
    d = h.dev_alloc(N * 32)
    t.fill_device(h, 0, N, d)
    h.submit_device(d, N)
    edges = h.flush()
    st = h.stats()
    assert st["events_in"] == N
    assert int(edges["count"].sum()) == st["rows_emitted"] == N - st["not_request"] - st["src_unresolved"]
    assert np.array_equal(edges["hist"].sum(axis=1), edges["count"])
    # the same records on the host (the device generator is bit-identical to the host one, tested above)
    ev_all = np.empty(N, dtype=abi.L7_REC)
    pods = np.sort(t.pod_ip)
    lat = rows = err = 0
    for c in range(0, N, CH):
        ev = h.d2h(d + c * 32, CH, abi.L7_REC)
        ev_all[c:c + CH] = ev
        emit = np.isin(ev["protocol"], [abi.PROTO_HTTP, abi.PROTO_AMQP, abi.PROTO_REDIS])
        idx = np.searchsorted(pods, ev["saddr"])
        idx[idx == len(pods)] = 0
        known = pods[idx] == ev["saddr"]
        m = emit & known
        rows += int(m.sum())
        lat += int(ev["duration_ns"][m].sum(dtype=np.uint64))
        err += int((m & (ev["protocol"] == abi.PROTO_HTTP) & (ev["status"] >= 500) & (ev["status"] < 600)).sum())
    assert rows == st["rows_emitted"]
    assert lat % (1 << 64) == int(edges["lat_sum_ns"].sum(dtype=np.uint64))
    assert err == int(edges["err5xx"].sum())
    # (1) the oracle on all 100M records
    o = ol.Oracle()
    o.load_tables(t.pod_ip, t.svc_ip)
    o.process(ev_all, max(1, os.cpu_count() or 1))
    exp = o.edges()
    assert edges_equal(edges, exp), explain_diff(edges, exp)
    ost = o.stats()
    for k in ("events_in", "rows_emitted", "not_request", "src_unresolved"):
        assert st[k] == ost[k], (k, st[k], ost[k])
    o.close()
    del ev_all
    # (3) linearity: two halves submitted separately into one window give the same edges
    h.submit_device(d, N // 2)
    h.submit_device(d + (N // 2) * 32, N - N // 2)
    again = h.flush()
    assert edges_equal(again, edges)
    h.dev_free(d)
    t.close()
    h.close()


@pytest.mark.parametrize("flags", PLANS)
def test_committed_golden_fixture(flags):
    """tests/golden/synth_small.npz: fixed vectors, independent of today's oracle build."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "synth_small.npz"))
    h = capi.Handle(max_endpoints=1024, max_pairs=1 << 14, flags=flags)
    h.load_tables(z["pod_ip"], z["svc_ip"])
    ev = z["events"].view(abi.L7_REC)
    h.submit(ev[:7777]); h.submit(ev[7777:])
    got = h.flush()
    exp = z["edges"].view(abi.EDGE_OUT)
    assert edges_equal(got, exp), explain_diff(got, exp)
    st = h.stats()
    assert [st["events_in"], st["rows_emitted"], st["not_request"], st["src_unresolved"]] == z["stats"].tolist()
    h.close()
