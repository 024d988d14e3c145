"""GPU: streaming windows (BASELINE configs[4] shape at test size): 1-second windows cut by
write_time, every window flushed and GNN-rescored, tables mutated between windows; plus
arithmetic edge cases of the accumulators (u64 wrap, 32-bit carry in the shared-memory rows)."""
import ctypes as C

import numpy as np
import pytest

import gnn_ref
import oracle_lib as ol
from alaz_b200 import abi, capi
from helpers import edges_equal, explain_diff

pytestmark = pytest.mark.gpu


def test_streaming_windows_with_rescore_and_table_churn():
    S, per_window, n_windows = 1500, 400_000, 6
    t = ol.Topo(S, seed=2024, mix=abi.MIX_ALL)
    h = capi.Handle(max_endpoints=4 * S, max_pairs=1 << 16)
    o = ol.Oracle()
    h.load_tables(t.pod_ip, t.svc_ip)
    o.load_tables(t.pod_ip, t.svc_ip)
    rng = np.random.default_rng(5)
    for w in range(n_windows):
        ev = t.events(w * per_window, per_window)
        # window = events of one second of write_time (the generator's clock is monotone: 100 ns per event)
        assert ev["write_time_ns"][0] <= ev["write_time_ns"][-1]
        cut = per_window // 3
        h.submit(ev[:cut]); o.process(ev[:cut])
        # informer churn in the middle of the window
        for k in rng.integers(0, len(t.pod_ip), 20):
            h.erase(abi.TABLE_POD, int(t.pod_ip[k])); o.erase(abi.TABLE_POD, int(t.pod_ip[k]))
        for k in rng.integers(0, len(t.pod_ip), 20):
            h.upsert(abi.TABLE_POD, int(t.pod_ip[k]), int(k)); o.upsert(abi.TABLE_POD, int(t.pod_ip[k]), int(k))
        h.commit()
        h.submit(ev[cut:]); o.process(ev[cut:])
        got, exp = h.flush(), o.edges()
        assert edges_equal(got, exp), f"window {w}: " + explain_diff(got, exp)
        o.reset_window()
        scores = np.zeros(len(got), dtype=np.float32)
        n = C.c_size_t(0)
        h._ck(h.L.alz_gnn_score(h.h, scores.ctypes.data_as(C.c_void_p), len(scores), C.byref(n)), "alz_gnn_score")
        _, _, ref = gnn_ref.run(got)
        err = np.abs(scores.astype(np.float64) - ref)
        assert np.all(err <= 1e-6 + 1e-5 * np.abs(ref)), (w, float(err.max()))
    st, ost = h.stats(), o.stats()
    for k in ("events_in", "rows_emitted", "not_request", "src_unresolved"):
        assert st[k] == ost[k], (k, st[k], ost[k])
    h.close()


@pytest.mark.parametrize("flags", [0, abi.CFG_EAGER_JOIN, abi.CFG_NO_SMEM_CACHE])
def test_accumulator_arithmetic_edges(flags):
    h = capi.Handle(max_endpoints=64, max_pairs=256, flags=flags)
    o = ol.Oracle()
    for x in (h, o):
        x.upsert(abi.TABLE_POD, abi.ip("10.0.0.1"), 1)
        x.upsert(abi.TABLE_SVC, abi.ip("172.16.0.1"), 2)
    h.commit()
    durs = [0xFFFFFFFF] * 40 + [0x1_0000_0000, 0xFFFF_FFFF_FFFF_FFFF, 1 << 63, 1 << 63, 3, 0, 255, 256,
                                 (1 << 40) - 1, 1 << 40, 0xFFFF_FFFF_0000_0001]
    recs = np.zeros(len(durs), dtype=abi.L7_REC)
    recs["saddr"] = abi.ip("10.0.0.1")
    recs["daddr"] = abi.ip("172.16.0.1")
    recs["protocol"] = abi.PROTO_HTTP
    recs["method_flags"] = 1
    recs["status"] = [500 + (i % 3) * 50 for i in range(len(durs))]     # 500, 550, 600 ...
    recs["status"][-1] = 65535                                            # saturated status
    recs["duration_ns"] = np.array(durs, dtype=np.uint64)
    # twice: the second submit finds the pair already hot (preloaded from the first fold)
    for rep in range(2):
        h.submit(recs); o.process(recs)
        got, exp = h.flush(), o.edges()
        assert edges_equal(got, exp), explain_diff(got, exp)
        assert int(got["lat_sum_ns"][0]) == sum(durs) % (1 << 64)
        o.reset_window()
    h.close()


def test_time_cut_windows_follow_the_records_clock():
    """SURVEY §8 row R13 / docs/SPEC.md §8: with alz_window_clock set, the window of a record is
    convertKernelTimeToUserspaceTime(write_time) / window_ns (aggregator/data.go:1740-1743), whatever the batch
    boundaries: records of later epochs wait on the device, each flush closes one epoch, late records join the
    open window and are counted."""
    S, N = 800, 1_200_000
    t = ol.Topo(S, seed=909, mix=abi.MIX_ALL)
    ev = t.events(0, N)
    wt = ev["write_time_ns"].astype(np.uint64)
    fk, fu = np.uint64(int(wt[0]) - 12345), np.uint64(1_700_000_000_123_456_789)
    W = np.uint64((int(wt[-1]) - int(wt[0])) // 4 + 1)
    epoch = (wt + (fu - fk)) // W
    assert ol.epoch(int(fk), int(fu), int(W), int(wt[777])) == int(epoch[777])
    e0, e_last = int(epoch.min()), int(epoch.max())
    assert epoch[0] == e0 and 3 <= e_last - e0 <= 5
    h = capi.Handle(max_endpoints=4 * S, max_pairs=1 << 16)
    h.load_tables(t.pod_ip, t.svc_ip)
    h.window_clock(int(fk), int(fu), int(W))
    # batches that ignore the window boundaries, one of them from device memory
    cuts = [0, 100_003, 350_000, 350_001, 900_017, N]
    for a, b in zip(cuts[:-1], cuts[1:]):
        if a == 350_001:
            d = h.dev_alloc((b - a) * 32)
            h.h2d(d, ev[a:b])
            h.submit_device(d, b - a)
            h.sync()
            h.dev_free(d)
        else:
            h.submit(ev[a:b])
    assert h.window_epoch() == e0
    st = h.stats()
    assert st["deferred_events"] == int((epoch > e0).sum()) and st["late_events"] == 0
    late = ev[:5000].copy()            # records of epoch e0 that show up after e0 was closed
    for k, e in enumerate(range(e0, e_last + 1)):
        got = h.flush()
        o = ol.Oracle(); o.load_tables(t.pod_ip, t.svc_ip)
        o.process(ev[epoch == e], 4)
        if k == 1:
            o.process(late)            # late records are reduced into the window that was open when they came
        exp = o.edges()
        assert edges_equal(got, exp), f"epoch {e}: " + explain_diff(got, exp)
        if k == 0:
            assert h.window_epoch() == e0 + 1
            h.submit(late)
            assert h.stats()["late_events"] == len(late)
    assert len(h.flush()) == 0
    st = h.stats()
    assert st["deferred_events"] == 0 and st["events_in"] == N + len(late)
    # packed records carry no time: refused while the clock is set; the clock can be switched off between windows
    r16, ovf = capi.pack_l7(ev[:10])
    with pytest.raises(capi.AlzError) as e:
        h.submit_packed(r16, ovf)
    assert e.value.status == abi.E_STATE
    h.window_clock(0, 0, 0)
    h.submit_packed(r16, ovf)
    assert int(h.flush()["count"].sum()) == h.stats()["rows_emitted"] - (st["rows_emitted"])
    h.close()
