"""GPU: tcp_state sink + temporal socket join (alz_submit_tcp / alz_sock_lookup) against the
SocketLine restatement, which is itself pinned by the reference's KATs (tests/test_oracle.py);
the same KATs are replayed here through the C ABI."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from alaz_b200 import abi, capi

pytestmark = pytest.mark.gpu


def _submit(h, recs):
    recs = np.ascontiguousarray(recs, dtype=abi.TCP_REC)
    h._ck(h.L.alz_submit_tcp(h.h, recs.ctypes.data_as(C.c_void_p), len(recs)), "alz_submit_tcp")


def _lookup(h, q):
    q = np.ascontiguousarray(q, dtype=abi.SOCK_QUERY)
    out = np.zeros(len(q), dtype=abi.SOCK_RESULT)
    h._ck(h.L.alz_sock_lookup(h.h, q.ctypes.data_as(C.c_void_p), len(q), out.ctypes.data_as(C.c_void_p)),
          "alz_sock_lookup")
    return out


def _ev(pid, fd, ts, typ, saddr=0, daddr=0, sport=0, dport=0):
    r = np.zeros(1, dtype=abi.TCP_REC)
    r[0] = (fd, ts, pid, saddr, daddr, sport, dport, typ, 0)
    return r


def test_reference_kats_through_the_abi():
    h = capi.Handle(max_endpoints=64, max_pairs=256)
    # sock_line_test.go:443-473 TestXxx2: later socket "xx" wins
    _submit(h, _ev(1, 0, 0, 1, saddr=0x7979))
    _submit(h, _ev(1, 0, 247453008321477, 1, saddr=0x7878))
    q = np.zeros(4, dtype=abi.SOCK_QUERY)
    q[0] = (0, 247453008321499, 1, 0)
    # sock_line_test.go:475-501 TestAlreadyEstablishCanBeFound: index-0 branch
    _submit(h, _ev(2, 5, 0, 1, saddr=0x7979))
    q[1] = (5, 0, 2, 0)
    # sock_line_test.go:11-349 TestSocketLine: identical opens collapse, lookup in between succeeds
    for ts in (33805065332163, 33805065990716, 33807002886899, 33945231235604, 33947004517045):
        _submit(h, _ev(3, 9, ts, 1))
    q[2] = (9, 33835107729129, 3, 0)
    q[3] = (77, 5, 99, 0)                     # unknown (pid, fd)
    r = _lookup(h, q)
    assert r[0]["found"] == 1 and r[0]["saddr"] == 0x7878
    assert r[1]["found"] == 1 and r[1]["saddr"] == 0x7979
    assert r[2]["found"] == 1
    assert r[3]["found"] == 0
    h.close()


def test_random_timelines_match_the_restatement():
    rng = np.random.default_rng(12345)
    n_lines, n_ev, n_q = 3000, 60000, 200000
    pids = rng.integers(1, 5000, n_lines).astype(np.uint32)
    fds = rng.integers(3, 200, n_lines).astype(np.uint64)
    ev = np.zeros(n_ev, dtype=abi.TCP_REC)
    line = rng.integers(0, n_lines, n_ev)
    ev["pid"], ev["fd"] = pids[line], fds[line]
    ev["timestamp_ns"] = rng.integers(0, 10**12, n_ev).astype(np.uint64)   # arrival order != time order
    ev["type"] = rng.choice([1, 5, 1, 5, 2, 3], n_ev)                       # some LISTEN / CONNECT_FAILED noise
    ev["saddr"] = 0x0A000000 + rng.integers(0, 50, n_ev)
    ev["daddr"] = 0x0A100000 + rng.integers(0, 4, n_ev)                     # few destinations: closed-gap rule fires
    ev["sport"] = rng.integers(30000, 30010, n_ev)
    ev["dport"] = rng.choice([80, 443], n_ev)
    lo = rng.random(n_ev) < 0.02
    ev["saddr"][lo] = 0x7F000001                                           # localhost: filtered
    dup = rng.random(n_ev) < 0.1                                           # duplicates: dedupe against the last element
    ev[1:][dup[1:]] = ev[:-1][dup[1:]]
    ev["timestamp_ns"][1:][dup[1:]] += 7
    q = np.zeros(n_q, dtype=abi.SOCK_QUERY)
    ql = rng.integers(0, n_lines, n_q)
    q["pid"], q["fd"] = pids[ql], fds[ql]
    q["timestamp_ns"] = rng.integers(0, 2 * 10**12, n_q).astype(np.uint64)
    miss = rng.random(n_q) < 0.05
    q["pid"][miss] = 999999
    o = ol.SockMaps()
    h = capi.Handle(max_endpoints=64, max_pairs=256)
    half = n_ev // 2
    for part in (ev[:half], ev[half:]):                                    # lookups between submits re-upload
        o.process(part)
        _submit(h, part)
        exp, got = o.lookup(q), _lookup(h, q)
        assert got.tobytes() == exp.tobytes()
    st = h.stats()
    assert st["tcp_events_in"] == n_ev
    assert st["tcp_localhost_dropped"] == o.localhost_dropped
    assert 0.2 < got["found"].mean() < 0.99
    h.close()


def _random_tcp(rng, n_lines, n_ev, t_hi=10**12, in_order=False):
    pids = rng.integers(1, 5000, n_lines).astype(np.uint32)
    fds = rng.integers(3, 200, n_lines).astype(np.uint64)
    ev = np.zeros(n_ev, dtype=abi.TCP_REC)
    line = rng.integers(0, n_lines, n_ev)
    ev["pid"], ev["fd"] = pids[line], fds[line]
    ts = rng.integers(0, t_hi, n_ev).astype(np.uint64)
    ev["timestamp_ns"] = np.sort(ts) if in_order else ts
    ev["type"] = rng.choice([1, 5], n_ev)
    ev["saddr"] = 0x0A000000 + rng.integers(0, 50, n_ev)
    ev["daddr"] = 0x0A100000 + rng.integers(0, 4, n_ev)
    ev["sport"] = rng.integers(30000, 30010, n_ev)
    ev["dport"] = rng.choice([80, 443], n_ev)
    return pids, fds, ev


def test_syncs_carry_only_the_inserts_since_the_last_one():
    """a lookup after new tcp events must not re-send every timeline. Mostly-in-order arrivals (what
    a kernel clock gives), many rounds: bytes copied per sync stay proportional to that round's events."""
    rng = np.random.default_rng(7)
    n_lines, rounds, per = 20000, 12, 4000
    pids, fds, ev = _random_tcp(rng, n_lines, rounds * per, in_order=True)
    q = np.zeros(50000, dtype=abi.SOCK_QUERY)
    ql = rng.integers(0, n_lines, len(q))
    q["pid"], q["fd"] = pids[ql], fds[ql]
    q["timestamp_ns"] = rng.integers(0, 10**12, len(q)).astype(np.uint64)
    o = ol.SockMaps()
    h = capi.Handle(max_endpoints=64, max_pairs=256)
    prev = h.sock_stats()
    per_round = []
    for r in range(rounds):
        part = ev[r * per:(r + 1) * per]
        o.process(part)
        h.submit_tcp(part)
        assert h.sock_lookup(q, now_ns=5).tobytes() == o.lookup(q, now_ns=5).tobytes()
        st = h.sock_stats()
        per_round.append((st["sync_bytes"] - prev["sync_bytes"], st["sync_ops"] - prev["sync_ops"]))
        prev = st
    st = h.sock_stats()
    assert st["syncs"] == rounds and st["pool_records"] >= o.records()
    total = o.records() * 32
    # after the index and pool have their size (the doublings are the big rounds), a sync costs its own inserts:
    # 48 B an insert + 40 B a touched line, nowhere near the full state
    tail = [b for b, ops in per_round[-4:]]
    assert max(tail) <= per * (48 + 40) + 4096, per_round
    assert max(tail) < total / 3, (per_round, total)
    h.close()


def test_out_of_order_inserts_and_segment_growth_on_the_device():
    """A few lines that grow to thousands of values with arrival order != time order: every insert shifts the tail
    of its line on the device, and the segments move as they outgrow their capacity."""
    rng = np.random.default_rng(8)
    pids, fds, ev = _random_tcp(rng, 5, 30000)
    ev["saddr"] = 0x0A000000 + rng.integers(0, 5000, len(ev))      # distinct sockets: the dedupe rarely fires
    q = np.zeros(100000, dtype=abi.SOCK_QUERY)
    ql = rng.integers(0, 5, len(q))
    q["pid"], q["fd"] = pids[ql], fds[ql]
    q["timestamp_ns"] = rng.integers(0, 10**12, len(q)).astype(np.uint64)
    o = ol.SockMaps()
    h = capi.Handle(max_endpoints=64, max_pairs=256)
    for part in np.array_split(ev, 40):
        o.process(part)
        h.submit_tcp(part)
        assert h.sock_lookup(q, now_ns=9).tobytes() == o.lookup(q, now_ns=9).tobytes()
    st = h.sock_stats()
    assert st["pool_garbage"] > 0 or st["repools"] > 0           # segments did move
    h.close()


def test_gc_follows_delete_unused_with_device_side_last_match():
    """alz_sock_gc = one tick of clearSocketLines: the LastMatch stamps are written by the lookups on the device and
    read back by the GC; afterwards lookups (and further inserts) still agree with the restatement."""
    rng = np.random.default_rng(9)
    M = 60 * 10**9
    n_lines = 2000
    pids, fds, ev = _random_tcp(rng, n_lines, 40000, t_hi=10**9)
    o = ol.SockMaps()
    h = capi.Handle(max_endpoints=64, max_pairs=256)
    o.process(ev)
    h.submit_tcp(ev)

    def queries(n, seed):
        r = np.random.default_rng(seed)
        q = np.zeros(n, dtype=abi.SOCK_QUERY)
        ql = r.integers(0, n_lines, n)
        q["pid"], q["fd"] = pids[ql], fds[ql]
        q["timestamp_ns"] = r.integers(0, 10**9, n).astype(np.uint64)
        return q
    # three batches of lookups at 1, 4 and 9 minutes: some pairs are last matched early, some late
    for k, now in enumerate((1 * M, 4 * M, 9 * M)):
        q = queries(15000, 100 + k)
        assert h.sock_lookup(q, now_ns=now).tobytes() == o.lookup(q, now_ns=now).tobytes()
    before = o.records()
    o.gc()
    h.sock_gc()
    assert o.records() < before
    q = queries(100000, 200)
    assert h.sock_lookup(q, now_ns=10 * M).tobytes() == o.lookup(q, now_ns=10 * M).tobytes()
    # life goes on: more events, another tick
    _, _, ev2 = _random_tcp(rng, n_lines, 1, t_hi=2 * 10**9)
    ev2 = np.zeros(20000, dtype=abi.TCP_REC)
    line = rng.integers(0, n_lines, len(ev2))
    ev2["pid"], ev2["fd"] = pids[line], fds[line]
    ev2["timestamp_ns"] = rng.integers(10**9, 2 * 10**9, len(ev2)).astype(np.uint64)
    ev2["type"] = rng.choice([1, 5], len(ev2))
    ev2["saddr"] = 0x0A000000 + rng.integers(0, 50, len(ev2))
    ev2["daddr"] = 0x0A100000 + rng.integers(0, 4, len(ev2))
    ev2["dport"] = 80
    o.process(ev2)
    h.submit_tcp(ev2)
    q["timestamp_ns"] = rng.integers(0, 2 * 10**9, len(q)).astype(np.uint64)
    assert h.sock_lookup(q, now_ns=16 * M).tobytes() == o.lookup(q, now_ns=16 * M).tobytes()
    o.gc()
    h.sock_gc()
    assert h.sock_lookup(q, now_ns=17 * M).tobytes() == o.lookup(q, now_ns=17 * M).tobytes()
    assert h.sock_stats()["pool_records"] >= o.records()
    h.close()


def test_l7_events_with_empty_5_tuples_are_joined_on_the_device():
    """alz_submit_l7_join: records whose addresses are zero take them from the (pid, fd) timeline at their write
    time, then go through the normal path; the window equals the oracle's over the restatement's joined records."""
    from helpers import edges_equal, explain_diff
    S, N = 300, 400_000
    t = ol.Topo(S, seed=5, mix=abi.MIX_ALL)
    ev = t.events(0, N)
    rng = np.random.default_rng(10)
    # a connection table: (pid, fd) -> the event's own addresses, opened before the event's write time
    conn = rng.integers(0, 20000, N)
    uc, first = np.unique(conn, return_index=True)
    tcp = np.zeros(len(uc), dtype=abi.TCP_REC)
    tcp["pid"] = 1000 + uc // 64
    tcp["fd"] = 3 + uc % 64
    tcp["timestamp_ns"] = 1                                       # opened at the beginning of time
    tcp["type"] = 1
    for f in ("saddr", "daddr", "sport", "dport"):
        tcp[f] = ev[f][first]
    keys = np.zeros(N, dtype=abi.SOCK_QUERY)
    keys["pid"] = 1000 + conn // 64
    keys["fd"] = 3 + conn % 64
    keys["timestamp_ns"] = np.maximum(ev["write_time_ns"], 2)
    blank = ev.copy()
    hide = rng.random(N) < 0.5                                    # half the events lost their 5-tuple in the kernel
    for f in ("saddr", "daddr", "sport", "dport"):
        blank[f][hide] = 0
    unknown = rng.random(N) < 0.02                                # ... and some have no timeline at all
    keys["pid"][unknown] = 7
    sm = ol.SockMaps()
    sm.process(tcp)
    joined, n_joined = sm.join(blank, keys, now_ns=3)
    o = ol.Oracle()
    o.load_tables(t.pod_ip, t.svc_ip)
    o.process(joined)
    exp = o.edges()
    h = capi.Handle(max_endpoints=4 * S, max_pairs=1 << 16, max_batch=1 << 17)   # several chunks
    h.load_tables(t.pod_ip, t.svc_ip)
    h.submit_tcp(tcp)
    h.submit_join(blank, keys, now_ns=3)
    got = h.flush()
    assert edges_equal(got, exp), explain_diff(got, exp)
    assert h.sock_stats()["joined_events"] == n_joined and n_joined > N // 3
    st, ost = h.stats(), o.stats()
    assert st["src_unresolved"] == ost["src_unresolved"] and st["src_unresolved"] > 0
    h.close()


def test_alive_connections_export():
    """alz_sock_alive = sendOpenConnection over every line (data.go:1628-1679)."""
    S = 200
    t = ol.Topo(S, seed=6, mix=abi.MIX_ALL)
    rng = np.random.default_rng(11)
    n = 30000
    ev = np.zeros(n, dtype=abi.TCP_REC)
    ev["pid"] = rng.integers(1, 3000, n)
    ev["fd"] = rng.integers(3, 40, n)
    ev["timestamp_ns"] = np.sort(rng.integers(1, 10**12, n)).astype(np.uint64)
    ev["type"] = rng.choice([1, 1, 5], n)
    pods = np.asarray(t.pod_ip)
    dst = np.concatenate([np.asarray(t.svc_ip), pods[:50], 0x08080000 + np.arange(50, dtype=np.uint32)])
    src = np.concatenate([pods, 0x0B000000 + np.arange(20, dtype=np.uint32)])   # some sources are no pods
    ev["saddr"] = src[rng.integers(0, len(src), n)]
    ev["daddr"] = dst[rng.integers(0, len(dst), n)]
    ev["sport"] = rng.integers(30000, 60000, n)
    ev["dport"] = rng.choice([80, 443, 5432], n)
    sm = ol.SockMaps()
    sm.process(ev)
    o = ol.Oracle()
    o.load_tables(t.pod_ip, t.svc_ip)
    exp = sm.alive(o)
    h = capi.Handle(max_endpoints=4 * S, max_pairs=256)
    h.load_tables(t.pod_ip, t.svc_ip)
    h.submit_tcp(ev)
    got = h.sock_alive()
    assert len(got) == len(exp) > 1000
    assert {int(x) for x in np.unique(got["to_type"])} == {abi.NODE_POD, abi.NODE_SVC, abi.NODE_OUTBOUND}
    assert np.sort(got, order=list(abi.ALIVE_CONN.names)[:-1]).tobytes() == \
        np.sort(exp, order=list(abi.ALIVE_CONN.names)[:-1]).tobytes()
    # too small a buffer: status + the number there are
    small = np.zeros(10, dtype=abi.ALIVE_CONN)
    k = C.c_size_t(0)
    rc = h.L.alz_sock_alive(h.h, small.ctypes.data_as(C.c_void_p), 10, C.byref(k))
    assert rc == abi.E_CAPACITY and k.value == len(exp)
    h.close()


def test_raw_tcp_samples_equal_decoded_ones():
    """alz_submit_tcp_raw takes struct tcp_event as the ring carries it (64 B, address bytes first octet first)."""
    rng = np.random.default_rng(13)
    pids, fds, ev = _random_tcp(rng, 500, 20000)
    raw = np.zeros(len(ev), dtype=abi.BPF_TCP_EVENT)
    raw["fd"], raw["timestamp"], raw["type"], raw["pid"] = ev["fd"], ev["timestamp_ns"], ev["type"], ev["pid"]
    raw["sport"], raw["dport"] = ev["sport"], ev["dport"]
    for f in ("saddr", "daddr"):
        for k in range(4):
            raw[f][:, k] = (ev[f] >> (24 - 8 * k)) & 0xFF
        raw[f][:, 4:] = 0xEE                                    # the other twelve bytes are not read
    q = np.zeros(50000, dtype=abi.SOCK_QUERY)
    ql = rng.integers(0, 500, len(q))
    q["pid"], q["fd"] = pids[ql], fds[ql]
    q["timestamp_ns"] = rng.integers(0, 10**12, len(q)).astype(np.uint64)
    a = capi.Handle(max_endpoints=64, max_pairs=256)
    b = capi.Handle(max_endpoints=64, max_pairs=256)
    a.submit_tcp(ev)
    b._ck(b.L.alz_submit_tcp_raw(b.h, raw.ctypes.data_as(C.c_void_p), len(raw)), "alz_submit_tcp_raw")
    ra, rb = a.sock_lookup(q, now_ns=1), b.sock_lookup(q, now_ns=1)
    assert ra.tobytes() == rb.tobytes() and ra["found"].sum() > 1000
    assert a.stats()["tcp_events_in"] == b.stats()["tcp_events_in"] == len(ev)
    a.close(); b.close()
