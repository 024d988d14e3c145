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
