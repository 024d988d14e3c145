"""CPU tests of the oracle itself: the oracle is only trusted after it passes
the hand-derived branch vectors, agrees with the independent Python
restatement, and reproduces the reference's own SocketLine KATs."""
import ctypes as C
import sys, os

import numpy as np
import pytest

import oracle_lib as ol
from alaz_b200 import abi
from helpers import load_branches, edges_equal, explain_diff, pyref_edges

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import ref_py  # noqa: E402


def _c_oracle(pods, svcs):
    o = ol.Oracle()
    for ip, i in pods.items():
        o.upsert(abi.TABLE_POD, ip, i)
    for ip, i in svcs.items():
        o.upsert(abi.TABLE_SVC, ip, i)
    return o


def _py_oracle(pods, svcs):
    a = ref_py.Aggregator()
    for ip, i in pods.items():
        a.pod_ip_to_uid[ref_py.ip_string(ip)] = "pod-%d" % i
    for ip, i in svcs.items():
        a.svc_ip_to_uid[ref_py.ip_string(ip)] = "svc-%d" % i
    return a


def test_c_oracle_matches_hand_derived_branches():
    pods, svcs, recs, exp, exp_stats = load_branches()
    o = _c_oracle(pods, svcs)
    o.process(recs)
    got = o.edges()
    assert edges_equal(got, exp), explain_diff(got, exp)
    st = o.stats()
    for k, v in exp_stats.items():
        assert st[k] == v, (k, st[k], v)


def test_py_restatement_matches_hand_derived_branches():
    pods, svcs, recs, exp, exp_stats = load_branches()
    a = _py_oracle(pods, svcs)
    for r in recs:
        a.process_l7(r)
    got = pyref_edges(a)
    assert edges_equal(got, exp), explain_diff(got, exp)
    assert a.stats == exp_stats


@pytest.mark.parametrize("mix", [abi.MIX_SURVEY, abi.MIX_ALL])
def test_c_oracle_matches_py_restatement_on_synthetic(mix):
    t = ol.Topo(50, seed=1234 + mix, mix=mix)
    ev = t.events(0, 20000)
    o = ol.Oracle()
    o.load_tables(t.pod_ip, t.svc_ip)
    o.process(ev)
    a = ref_py.Aggregator()
    for k, v in enumerate(t.pod_ip):
        a.pod_ip_to_uid[ref_py.ip_string(int(v))] = "pod-%d" % k
    for k, v in enumerate(t.svc_ip):
        a.svc_ip_to_uid[ref_py.ip_string(int(v))] = "svc-%d" % k
    for r in ev:
        a.process_l7(r)
    got, exp = o.edges(), pyref_edges(a)
    assert edges_equal(got, exp), explain_diff(got, exp)
    st = o.stats()
    for k, v in a.stats.items():
        assert st[k] == v
    assert st["rows_emitted"] == int(got["count"].sum())


def test_multithreaded_oracle_is_identical():
    t = ol.Topo(100, seed=7, mix=abi.MIX_ALL)
    ev = t.events(0, 200000)
    a, b = ol.Oracle(), ol.Oracle()
    a.load_tables(t.pod_ip, t.svc_ip)
    b.load_tables(t.pod_ip, t.svc_ip)
    a.process(ev, 1)
    b.process(ev, 4)
    assert a.edges().tobytes() == b.edges().tobytes()
    assert a.stats() == b.stats()


@pytest.mark.parametrize("mix", [abi.MIX_SURVEY, abi.MIX_ALL])
def test_fair_cpu_arm_is_bit_exact_to_the_faithful_port(mix):
    # oracle/alz_fastcpu.c (bench's second cpu_baseline) against oracle/alz_oracle.c, single- and multi-threaded
    t = ol.Topo(300, seed=99 + mix, mix=mix)
    ev = t.events(0, 300_000)
    o = ol.Oracle(); o.load_tables(t.pod_ip, t.svc_ip); o.process(ev, 3)
    for nt in (1, 5):
        f = ol.FastCpu(4 * 300); f.load_tables(t.pod_ip, t.svc_ip); f.process(ev, nt)
        got, exp = f.edges(), o.edges()
        assert edges_equal(got, exp), explain_diff(got, exp)
        fs, os_ = f.stats(), o.stats()
        for k in ("events_in", "rows_emitted", "not_request", "src_unresolved"):
            assert fs[k] == os_[k]
        f.close()


HOST_VECTORS = [   # parseHttpPayload's hostHeader, aggregator/data.go:508-531 (same vectors: tests/cpp/host_unit_test.cc)
    (b"GET /x HTTP/1.1\r\nHost: example.com\r\nAccept: */*\r\n\r\n", "example.com"),
    (b"GET /x HTTP/1.1\nHost: a.b:8080\n\n", "a.b:8080"),
    (b"GET /x HTTP/1.1\r\nAccept: */*\r\n\r\n", ""),
    (b"Host: first.line\r\nX: y\r\n", ""),
    (b"GET / HTTP/1.1\r\nHost:nospace.com\r\nHost: second.com\r\n", "second.com"),
    (b"GET / HTTP/1.1\r\nHost:  two.spaces\r\n", ""),
    (b"GET / HTTP/1.1\r\nhost: lower.case\r\n", ""),
    (b"GET / HTTP/1.1\r\nHost: h.com extra words\r\n", "h.com"),
    (b"GET / HTTP/1.1\r\nX-Host: no\r\nHost: yes.com", "yes.com"),
    (b"", ""),
]


def test_parse_http_host_header_vectors():
    for payload, want in HOST_VECTORS:
        assert ol.parse_http_host(payload) == want, payload


def test_host_header_keys_outbound_destinations_only():
    # setFromToV2 :838-866: service first, then pod, then the Host header, then the raw daddr
    o = ol.Oracle()
    o.upsert(abi.TABLE_POD, abi.ip("10.0.0.1"), 1)
    o.upsert(abi.TABLE_POD, abi.ip("10.0.0.2"), 2)
    o.upsert(abi.TABLE_SVC, abi.ip("172.16.0.1"), 7)
    names = ["api.example.com", "8.8.4.4"]
    recs = np.zeros(6, dtype=abi.L7_REC)
    recs["saddr"] = abi.ip("10.0.0.1")
    recs["protocol"] = abi.PROTO_HTTP
    recs["method_flags"] = 1
    recs["duration_ns"] = 1000
    recs["daddr"] = [abi.ip(x) for x in ("9.9.9.9", "9.9.9.9", "172.16.0.1", "10.0.0.2", "9.9.9.9", "8.8.4.4")]
    recs[4]["protocol"] = abi.PROTO_REDIS            # only HTTP rows carry a Host header
    host_idx = np.array([1, 0, 1, 1, 1, 2], dtype=np.uint32)
    o.process_hosts(recs, host_idx, names)
    e = o.edges()
    got = {(int(x["to_type"]), int(x["to"])): int(x["count"]) for x in e}
    assert got == {
        (abi.NODE_OUTBOUND_HOST, 0): 1,                 # 9.9.9.9 with Host api.example.com
        (abi.NODE_OUTBOUND, abi.ip("9.9.9.9")): 2,      # no header; and the REDIS row
        (abi.NODE_SVC, 7): 1, (abi.NODE_POD, 2): 1,     # the header is ignored for in-cluster destinations
        (abi.NODE_OUTBOUND, abi.ip("8.8.4.4")): 1,      # a header that is a dotted quad = that raw daddr node
    }


def test_epoch_is_the_references_time_conversion():
    # convertKernelTimeToUserspaceTime (data.go:1740-1743): FirstUserspaceTime - (FirstKernelTime - t), uint64
    fk, fu, w = 5_000_000_000, 1_700_000_000_000_000_000, 1_000_000_000
    assert ol.epoch(fk, fu, w, fk) == fu // w
    assert ol.epoch(fk, fu, w, fk + 999_999_999 - (fu % w)) == fu // w
    assert ol.epoch(fk, fu, w, fk + w - (fu % w)) == fu // w + 1
    assert ol.epoch(fk, fu, w, fk - 1) == (fu - 1) // w          # events older than the first sample


def test_table_erase_and_update_follow_persist_go():
    # persist.go:55-71: UPDATE overwrites, DELETE removes
    o = ol.Oracle()
    o.upsert(abi.TABLE_POD, abi.ip("10.1.1.1"), 5)
    o.upsert(abi.TABLE_SVC, abi.ip("172.16.1.1"), 9)
    rec = np.zeros(1, dtype=abi.L7_REC)
    rec[0] = (abi.ip("10.1.1.1"), abi.ip("172.16.1.1"), 1, 2, 200, abi.PROTO_HTTP, 1, 10, 0)
    o.process(rec)
    o.upsert(abi.TABLE_POD, abi.ip("10.1.1.1"), 6)      # UPDATE: same IP, new UID
    o.process(rec)
    o.erase(abi.TABLE_SVC, abi.ip("172.16.1.1"))        # DELETE: now outbound by raw IP
    o.process(rec)
    o.erase(abi.TABLE_POD, abi.ip("10.1.1.1"))          # source gone: dropped
    o.process(rec)
    e = o.edges()
    keys = {(int(x["from_type"]), int(x["from"]), int(x["to_type"]), int(x["to"])) for x in e}
    assert keys == {(0, 5, 1, 9), (0, 6, 1, 9), (0, 6, 2, abi.ip("172.16.1.1"))}
    assert o.stats()["src_unresolved"] == 1


def test_bucket_function_edges():
    assert ol.bucket(0) == 0 and ol.bucket(255) == 0 and ol.bucket(256) == 0
    assert ol.bucket(383) == 0 and ol.bucket(384) == 1 and ol.bucket(511) == 1
    assert ol.bucket(512) == 2 and ol.bucket(767) == 2 and ol.bucket(768) == 3
    assert ol.bucket((1 << 40) - 1) == 63 and ol.bucket(1 << 40) == 63
    assert ol.bucket((1 << 64) - 1) == 63
    for d in [1, 300, 1000, 123456, 2_000_000, 10**9, 10**12, 10**15]:
        assert ol.bucket(d) == ref_py.bucket(d)
    prev = 0
    for sh in range(0, 50):
        for m in (2, 3):
            b = ol.bucket(m << sh)
            assert b >= prev
            prev = b


def test_quantile_interpolation():
    h = np.zeros(64, dtype=np.uint32)
    h[4] = 10           # [1024, 1536)
    assert ol.quantile(h, 0.5) == pytest.approx(1024 + 0.5 * 512, rel=1e-12)
    assert ol.quantile(h, 1.0) == pytest.approx(1536, rel=1e-12)
    h[5] = 10           # [1536, 2048)
    assert ol.quantile(h, 0.75) == pytest.approx(1536 + 0.5 * 512, rel=1e-12)
    assert ol.quantile(np.zeros(64, dtype=np.uint32), 0.5) == 0.0


def test_compact_raw_reads_go_struct_offsets():
    # bpfL7Event, ebpf/l7_req/l7.go:345-369
    raw = np.zeros(2 * abi.BPF_L7_EVENT_SIZE, dtype=np.uint8)
    v = raw[abi.BPF_L7_EVENT_SIZE:]
    v[8:16] = np.frombuffer(np.uint64(111).tobytes(), np.uint8)       # WriteTimeNs
    v[20:24] = np.frombuffer(np.uint32(70000).tobytes(), np.uint8)    # Status (saturates)
    v[24:32] = np.frombuffer(np.uint64(222).tobytes(), np.uint8)      # Duration
    v[32], v[33] = 5, 2                                               # Protocol, Method
    v[1066] = 1                                                       # IsTls
    v[1076:1080] = np.frombuffer(np.uint32(abi.ip("10.0.0.1")).tobytes(), np.uint8)
    v[1080:1082] = np.frombuffer(np.uint16(4444).tobytes(), np.uint8)
    v[1084:1088] = np.frombuffer(np.uint32(abi.ip("10.0.0.2")).tobytes(), np.uint8)
    v[1088:1090] = np.frombuffer(np.uint16(6379).tobytes(), np.uint8)
    out = ol.compact_raw(raw)
    assert out[0].tobytes() == bytes(32)
    r = out[1]
    assert (int(r["saddr"]), int(r["daddr"]), int(r["sport"]), int(r["dport"])) == (
        abi.ip("10.0.0.1"), abi.ip("10.0.0.2"), 4444, 6379)
    assert int(r["status"]) == 65535 and int(r["protocol"]) == 5
    assert int(r["method_flags"]) == (2 | abi.MF_TLS)
    assert int(r["duration_ns"]) == 222 and int(r["write_time_ns"]) == 111


# ---------------- SocketLine KATs: the reference's own tests ----------------
class _SI(C.Structure):
    _fields_ = [("saddr", C.c_uint32), ("daddr", C.c_uint32), ("sport", C.c_uint16), ("dport", C.c_uint16)]


class _Line:
    def __init__(self):
        self.L = ol.lib()
        self.h = self.L.orc_sockline_create()

    def add(self, ts, si):
        self.L.orc_sockline_add(self.h, ts, C.byref(si) if si is not None else None)

    def get(self, ts):
        out = _SI()
        ok = self.L.orc_sockline_get(self.h, ts, C.byref(out))
        return out if ok else None

    def get_at(self, ts, now):
        out = _SI()
        ok = self.L.orc_sockline_get_at(self.h, ts, now, C.byref(out))
        return out if ok else None

    def delete_unused(self):
        self.L.orc_sockline_delete_unused(self.h)

    def __len__(self):
        return self.L.orc_sockline_len(self.h)


def _sockline_kat():
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sockline_kat.json")) as f:
        return json.load(f)


def test_kat_TestSocketLine():
    # aggregator/sock_line_test.go:11-349, replayed in full: every timestamp of the reference's tsList
    # (tests/golden/sockline_kat.json, extracted by tests/golden/make_sockline_fixture.py) is added as an
    # open with the same (empty) SockInfo; AddValue's dedupe collapses them to the first
    # (sock_num_line.go:70-78); GetValue(33835107729129) must succeed (sock_line_test.go:342-347).
    kat = _sockline_kat()
    assert len(kat["ts_list"]) >= 300 and kat["query"] == 33835107729129
    ln = _Line()
    for ts in kat["ts_list"]:
        ln.add(ts, _SI())
    assert len(ln) == 1
    assert ln.get(kat["query"]) is not None
    # the same inserts with DISTINCT sockets keep every entry, sorted, and the query lands on the last open
    # socket at or before it (sock_num_line.go:90-92, :121)
    ln2 = _Line()
    for i, ts in enumerate(kat["ts_list"]):
        ln2.add(ts, _SI(saddr=i + 1))
    assert len(ln2) == len(set(kat["ts_list"]))
    before = max(i for i, ts in enumerate(kat["ts_list"]) if ts < kat["query"])
    assert ln2.get(kat["query"]).saddr == before + 1


def test_kat_TestXxx_open_close_pairs():
    # sock_line_test.go:351-441: opens at 10/30/50, closes at 20/40/60
    ln = _Line()
    for ts, si in [(10, _SI()), (20, None), (30, _SI()), (40, None), (50, _SI()), (60, None)]:
        ln.add(ts, si)
    assert ln.get(52) is not None     # inside [50,60): "should return 50"
    assert ln.get(33) is not None     # inside [30,40)
    assert ln.get(45) is not None     # closed gap, same daddr/dport both sides -> closest
    assert ln.get(5) is not None      # before first entry, first is open (:107-115)


def test_kat_TestXxx2_later_socket_wins():
    # sock_line_test.go:443-473
    ln = _Line()
    s1, s2 = _SI(saddr=0x7878), _SI(saddr=0x7979)   # "xx", "yy"
    ln.add(0, s2)
    ln.add(247453008321477, s1)
    got = ln.get(247453008321499)
    assert got is not None and got.saddr == 0x7878


def test_kat_TestAlreadyEstablishCanBeFound():
    # sock_line_test.go:475-501: index-0 branch returns the first open socket
    ln = _Line()
    ln.add(0, _SI(saddr=0x7979))
    got = ln.get(0)
    assert got is not None and got.saddr == 0x7979


def test_delete_unused_hand_derived():
    # sock_num_line.go:160-209, by hand. M = one minute of ns.
    M = 60 * 10**9
    # (a) <= 1 value: untouched (:165-167)
    ln = _Line(); ln.add(10, _SI(saddr=1)); ln.delete_unused(); assert len(ln) == 1
    # (b) [open, close]: the first loop runs while i < len-1, appends the open and stops: the close is gone.
    ln = _Line(); ln.add(10, _SI(saddr=1)); ln.add(20, None); ln.delete_unused()
    assert len(ln) == 1 and ln.get(15).saddr == 1 and ln.get(10**6).saddr == 1   # last value is the open now
    # (c) two opens in a row: the first is dropped (its close never arrived, :170-176); here they are the
    # last two values, the loop steps over both, so nothing else is lost
    ln = _Line(); ln.add(10, _SI(saddr=1)); ln.add(20, _SI(saddr=2)); ln.delete_unused()
    assert len(ln) == 1 and ln.get(15).saddr == 2
    # (d) closed pairs whose open was last matched more than five minutes before the line's latest match go
    # (:193-208): opens at 10/30/50 closes at 20/40/60, plus an open at 70 that keeps 60 in the first loop
    ln = _Line()
    for ts, si in [(10, _SI(saddr=1)), (20, None), (30, _SI(saddr=3)), (40, None), (50, _SI(saddr=5)), (60, None),
                   (70, _SI(saddr=7))]:
        ln.add(ts, si)
    assert ln.get_at(15, 1 * M).saddr == 1          # pair (10,20) last matched at 1 min
    assert ln.get_at(35, 3 * M).saddr == 3          # pair (30,40) at 3 min
    assert ln.get_at(55, 7 * M).saddr == 5          # pair (50,60) at 7 min = the latest match
    ln.delete_unused()
    # first loop: 7 values, no two opens adjacent -> values 0..5 kept, the open at 70 dropped.
    # second loop from the back: (50,60): 7M+5M < 7M no. (30,40): 3M+5M < 7M no. (10,20): 1M+5M < 7M yes -> gone
    assert len(ln) == 4
    assert ln.get(15).saddr == 3                    # before the first value, which is the open at 30 (:107-115)
    assert ln.get(65).saddr == 5                    # after the close at 60, within a minute of the open at 50 (:97-101)


def test_sockline_closed_last_entry_rules():
    # sock_num_line.go:94-105
    ln = _Line()
    ln.add(100, _SI(saddr=1, daddr=2, dport=80))
    ln.add(200, None)
    assert ln.get(250).saddr == 1                      # within one minute of the open
    assert ln.get(100 + 60 * 10**9 + 1) is None        # too late: "closed socket on last entry"
    ln2 = _Line()
    ln2.add(100, None)
    assert ln2.get(50) is None                         # :117-118 first entry is a close


def test_committed_synthetic_fixture_still_matches_the_oracle():
    # tests/golden/synth_small.npz (made by tests/golden/make_synth_golden.py)
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "synth_small.npz"))
    o = ol.Oracle()
    o.load_tables(z["pod_ip"], z["svc_ip"])
    o.process(z["events"].view(abi.L7_REC))
    assert edges_equal(o.edges(), z["edges"].view(abi.EDGE_OUT))
    a = _py_oracle({int(ip): k for k, ip in enumerate(z["pod_ip"])}, {int(ip): k for k, ip in enumerate(z["svc_ip"])})
    for r in z["events"].view(abi.L7_REC):
        a.process_l7(r)
    assert edges_equal(pyref_edges(a), z["edges"].view(abi.EDGE_OUT))
    st = o.stats()
    assert [st["events_in"], st["rows_emitted"], st["not_request"], st["src_unresolved"]] == z["stats"].tolist()
