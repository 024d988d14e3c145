"""Property test (hypothesis): the C oracle and the independent Python restatement agree on arbitrary small
tables and event lists — including IPs that are pod AND service, unknown sources, every protocol value,
method values that alias across protocols, extreme durations and statuses."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle_lib as ol
from alaz_b200 import abi
from helpers import edges_equal, explain_diff, pyref_edges
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import ref_py  # noqa: E402

IPS = [abi.ip("10.0.0.%d" % i) for i in range(1, 9)] + [0, 0xFFFFFFFF, abi.ip("8.8.8.8"), abi.ip("127.0.0.1")]
ip_s = st.sampled_from(IPS)
event_s = st.tuples(ip_s, ip_s, st.integers(0, 12), st.integers(0, 255),
                    st.sampled_from([0, 1, 2, 200, 404, 499, 500, 503, 599, 600, 65535]),
                    st.sampled_from([0, 1, 255, 256, 383, 384, 10**6, (1 << 40) - 1, 1 << 40, (1 << 64) - 1]))
table_s = st.dictionaries(ip_s, st.integers(0, 50), max_size=8)


@settings(max_examples=150, deadline=None)
@given(pods=table_s, svcs=table_s, events=st.lists(event_s, max_size=60))
def test_c_oracle_equals_python_restatement(pods, svcs, events):
    o = ol.Oracle()
    a = ref_py.Aggregator()
    for ip, i in pods.items():
        o.upsert(abi.TABLE_POD, ip, i)
        a.pod_ip_to_uid[ref_py.ip_string(ip)] = "pod-%d" % i
    for ip, i in svcs.items():
        o.upsert(abi.TABLE_SVC, ip, i)
        a.svc_ip_to_uid[ref_py.ip_string(ip)] = "svc-%d" % i
    recs = np.zeros(len(events), dtype=abi.L7_REC)
    for k, (s, d, proto, mf, status, dur) in enumerate(events):
        recs[k] = (s, d, 1000 + k, 80, status, proto, mf, dur, k)
    o.process(recs)
    for r in recs:
        a.process_l7(r)
    got, exp = o.edges(), pyref_edges(a)
    assert edges_equal(got, exp), explain_diff(got, exp)
    stc = o.stats()
    for key, v in a.stats.items():
        assert stc[key] == v


# ---- SocketLine: C restatement vs the independent Python one, arbitrary interleavings -------------------------
_ts_s = st.integers(0, 40) .map(lambda k: k * 20 * 10**9)          # 20 s steps: the one-minute rule is exercised
_op_s = st.one_of(
    st.tuples(st.just("tcp"), st.sampled_from([1, 5, 1, 5, 2]), st.integers(1, 3), st.integers(3, 5), _ts_s,
              st.sampled_from([abi.ip("10.0.0.1"), abi.ip("10.0.0.2"), abi.ip("127.0.0.1")]),
              st.sampled_from([abi.ip("10.0.0.5"), abi.ip("10.0.0.6"), abi.ip("8.8.8.8")]),
              st.sampled_from([1000, 1001]), st.sampled_from([80, 443])),
    st.tuples(st.just("get"), st.integers(1, 3), st.integers(3, 5), st.integers(0, 45).map(lambda k: k * 17 * 10**9 + 3),
              st.integers(1, 20).map(lambda m: m * 60 * 10**9)),
    st.tuples(st.just("gc")),
)


@settings(max_examples=300, deadline=None)
@given(ops=st.lists(_op_s, max_size=60))
def test_sockline_c_equals_python_restatement(ops):
    c = ol.SockMaps()
    p = ref_py.SocketMaps()
    for op in ops:
        if op[0] == "tcp":
            _, typ, pid, fd, ts, sa, da, sp, dp = op
            r = np.zeros(1, dtype=abi.TCP_REC)
            r[0] = (fd, ts, pid, sa, da, sp, dp, typ, 0)
            c.process(r)
            p.process_tcp(typ, pid, fd, ts, sa, da, sp, dp)
        elif op[0] == "get":
            _, pid, fd, ts, now = op
            q = np.zeros(1, dtype=abi.SOCK_QUERY)
            q[0] = (fd, ts, pid, 0)
            got = c.lookup(q, now_ns=now)[0]
            exp = p.lookup(pid, fd, ts, now)
            if exp is None:
                assert got["found"] == 0, (op, got)
            else:
                assert got["found"] == 1 and (int(got["saddr"]), int(got["daddr"]), int(got["sport"]), int(got["dport"])) == exp, (op, got, exp)
        else:
            c.gc()
            p.gc()
        assert c.records() == sum(len(l.values) for l in p.lines.values())
    assert c.localhost_dropped == p.localhost_dropped
    # sendOpenConnection over what is left
    o = ol.Oracle()
    pods = {abi.ip("10.0.0.1"): 7, abi.ip("10.0.0.5"): 9}
    svcs = {abi.ip("10.0.0.6"): 3}
    for ip, i in pods.items():
        o.upsert(abi.TABLE_POD, ip, i)
    for ip, i in svcs.items():
        o.upsert(abi.TABLE_SVC, ip, i)
    got = c.alive(o, cap=256)
    exp = p.alive({ref_py.ip_string(ip): i for ip, i in pods.items()}, {ref_py.ip_string(ip): i for ip, i in svcs.items()})
    tname = {abi.NODE_POD: "pod", abi.NODE_SVC: "service", abi.NODE_OUTBOUND: "outbound"}
    got_rows = sorted((int(g["from_ip"]), int(g["from_id"]), int(g["from_port"]), int(g["to_ip"]), tname[int(g["to_type"])],
                       int(g["to_id"]) if int(g["to_type"]) != abi.NODE_OUTBOUND else ref_py.ip_string(int(g["to_id"])),
                       int(g["to_port"])) for g in got)
    assert got_rows == sorted(exp)


# ---- Host header: the C parser vs the Python one on arbitrary payloads; host-keyed rows through both aggregators --
_frag = st.sampled_from(["GET / HTTP/1.1", "Host:", "Host: ", "Host:  ", "host: x", "Host: a.com", "Host: b.org:80", " ", "\r",
                         "X-Host: no", "Host:nospace", "8.8.8.8", "Host: 8.8.8.8", "Host: a.com extra", "", "Accept: */*"])
_payload_s = st.lists(_frag, max_size=8).flatmap(
    lambda fr: st.lists(st.sampled_from(["\n", "\r\n", " ", ""]), min_size=len(fr), max_size=len(fr)).map(
        lambda seps: "".join(f + s_ for f, s_ in zip(fr, seps))))


@settings(max_examples=400, deadline=None)
@given(payload=_payload_s)
def test_host_header_parsers_agree(payload):
    assert ol.parse_http_host(payload.encode("latin-1")) == ref_py.parse_http_payload_host(payload)


@settings(max_examples=100, deadline=None)
@given(pods=table_s, svcs=table_s,
       events=st.lists(st.tuples(ip_s, ip_s, st.sampled_from([1, 1, 1, 2, 5, 3]), st.integers(0, 3), st.integers(0, 2)), max_size=40))
def test_host_keyed_rows_c_equals_python(pods, svcs, events):
    """orc_process_l7_hosts against the Python aggregator fed the payloads themselves: only HTTP events carry a payload
    (processHttpEvent is the only handler that parses one), the header keys outbound destinations only."""
    names = ["api.example.com", "b.org:80", "8.8.8.8"]
    o = ol.Oracle()
    a = ref_py.Aggregator()
    for ip, i in pods.items():
        o.upsert(abi.TABLE_POD, ip, i)
        a.pod_ip_to_uid[ref_py.ip_string(ip)] = "pod-%d" % i
    for ip, i in svcs.items():
        o.upsert(abi.TABLE_SVC, ip, i)
        a.svc_ip_to_uid[ref_py.ip_string(ip)] = "svc-%d" % i
    recs = np.zeros(len(events), dtype=abi.L7_REC)
    host_idx = np.zeros(len(events), dtype=np.uint32)
    for k, (s_, d, proto, meth, hn) in enumerate(events):
        recs[k] = (s_, d, 1000 + k, 80, 200, proto, meth, 1000 + k, k)
        has = proto == 1 and hn > 0
        host_idx[k] = hn if has else 0
        payload = ("GET / HTTP/1.1\r\nHost: %s\r\n\r\n" % names[hn - 1]) if has else ("GET / HTTP/1.1\r\n\r\n" if proto == 1 else None)
        a.process_l7(recs[k], payload)
    o.process_hosts(recs, host_idx, names)
    got = o.edges()
    # the C side reports a host-keyed node as (ALZ_NODE_OUTBOUND_HOST, index into names), or as a raw address when the
    # text is a dotted quad; translate the Python groups the same way
    exp = {}
    for (ft, fu, tt, tu), g in a.groups.items():
        exp[(ft, fu, tt, tu)] = g["count"]
    got_map = {}
    tname = {abi.NODE_POD: "pod", abi.NODE_SVC: "service", abi.NODE_OUTBOUND: "outbound", abi.NODE_OUTBOUND_HOST: "outbound"}

    def uid(t, v):
        if t == abi.NODE_POD:
            return "pod-%d" % v
        if t == abi.NODE_SVC:
            return "svc-%d" % v
        if t == abi.NODE_OUTBOUND_HOST:
            return names[v]
        return ref_py.ip_string(v)
    for e in got:
        k = (tname[int(e["from_type"])], uid(int(e["from_type"]), int(e["from"])), tname[int(e["to_type"])], uid(int(e["to_type"]), int(e["to"])))
        got_map[k] = got_map.get(k, 0) + int(e["count"])
    assert got_map == exp
