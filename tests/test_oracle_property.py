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
