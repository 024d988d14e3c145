#!/usr/bin/env python
"""Regenerates tests/golden/synth_small.npz: a seeded synthetic stream (all protocol branches), its tables
and the per-edge result of the CPU oracle (oracle/alz_oracle.c, cross-checked by oracle/ref_py.py in
tests/test_oracle.py). The committed file lets a GPU box check the CUDA path against fixed vectors
even if the oracle library were to change. Run from the repo root: python tests/golden/make_synth_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
from alaz_b200 import abi  # noqa: E402
from helpers import sort_edges  # noqa: E402

S, N, SEED = 40, 20_000, 0xA1A2F1F1
t = ol.Topo(S, seed=SEED, mix=abi.MIX_ALL)
ev = t.events(0, N)
o = ol.Oracle()
o.load_tables(t.pod_ip, t.svc_ip)
o.process(ev)
edges = sort_edges(o.edges())
st = o.stats()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "synth_small.npz"), events=ev, pod_ip=t.pod_ip,
                    svc_ip=t.svc_ip, edges=edges,
                    stats=np.array([st["events_in"], st["rows_emitted"], st["not_request"], st["src_unresolved"]],
                                   dtype=np.uint64))
print(len(ev), "events ->", len(edges), "edges", st)
