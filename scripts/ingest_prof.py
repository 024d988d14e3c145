#!/usr/bin/env python
"""Where does a warp of the ingest kernel spend its cycles in a REAL run (warm caches, no profiler)?
Runs the profiling build (alaz_b200/build.py --prof -> libalazgpu_prof.so) on config 2 and prints the
per-section cycle shares. Usage (GPU box):
  ALZ_LIB_PATH=alaz_b200/lib/libalazgpu_prof.so python scripts/ingest_prof.py [--max-pairs N] [--steps K]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alaz_b200 import abi, capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--services", type=int, default=10_000)
ap.add_argument("--events", type=int, default=100_000_000)
ap.add_argument("--max-pairs", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--windows", type=int, default=3)
a = ap.parse_args()

import torch  # noqa: E402
L = capi.load()
L.alz_debug_ingest_prof.argtypes = [C.c_void_p]
topo = capi.Topo(a.services, seed=0xA1A20001)
h = capi.Handle(max_endpoints=4 * a.services, max_pairs=a.max_pairs)
h.load_tables(topo.pod_ip, topo.svc_ip)
wins = []
for k in range(a.windows):
    d = h.dev_alloc(a.events * 32)
    topo.fill_device(h, k * a.events, a.events, d)
    wins.append(d)
h.sync()
buf = (C.c_ulonglong * 12)()
for k in range(3):
    h.submit_device(wins[k % a.windows], a.events); h.flush_device()
h.sync()
rc = L.alz_debug_ingest_prof(buf)
if rc != 0:
    raise SystemExit("not a profiling build: set ALZ_LIB_PATH=alaz_b200/lib/libalazgpu_prof.so")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ms = 0.0
for k in range(a.steps):
    torch.cuda.synchronize()
    e0.record()
    h.submit_device(wins[(3 + k) % a.windows], a.events)
    e1.record()
    h.flush_device()
    torch.cuda.synchronize()
    ms += e0.elapsed_time(e1)
L.alz_debug_ingest_prof(buf)
v = [int(x) for x in buf]
tot = max(1, v[0])
names = ["loop total", "wait for the TMA stage", "wait for the probes (cp.async)", "cold tier incl. its waits + slow",
         "slow batches", "hot tier (record loads .. queue pushes)", "iterations", "cold batches"]
out = {"ingest_ms_with_instrumentation": ms / a.steps, "max_pairs": a.max_pairs,
       "cycles_per_iteration": v[0] / max(1, v[6]),
       "shares": {names[i]: round(v[i] / tot, 4) for i in (1, 2, 3, 4, 5)},
       "cold_batches_per_iteration": v[7] / max(1, v[6])}
ctas = max(1, v[11])
warps_per_cta = 16
out["phases_cycles_per_cta"] = {"prologue (smem init, hot list preload)": v[8] / ctas,
                                "main loop until the slowest warp is done": v[9] / ctas,
                                "mean warp's main loop (16 warps assumed)": v[0] / ctas / warps_per_cta,
                                "drain of the private rows": v[10] / ctas}
out["phases_ms_at_1965MHz"] = {k: round(c / 1.965e6, 4) for k, c in out["phases_cycles_per_cta"].items()}
print(json.dumps(out))
