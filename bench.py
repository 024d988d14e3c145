#!/usr/bin/env python
"""bench.py — l7_req events/sec aggregated into the per-edge service graph.

One "step" = one pass of the hot path over one batch of the synthetic stream:
alz_submit_l7_device (ingest kernel) + alz_window_flush_device (join of the
distinct socket pairs, canonical edge list; at N>1 the cross-rank merge with
its single all-reduce on the per-edge accumulators). Workload at N=1 is
BASELINE.json configs[1]: 10k services / 100M l7_req events on one B200;
at N>1 every rank gets its own 100M-event shard of the global stream
(events owned by alz_owner_rank(saddr)), i.e. weak scaling.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`--impl reference` times the CPU restatement of the reference aggregator
(oracle/alz_oracle.c — the Go binary cannot be built here) on
the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "l7_req events/sec aggregated (per-edge count/5xx/latency histogram)"
UNIT = "events/s"
ALG_BYTES_PER_EVENT = 32      # each compact record read once
ALG_BYTES_PER_EDGE = 296      # each live edge row written once per window (alz_edge_out)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--services", type=int, default=10_000)
    ap.add_argument("--events", type=int, default=100_000_000, help="events per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=None,
                    help="events per step timed on the CPU arm (default: 32M, or less so that --impl reference ends in ~1-2 min)")
    ap.add_argument("--eager", action="store_true", help="ALZ_CFG_EAGER_JOIN plan")
    ap.add_argument("--no-smem-cache", action="store_true", help="ingest v1: global reductions only")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gnn", action="store_true", help="skip the GNN-update timing (extra key gnn_update)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.p, self.index = [], None, index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.rows.append([time.perf_counter()] + [x.strip() for x in ln.strip().split(",")])

    def mark(self):
        """Samples from here on are 'under load' (the timed regions); earlier ones are kept as a fallback."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_mark = getattr(self, "t_mark", 0.0)
        loaded = [r for r in self.rows if r[0] >= t_mark]
        rows, scope = (loaded, "timed regions (device-resident steps + e2e steps)") if loaded else \
                      (self.rows, "whole run incl. warm-up (timed regions shorter than the sampling period)")
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm), "scope": scope}


def cpu_arm(services, n_events, seed, nthreads, steps=1, warmup=0):
    """Time the CPU restatement (oracle/alz_oracle.c) on a bounded sample of the workload."""
    import oracle_lib as ol
    t = ol.Topo(services, seed=seed)
    ev = t.events(0, n_events)
    times = []
    for it in range(warmup + steps):
        o = ol.Oracle()
        o.load_tables(t.pod_ip, t.svc_ip)
        t0 = time.perf_counter()
        o.process(ev, nthreads)
        edges = o.edges()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        o.close()
    return n_events / (sum(times) / len(times)), len(edges), sum(times) / len(times)


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if args.cpu_sample is None:   # bounded: ~400M events in total over all steps, 2M..32M per step
        args.cpu_sample = max(2_000_000, min(32_000_000, 400_000_000 // max(1, args.steps + args.warmup)))
    v, n_edges, sec = cpu_arm(args.services, args.cpu_sample, 0xA1A20001, cores, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64/u32 integer", "data": "synthetic",
        "config": {"workload": f"{args.services} services / {args.events} l7_req events per GPU per step "
                               "(BASELINE.json configs[1])",
                   "note": "CPU restatement of aggregator/data.go resolve/emit + group-by (the Go binary "
                           "cannot be built here: no Go toolchain)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"first {args.cpu_sample} events of the workload stream per step, "
                                   f"{cores} threads, tables preloaded"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return 0
    if args.cpu_sample is None:
        args.cpu_sample = 32_000_000

    import torch
    import torch.distributed as dist
    from alaz_b200 import abi, capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference "
                         "for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    S, N = args.services, args.events
    seed = 0xA1A20001
    topo = capi.Topo(S, seed=seed)
    flags = (abi.CFG_EAGER_JOIN if args.eager else 0) | (abi.CFG_NO_SMEM_CACHE if args.no_smem_cache else 0)
    h = capi.Handle(device=local_rank, max_endpoints=4 * S, max_pairs=max(1 << 20, 40 * S),
                    max_batch=1 << 22, flags=flags)
    stream = torch.cuda.Stream()          # a real stream: handle 0 would mean "library's own"
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)
    h.load_tables(topo.pod_ip, topo.svc_ip)

    # ---- inputs resident in HBM before the timed region
    d_ev = h.dev_alloc(N * 32)
    if world > 1:
        comm_setup(h, dist, rank, world, torch)
        n_scanned = fill_owned(h, topo, d_ev, N, world, rank)
    else:
        topo.fill_device(h, 0, N, d_ev)
        n_scanned = N
    h.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        h.submit_device(d_ev, N)
        return h.flush_device()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        _, n_edges = step()
    barrier()
    sampler.mark()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_all0.record(stream)
    for k in range(args.steps):
        ev[k][0].record(stream)
        h.submit_device(d_ev, N)
        ev[k][1].record(stream)
        _, n_edges = h.flush_device()
        ev[k][2].record(stream)
    t_all1.record(stream)
    barrier()
    total_ms = t_all0.elapsed_time(t_all1)
    ingest_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    flush_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    st = h.stats()

    tt = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    ti = torch.tensor([ingest_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ti, op=dist.ReduceOp.MAX)
    total_ms, ingest_ms_max = float(tt.item()), float(ti.item())
    ms_per_step = total_ms / args.steps
    value = world * N / (ms_per_step * 1e-3)

    # ---- end to end through the C ABI with host buffers (H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(h, capi, abi, d_ev, N, n_edges, args, world, dist if world > 1 else None, torch)

    clocks = sampler.stop() if rank == 0 else None
    gnn_ms = None
    if not args.no_gnn:
        gnn_ms = run_gnn(h, step, args, torch)

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = N * ALG_BYTES_PER_EVENT
        achieved = alg_bytes / (ingest_ms_max * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64/u32 integer", "data": "synthetic",
            "config": {
                "workload": f"{S} services / {N} l7_req events per GPU per step (BASELINE.json configs[1]), "
                            f"Zipf(1.1) over {topo.n_edges} edges, 32-B compact records",
                "plan": "eager-join" if args.eager else "reduce-per-socket-pair then join distinct pairs",
                "l2": "inputs (3.2 GB/step) larger than L2; no explicit flush",
                "parallelism": f"dp{world} by alz_owner_rank(saddr)" if world > 1 else "single GPU",
                "live_edges": int(n_edges), "rows_emitted_per_step": int(st["rows_emitted"] // max(1, st["events_in"] // N)),
            },
            "roofline": {"bound": "hbm", "kernel": "ingest_eager_kernel" if args.eager else ("ingest_pairs_kernel(v1)" if args.no_smem_cache else "ingest_pairs_v4_kernel"),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full on this
                         # workload (profiles/r1_v5_ingest_ncu.txt); only valid for the default configuration
                         "traffic": (3.453e9 if (S == 10_000 and N == 100_000_000 and world == 1 and not args.eager
                                                and not args.no_smem_cache) else None),
                         "kernel_ms": ingest_ms_max, "flush_ms": flush_ms,
                         "algorithmic_bytes_per_launch": alg_bytes},
            "clocks": clocks,
            # own kernels per step: ingest, 2x fold_resolve, 2x fold_pairs, 2x hot_pick, 2x hot_emit, iota, gather
            # (CUB sort/scan kernels and memsets not counted)
            "gpu_launches": args.steps * 11,
            "e2e": e2e,
        }
        if gnn_ms is not None:
            line["gnn_update"] = gnn_ms
        if not args.no_cpu:
            cores = os.cpu_count() or 1
            v, _, sec = cpu_arm(S, args.cpu_sample, seed, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {args.cpu_sample} events of the same stream, {cores} threads "
                                              f"({sec:.1f} s); restatement of aggregator/data.go, not the Go binary"}
        print(json.dumps(line), flush=True)
    h.dev_free(d_ev)
    topo.close()
    h.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_e2e(h, capi, abi, d_ev, N, n_edges, args, world, dist, torch):
    """Same metric through alz_submit_l7 / alz_window_flush with HOST buffers."""
    n_e2e = N
    pin = capi.PinnedBuffer(n_e2e, abi.L7_REC)
    got = h.d2h(d_ev, n_e2e, abi.L7_REC)     # host copy of the very same stream
    pin.array[:] = got
    del got
    out = capi.PinnedBuffer(h.max_edges, abi.EDGE_OUT)
    import ctypes as C
    n_out = C.c_size_t(0)
    steps = max(2, min(args.steps, 5))

    def one():
        h.submit_ptr(pin.ptr, n_e2e)
        h._ck(h.L.alz_window_flush(h.h, C.c_void_p(out.ptr), h.max_edges, C.byref(n_out)), "alz_window_flush")

    one()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    res = {"value": world * n_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": n_e2e * 32,
           "d2h_bytes_per_step": int(n_out.value) * abi.EDGE_OUT.itemsize, "ms_per_step": dt * 1e3,
           "steps": steps, "api": "alz_submit_l7 (pinned host records) + alz_window_flush (host edge rows)"}
    pin.free()
    out.free()
    return res


def run_gnn(h, step, args, torch):
    """GNN-update ms = CSR build + 2 GraphSAGE layers + edge scoring over the flushed window (device time)."""
    import ctypes as C
    step()
    p, n_out = C.c_void_p(), C.c_size_t(0)
    rc = h.L.alz_gnn_score_device(h.h, C.byref(p), C.byref(n_out))
    if rc != 0:
        return None
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    reps = 5
    e0.record(stream)
    for _ in range(reps):
        h.L.alz_gnn_score_device(h.h, C.byref(p), C.byref(n_out))
    e1.record(stream)
    torch.cuda.synchronize()
    return {"ms": e0.elapsed_time(e1) / reps, "edges": int(n_out.value), "layers": 2, "d": 64,
            "what": "node set + CSR build + 2x GraphSAGE-mean + edge scores, replicated per rank"}


def comm_setup(h, dist, rank, world, torch):
    import ctypes as C
    from alaz_b200 import abi
    idbuf = (C.c_uint8 * abi.COMM_ID_BYTES)()
    if rank == 0:
        h._ck(h.L.alz_comm_unique_id(idbuf), "alz_comm_unique_id")
    t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().tolist())
    idbuf = (C.c_uint8 * abi.COMM_ID_BYTES).from_buffer_copy(raw)
    h._ck(h.L.alz_comm_init(h.h, world, rank, idbuf), "alz_comm_init")


def fill_owned(h, topo, d_ev, N, world, rank):
    """Fill d_ev with the first N events of the global stream owned by this rank."""
    import ctypes as C
    if topo.dev is None:
        topo.to_device(h)
    n_written, n_scanned = C.c_uint64(0), C.c_uint64(0)
    h._ck(h.L.alz_synth_dev_fill_owned(h.h, topo.dev, 0, world, rank, C.c_void_p(d_ev), N,
                                       C.byref(n_written), C.byref(n_scanned)), "alz_synth_dev_fill_owned")
    assert n_written.value == N
    return n_scanned.value


if __name__ == "__main__":
    sys.exit(main())
