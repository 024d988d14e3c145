#!/bin/bash
# round 2: GNN TMA-gather layer + sync-free run; ingest L2 evict-last hint A/B
set -u
mkdir -p gpurun_out
echo "== gpu tests (all)"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_g.txt
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e"
for k in 1 0; do for s in 0 1; do
  echo "== keep $k shape $s"; ALZ_INGEST_KEEP=$k ALZ_INGEST_SHAPE=$s timeout 400 $B 2>&1 | tail -1 | tee gpurun_out/bench_g_keep${k}_shape$s.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['verify'])"
done; done
echo "== ncu warm"
timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k regex:ingest_pairs_v8 -s 4 -c 1 -o gpurun_out/prof_r2g_ingest_warm -f \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-gnn --no-verify > gpurun_out/ncu_full_g.log 2>&1
tail -1 gpurun_out/ncu_full_g.log | cut -c1-200
echo "== ncu gnn"
timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k regex:sage_layer -c 2 -o gpurun_out/prof_r2g_gnn -f \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-verify > gpurun_out/ncu_gnn_g.log 2>&1
tail -1 gpurun_out/ncu_gnn_g.log | cut -c1-200
