#!/bin/bash
# round 2, call 2: parity after the stage-release fix, A/B of the cold tier rewrite, new bench.py, ncu capture
set -u
mkdir -p gpurun_out
echo "== boundary+parity tests"; timeout 1500 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_parity.py tests/test_gpu_windows.py tests/test_gpu_property.py -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_b.txt
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-gnn --no-e2e"
for s in 0 1 2 3; do
  echo "== v6.1 shape $s"; ALZ_INGEST_SHAPE=$s timeout 400 $B 2>&1 | tail -1 | tee gpurun_out/bench_b_shape$s.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['verify'])"
done
echo "== full default bench"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_b_full.json | cut -c1-3000
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ingest_pairs_v6 -s 3 -c 1 -o gpurun_out/prof_r2b_ingest -f \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-gnn --no-verify > gpurun_out/ncu_full_b.log 2>&1
tail -3 gpurun_out/ncu_full_b.log
