#!/bin/bash
# round 2: hot-tier reductions issued by all lanes (selected operands) instead of divergent regions
set -u
mkdir -p gpurun_out
echo "== gpu parity"; timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_property.py tests/test_gpu_boundary.py tests/test_gpu_windows.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_k3.txt
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-gnn"
for s in 0 4 0; do
  echo "== shape $s"; ALZ_INGEST_SHAPE=$s timeout 400 $B 2>&1 | tail -1 | tee gpurun_out/bench_k3_shape$s.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d['verify'])"
done
echo "== ncu warm v8"
timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k regex:ingest_pairs_v8 -s 4 -c 1 -o gpurun_out/prof_r2k3_v8_warm -f \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-gnn --no-verify > gpurun_out/ncu_full_k3.log 2>&1
tail -1 gpurun_out/ncu_full_k3.log | cut -c1-200
