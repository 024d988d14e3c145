#!/bin/bash
# round 2: tier-S replicated rows: parity + bench on one GPU
set -u
mkdir -p gpurun_out
echo "== gpu parity"; timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_property.py tests/test_gpu_boundary.py tests/test_gpu_windows.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_n.txt
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-gnn"
for s in 0; do
  echo "== shape $s"; ALZ_INGEST_SHAPE=$s timeout 400 $B 2>&1 | tail -1 | tee gpurun_out/bench_n_shape$s.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d['verify'])"
done
