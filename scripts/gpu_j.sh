#!/bin/bash
# round 2: sectored pair rows (cell + latency of an event in one 32-byte sector, one red.u64 per lane pair)
set -u
mkdir -p gpurun_out
echo "== gpu tests (all)"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_j.txt
for s in 10; do
echo "== parity under shape $s"; ALZ_INGEST_SHAPE=$s timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_property.py tests/test_gpu_boundary.py tests/test_gpu_windows.py -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_j_shape$s.txt
done
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-gnn"
for s in 0 1 9 10; do
  echo "== shape $s"; ALZ_INGEST_SHAPE=$s timeout 400 $B 2>&1 | tail -1 | tee gpurun_out/bench_j_shape$s.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d['verify'])"
done
echo "== ncu warm v9-24"
ALZ_INGEST_SHAPE=10 timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k regex:ingest_pairs_v9 -s 4 -c 1 -o gpurun_out/prof_r2j_v9_24_warm -f \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-gnn --no-verify > gpurun_out/ncu_full_j.log 2>&1
tail -1 gpurun_out/ncu_full_j.log | cut -c1-200
