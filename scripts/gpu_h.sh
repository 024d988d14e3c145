#!/bin/bash
# round 2: device-incremental socket timelines (join / gc / alive); ingest probe-wait diagnosis; v9 (32 warps)
set -u
mkdir -p gpurun_out
echo "== gpu tests (all)"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_h.txt
for s in 9 10; do
echo "== parity under shape $s"; ALZ_INGEST_SHAPE=$s timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_property.py tests/test_gpu_boundary.py -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_h_shape$s.txt
done
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-gnn"
for s in 0 9 10; do
  echo "== shape $s"; ALZ_INGEST_SHAPE=$s timeout 400 $B 2>&1 | tail -1 | tee gpurun_out/bench_h_shape$s.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d['verify'])"
done
for d in 1 2 3; do
  echo "== diag $d"; ALZ_INGEST_DIAG=$d timeout 400 $B --no-verify 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'])"
done
echo "== ncu warm v9"
ALZ_INGEST_SHAPE=9 timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k regex:ingest_pairs_v9 -s 4 -c 1 -o gpurun_out/prof_r2h_v9_warm -f \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-gnn --no-verify > gpurun_out/ncu_full_h.log 2>&1
tail -1 gpurun_out/ncu_full_h.log | cut -c1-200
