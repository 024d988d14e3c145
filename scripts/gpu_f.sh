#!/bin/bash
# round 2: ingest v8 (TMA ring + queued cold tier + windowed 16-bit rows + pod filter): parity, sections, bench
set -u
mkdir -p gpurun_out
echo "== gpu tests (all)"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_f.txt
export ALZ_LIB_PATH=alaz_b200/lib/libalazgpu_prof.so
for s in 0 1 2 3; do echo "== sections shape $s"; ALZ_INGEST_SHAPE=$s timeout 300 python scripts/ingest_prof.py | tee gpurun_out/ingest_prof_v8_s$s.json; done
unset ALZ_LIB_PATH
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-gnn --no-e2e"
for s in 0 1 2 3; do
  echo "== v8 shape $s"; ALZ_INGEST_SHAPE=$s timeout 400 $B 2>&1 | tail -1 | tee gpurun_out/bench_f_shape$s.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['verify'])"
done
echo "== ncu warm"
timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k regex:ingest_pairs_v8 -s 4 -c 1 -o gpurun_out/prof_r2f_ingest_warm -f \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-gnn --no-verify > gpurun_out/ncu_full_f.log 2>&1
tail -1 gpurun_out/ncu_full_f.log | cut -c1-200
