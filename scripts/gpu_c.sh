#!/bin/bash
# round 2, call 3: where the ingest kernel's time goes (instrumented build, warm caches), dictionary size,
# an ncu capture that keeps the caches warm between replay passes
set -u
mkdir -p gpurun_out
P="python scripts/ingest_prof.py"
export ALZ_LIB_PATH=alaz_b200/lib/libalazgpu_prof.so
echo "== prof default";  timeout 300 $P | tee gpurun_out/ingest_prof_default.json
echo "== prof max_pairs 2^18"; timeout 300 $P --max-pairs 262144 | tee gpurun_out/ingest_prof_p18.json
echo "== prof shape 2 (12 warps)"; ALZ_INGEST_SHAPE=2 timeout 300 $P | tee gpurun_out/ingest_prof_s2.json
echo "== prof shape 3 (20 warps)"; ALZ_INGEST_SHAPE=3 timeout 300 $P | tee gpurun_out/ingest_prof_s3.json
unset ALZ_LIB_PATH
echo "== ncu warm"
timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k regex:ingest_pairs_v6 -s 4 -c 1 -o gpurun_out/prof_r2c_ingest_warm -f \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-gnn --no-verify > gpurun_out/ncu_full_c.log 2>&1
tail -2 gpurun_out/ncu_full_c.log | cut -c1-300
