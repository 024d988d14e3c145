#!/bin/bash
# One gpurun call of round 2: parity tests, A/B of the ingest kernel (r1 build vs this build, CTA shapes),
# launch list and one full ncu capture of the ingest kernel. Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-gnn"
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== r1 build"; ALZ_LIB_PATH=alaz_b200/lib/libalazgpu_r1.so ALZ_ABI_VERSION=1 timeout 300 $B --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_r1lib.json | cut -c1-600
for s in 0 1 2 3; do
  echo "== v6 shape $s"; ALZ_INGEST_SHAPE=$s timeout 300 $B --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_v6_shape$s.json | cut -c1-600
done
echo "== v6 e2e"; timeout 400 $B 2>&1 | tail -1 | tee gpurun_out/bench_v6_e2e.json | cut -c1-900
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2a.csv \
  python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ingest_pairs_v6 -s 3 -c 1 -o gpurun_out/prof_r2a_ingest -f \
  python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e --no-gnn > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -20
