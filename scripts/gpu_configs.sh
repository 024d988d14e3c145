#!/bin/bash
# BASELINE configs[2] and configs[4] at one GPU (config 4 = 5B events needs the 8-GPU box: gpu_multi.sh 8 all)
set -u
mkdir -p gpurun_out
for c in 3 5; do
  echo "== bench config $c N=1"
  timeout 1500 python bench.py --config $c --no-cpu > gpurun_out/bench_c${c}_n1.log 2>&1; grep -iE "error|Traceback" gpurun_out/bench_c${c}_n1.log | head -5
  tail -1 gpurun_out/bench_c${c}_n1.log | tee gpurun_out/bench_c${c}_n1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases_ms'], d.get('e2e'), d['verify'], d.get('gnn_update'))"
done
