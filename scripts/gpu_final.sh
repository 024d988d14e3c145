#!/bin/bash
# round 2 wrap-up on one GPU: the whole GPU test suite, configs 3 and 5, evidence for profiles/, the default bench lines
set -u
mkdir -p gpurun_out
echo "== gpu tests (all)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_final.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
show='import json,sys
d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["phases_ms"], d["roofline"]["frac"], d["verify"], d.get("gnn_update"))'
for c in 3 5; do
  echo "== bench config $c N=1"
  timeout 600 python bench.py --config $c --no-cpu --no-e2e > gpurun_out/bench_c${c}_n1.log 2>&1; grep -iE "^bench.py|Error" gpurun_out/bench_c${c}_n1.log | head -5
  tail -1 gpurun_out/bench_c${c}_n1.log | tee gpurun_out/bench_c${c}_n1.json | python -c "$show"
done
bash scripts/gpu_profiles.sh
