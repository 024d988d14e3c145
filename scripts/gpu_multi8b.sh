#!/bin/bash
# round 2: 8 GPUs again after the tier-S rows and the edge-capacity fix: configs 2 and 4
set -u
N=8
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
show='import json,sys
d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["phases_ms"], (d.get("e2e") or {}).get("value"), d["verify"], d.get("multi_gpu"), d.get("gnn_update"))'
echo "== bench config 2 N=$N"
timeout 400 $T bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c2_n${N}b.log 2>&1; grep -iE "^bench.py|Error" gpurun_out/bench_c2_n${N}b.log | head -5
tail -1 gpurun_out/bench_c2_n${N}b.log | tee gpurun_out/bench_c2_n${N}b.json | python -c "$show"
echo "== bench config 4 N=$N"
timeout 300 $T bench.py --gpus $N --config 4 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c4_n$N.log 2>&1; grep -iE "^bench.py|Error" gpurun_out/bench_c4_n$N.log | head -5
tail -1 gpurun_out/bench_c4_n$N.log | tee gpurun_out/bench_c4_n$N.json | python -c "$show"
