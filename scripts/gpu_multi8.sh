#!/bin/bash
# 8 GPUs of one box: merge parity at 2/4/8 ranks, bench configs 2 (weak), 4 and 5 (strong).
# Budget note: an 8-GPU call is charged 8x its wall time; give gpurun a --timeout that the remaining budget covers.
set -u
N=8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "== multi-GPU tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_n$N.txt 2>&1; grep -E "^E |passed|failed|Error|error" gpurun_out/pytest_multi_n$N.txt | head -10
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
show='import json,sys
d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["phases_ms"], (d.get("e2e") or {}).get("value"), d["verify"], d.get("multi_gpu"), d.get("gnn_update"))'
echo "== bench config 2 N=$N"
timeout 600 $T bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_c2_n$N.log 2>&1; grep -iE "^bench.py|Error" gpurun_out/bench_c2_n$N.log | head -5
tail -1 gpurun_out/bench_c2_n$N.log | tee gpurun_out/bench_c2_n$N.json | python -c "$show"
echo "== bench config 4 N=$N"
timeout 420 $T bench.py --gpus $N --config 4 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c4_n$N.log 2>&1; grep -iE "^bench.py|Error" gpurun_out/bench_c4_n$N.log | head -5
tail -1 gpurun_out/bench_c4_n$N.log | tee gpurun_out/bench_c4_n$N.json | python -c "$show"
echo "== bench config 5 N=$N"
timeout 420 $T bench.py --gpus $N --config 5 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c5_n$N.log 2>&1; grep -iE "^bench.py|Error" gpurun_out/bench_c5_n$N.log | head -5
tail -1 gpurun_out/bench_c5_n$N.log | tee gpurun_out/bench_c5_n$N.json | python -c "$show"
