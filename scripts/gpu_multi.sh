#!/bin/bash
# round 2: multi-GPU merge (one all-gather of fixed-capacity blocks), N = number of GPUs on the box
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_n$N.txt 2>&1; grep -E "^E |passed|failed|Error|error" gpurun_out/pytest_multi_n$N.txt | head -30
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== bench config 2 N=$N"
timeout 900 $T bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_c2_n$N.log 2>&1; grep -iE "error|Traceback|status" gpurun_out/bench_c2_n$N.log | head -10; tail -1 gpurun_out/bench_c2_n$N.log | tee gpurun_out/bench_c2_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases_ms'], d.get('e2e'), d['verify'], d.get('multi'))"
if [ "${2:-}" = "all" ]; then
for c in 3 4 5; do
  echo "== bench config $c N=$N"
  timeout 1500 $T bench.py --gpus $N --config $c --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_c${c}_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['phases_ms'], d.get('e2e'), d['verify'], d.get('multi'))"
done
fi
