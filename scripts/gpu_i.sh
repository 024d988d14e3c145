#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== red merge micro"; timeout 120 scripts/micro/red_merge | tee gpurun_out/micro_red_merge.txt
