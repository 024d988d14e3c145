#!/bin/bash
set -u
mkdir -p gpurun_out
ALZ_LIB_PATH=alaz_b200/lib/libalazgpu_prof.so timeout 300 python scripts/ingest_prof.py | tee gpurun_out/ingest_phases.json
