#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/gpu_tests.log
for wl in c2 c3 c4; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/quick_$wl.json
done
timeout 900 python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu --batch 64 2>&1 | tail -1 > gpurun_out/quick_c5b64.json
