#!/bin/bash
# Fresh numbers for every BASELINE config on one B200 (bench.py lines kept under gpurun_out/measure_*.json).
set -u
mkdir -p gpurun_out
for wl in c1 c2 c3 c4; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/measure_$wl.json
done
timeout 900 python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/measure_c5.json
timeout 600 python bench.py --workload c2 --path tma64 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/measure_c2_tma64.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/measure_reference.json
nproc > gpurun_out/nproc.txt
