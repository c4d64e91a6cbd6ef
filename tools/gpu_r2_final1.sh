#!/bin/bash
# Driver-like single-GPU run: smoke, bench (engine arm, then reference arm) with the driver's arguments.
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_final_1gpu.json 2> gpurun_out/r2_bench_final_1gpu.err
tail -3 gpurun_out/r2_bench_final_1gpu.err; python tools/show_bench.py gpurun_out/r2_bench_final_1gpu.json
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_final_reference.json ) 2>&1 | tail -3
cut -c1-400 gpurun_out/r2_bench_final_reference.json
