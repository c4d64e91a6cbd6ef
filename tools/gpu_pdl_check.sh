#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/gpu_tests.log
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_c2_auto.json
timeout 600 python bench.py --workload c2 --grid-n 400 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_n400_auto.json
timeout 600 python bench.py --workload c5 --batch 32 --grid-n 256 --iters 500 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_b32n256_auto.json
