#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/sanity_gpu_tests.log
timeout 600 python bench.py --workload c1 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/sanity_c1.json
timeout 600 python bench.py --workload c3 --grid-n 400 --batch 16384 --iters 500 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/sanity_c3_n400.json
