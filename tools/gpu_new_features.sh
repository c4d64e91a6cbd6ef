#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/engine_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk4_resident -s 2 -c 1 -o gpurun_out/r1_resident_c2 \
    python bench.py --workload c2 --path resident --steps 1 --warmup 3 --no-cpu --iters 200 > gpurun_out/ncu_resident.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk4_step_fused -s 200 -c 1 -o gpurun_out/r1_tile_c2 \
    python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu --iters 128 > gpurun_out/ncu_tile.log 2>&1
tail -2 gpurun_out/ncu_tile.log
