#!/bin/bash
# GPU check of the register-resident 2D kernel: parity tests, C2 throughput against the tile kernel, one ncu capture.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "resident" 2>&1 | tail -8 | tee gpurun_out/resident_tests.log
for path in resident auto; do
  timeout 600 python bench.py --workload c2 --path $path --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_c2_$path.json
done
timeout 600 python bench.py --workload c2 --grid-n 400 --path resident --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_n400_resident.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk4_resident -s 2 -c 1 -o gpurun_out/r1_resident_c2 \
    python bench.py --workload c2 --path resident --steps 1 --warmup 3 --no-cpu --iters 200 > gpurun_out/ncu_resident.log 2>&1
tail -3 gpurun_out/ncu_resident.log
