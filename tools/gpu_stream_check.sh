#!/bin/bash
# First GPU check of the streaming 2D kernel: parity tests, C4 / C5 throughput against the tile kernel, one ncu capture.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stream" 2>&1 | tail -15 | tee gpurun_out/stream_tests.log
for path in stream auto; do
  timeout 600 python bench.py --workload c4 --path $path --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_c4_$path.json
done
for path in stream auto; do
  timeout 600 python bench.py --workload c5 --batch 32 --path $path --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_c5_$path.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk4_stream -s 14 -c 1 -o gpurun_out/r1_stream_c4 \
    python bench.py --workload c4 --path stream --steps 1 --warmup 3 --no-cpu --iters 4 > gpurun_out/ncu_stream.log 2>&1
tail -3 gpurun_out/ncu_stream.log
