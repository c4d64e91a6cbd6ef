#!/bin/bash
# GPU check of the streaming 2D kernel: parity tests, throughput against the tile kernel over grid sizes, one ncu capture.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py -x -q -m gpu -k "stream or slab" 2>&1 | tail -5 | tee gpurun_out/stream_tests.log
for n in 8192 4096 2048 1024; do
  for path in stream auto; do
    timeout 600 python bench.py --workload c4 --grid-n $n --path $path --steps 3 --warmup 3 --no-cpu --iters 40 2>&1 | tail -1 > gpurun_out/bench_n${n}_$path.json
  done
done
timeout 600 python bench.py --workload c5 --batch 32 --path stream --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_c5_stream.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk4_stream -s 14 -c 1 -o gpurun_out/r1_stream_c4 \
    python bench.py --workload c4 --path stream --steps 1 --warmup 3 --no-cpu --iters 4 > gpurun_out/ncu_stream.log 2>&1
tail -3 gpurun_out/ncu_stream.log
