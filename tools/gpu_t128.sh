#!/bin/bash
set -u
mkdir -p gpurun_out
NLSB_STREAM_T=128 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stream" 2>&1 | tail -2 | tee gpurun_out/t128_tests.log
for T in 128 256; do
  NLSB_STREAM_T=$T timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/t${T}_c4.json
  NLSB_STREAM_T=$T timeout 600 python bench.py --workload c5 --batch 64 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/t${T}_c5b64.json
  NLSB_STREAM_T=$T timeout 600 python bench.py --workload c4 --grid-n 2048 --iters 100 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/t${T}_n2048.json
done
