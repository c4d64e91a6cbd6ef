#!/bin/bash
# Round 2, call 1: regression of the GPU suite, the stream kernel's synchronisation flavours under the parity tests,
# and the (sync x width) sweep on the C4 grid and a C5 slice.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_gpu_tests.log
for sync in 1 2; do
  NLSB_STREAM_SYNC=$sync timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py tests/test_gpu_engine.py -x -q -m gpu -k "stream or slab or c4 or c5 or batch_2d or large_grid" 2>&1 | tail -3 | tee gpurun_out/r2_stream_tests_sync$sync.log
done
timeout 900 python tools/stream_sweep.py --n 8192 --steps 30 --out gpurun_out/r2_sweep_c4.jsonl 2>&1 | tail -40
timeout 600 python tools/stream_sweep.py --n 1024 --batch 32 --steps 30 --out gpurun_out/r2_sweep_c5.jsonl 2>&1 | tail -40
timeout 600 python tools/stream_sweep.py --n 8192 --steps 30 --syncs 0,2 --widths 256 --iters 0,110,210,410,810 --out gpurun_out/r2_sweep_iters.jsonl 2>&1 | tail -12
timeout 600 python tools/stream_sweep.py --n 4096 --order 3 --steps 30 --widths 0,128,192,256 --out gpurun_out/r2_sweep_o3.jsonl 2>&1 | tail -14
timeout 600 python tools/stream_sweep.py --n 4096 --order 7 --steps 20 --widths 0 --out gpurun_out/r2_sweep_o7.jsonl 2>&1 | tail -5
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
