#!/bin/bash
# Round 2: the mid-iteration barrier flavour (sync 3) under the parity tests, then round-robin timing of the flavours.
set -u
mkdir -p gpurun_out
NLSB_STREAM_SYNC=3 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py tests/test_gpu_engine.py -x -q -m gpu -k "stream or slab or c4 or c5 or batch_2d or large_grid" 2>&1 | tail -3 | tee gpurun_out/r2_stream_tests_sync3.log
timeout 900 python tools/stream_sweep.py --n 8192 --steps 30 --syncs 0,1,3 --widths 128,256 --out gpurun_out/r2_sweep2_c4.jsonl 2>&1 | tail -8
timeout 600 python tools/stream_sweep.py --n 1024 --batch 32 --steps 30 --syncs 0,1,3 --widths 128,224,256 --out gpurun_out/r2_sweep2_c5.jsonl 2>&1 | tail -10
timeout 600 python tools/stream_sweep.py --n 4096 --order 3 --steps 30 --syncs 0,3 --widths 128,256 --out gpurun_out/r2_sweep2_o3.jsonl 2>&1 | tail -5
timeout 600 python tools/stream_sweep.py --n 4096 --order 7 --steps 20 --syncs 0,2,3 --widths 0 --out gpurun_out/r2_sweep2_o7.jsonl 2>&1 | tail -5
