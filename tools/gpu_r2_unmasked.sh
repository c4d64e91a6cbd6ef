#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py tests/test_gpu_engine.py -x -q -m gpu -k "stream or slab or c4 or c5 or batch_2d or large_grid or fused_diag or fixture or oracle" 2>&1 | tail -3
timeout 900 python tools/stream_sweep.py --n 8192 --steps 1000 --rounds 2 --check 2 --syncs 0,1 --widths 128,256 --out gpurun_out/r2_unmasked_c4.jsonl 2>&1 | tail -5
timeout 900 python tools/stream_sweep.py --n 1024 --batch 64 --steps 600 --rounds 2 --check 2 --syncs 1 --widths 128,256 --out gpurun_out/r2_unmasked_c5.jsonl 2>&1 | tail -3
