#!/bin/bash
# Sustained (power-capped) comparison of the strip shapes: every setting runs 1500 steps back to back per round.
set -u
mkdir -p gpurun_out
timeout 900 python tools/stream_sweep.py --n 8192 --steps 1500 --rounds 2 --check 2 --syncs 0,1 --widths 128,256 --out gpurun_out/r2_sustain_c4.jsonl 2>&1 | tail -6
timeout 900 python tools/stream_sweep.py --n 1024 --batch 64 --steps 800 --rounds 2 --check 2 --syncs 0,1 --widths 128,256 --out gpurun_out/r2_sustain_c5.jsonl 2>&1 | tail -6
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk4_1d_resident -s 3 -c 1 -o gpurun_out/r2_c3 \
      python bench.py --workload c3 --also none --steps 1 --warmup 3 --no-cpu --iters 200 --batch 8192 > gpurun_out/ncu_c3.log 2>&1
tail -2 gpurun_out/ncu_c3.log | cut -c1-200
