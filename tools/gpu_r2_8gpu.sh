#!/bin/bash
# Round 2, 8 GPUs: where a slab step's time goes, and BASELINE config 5 at its FULL horizon (1e5 RK steps) once.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29520 tools/slab_timeline.py --out gpurun_out/r2_slab_timeline_8gpu.json 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3
timeout 600 $TR --master-port 29521 bench.py --gpus 8 --workload c5 --iters 100000 --warmup-iters 100 --steps 1 --warmup 3 --also none > gpurun_out/r2_c5_full_horizon_8gpu.json 2> gpurun_out/r2_c5_full.err
tail -2 gpurun_out/r2_c5_full.err; cut -c1-400 gpurun_out/r2_c5_full_horizon_8gpu.json
