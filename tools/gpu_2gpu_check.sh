#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_slab_nccl.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -8 | tee gpurun_out/slab_nccl_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c4 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/scale_c4_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/scale_c2_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 2>&1 | tail -1 > gpurun_out/scale_reference_2gpu.json
