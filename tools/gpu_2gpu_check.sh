#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slabs.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/slab_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_slab_nccl.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -8 | tee gpurun_out/slab_nccl_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c4 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/scale_c4_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload c4 --n 2048 --iters 200 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/scale_n2048_2gpu.json
tail -c 300 gpurun_out/scale_c4_2gpu.json
