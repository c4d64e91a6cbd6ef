#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 tests/run_slab_nccl.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -5 | tee gpurun_out/slab_nccl_8gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 \
      bench.py --gpus 8 --workload c4 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/scale_c4_8gpu.json
tail -c 300 gpurun_out/scale_c4_8gpu.json
