#!/bin/bash
# Partitioned configs on N GPUs of one box: C4 (slabs + NCCL halo exchange), C3 and C5 (members sharded).
set -u
N=${1:-8}
WLS=${2:-"c4 c3 c5"}
mkdir -p gpurun_out
port=29600
for wl in $WLS; do
  port=$((port+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --workload $wl --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/scale_${wl}_${N}gpu.json
  tail -c 300 gpurun_out/scale_${wl}_${N}gpu.json; echo
done
