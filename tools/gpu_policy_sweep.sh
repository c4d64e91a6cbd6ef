#!/bin/bash
# Measurements behind the automatic kernel choice: tile shapes on small grids, streaming vs tile kernel per order.
set -u
mkdir -p gpurun_out
for n in 256 400 512 768; do
  for path in tma32 tma64; do
    timeout 300 python bench.py --workload c2 --grid-n $n --path $path --iters 2000 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/policy_n${n}_$path.json
  done
done
for order in 3 7; do
  for path in stream tma32; do
    timeout 300 python bench.py --workload c4 --grid-n 4096 --order $order --path $path --iters 20 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/policy_o${order}_$path.json
  done
done
