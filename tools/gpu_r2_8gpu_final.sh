#!/bin/bash
# Round 2 (final kernels), N GPUs (default 8): slabs bitwise against one GPU (exchange inside the step launch), the
# driver-like bench line, exchange-kernel / halo-depth variants, and the slab timeline.
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29519 tests/run_slab_nccl.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -6 | tee gpurun_out/r2_slab_check_${N}gpu.log
timeout 400 $TR --master-port 29520 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_final_${N}gpu.json 2> gpurun_out/r2_bench_final_${N}gpu.err
tail -2 gpurun_out/r2_bench_final_${N}gpu.err | cut -c1-300; python tools/show_bench.py gpurun_out/r2_bench_final_${N}gpu.json
if [ "${2:-}" = "variants" ]; then
for cfg in "off 4" "on 2"; do set -- $cfg
  timeout 200 $TR --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --also none --no-cpu --fused-exchange $1 --halo-steps $2 > gpurun_out/r2_bench_${N}gpu_$1_$2.json 2> gpurun_out/r2_bench_${N}gpu_$1_$2.err
  python tools/show_bench.py gpurun_out/r2_bench_${N}gpu_$1_$2.json | head -3
done
timeout 300 $TR --master-port 29522 tools/slab_timeline.py --out gpurun_out/r2_slab_timeline_${N}gpu.json 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3
fi
