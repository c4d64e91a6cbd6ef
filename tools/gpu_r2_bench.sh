#!/bin/bash
# Round 2: full GPU suite, the default bench line (C4 + sub-records), the reference arm, ncu captures of the dominant
# kernels of every config, and the launch list of the default command.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2_gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 600 gpurun_out/r2_bench_default.err; cut -c1-1500 gpurun_out/r2_bench_default.json
if [ "${1:-}" = "ncu" ]; then
for wl in c4 c5 c2 c3; do
  skip=14
  case $wl in
    c4) pat=rk4_stream; extra="--iters 4";;
    c5) pat=rk4_stream; extra="--iters 4 --batch 32";;
    c2) pat=rk4_step_fused; extra="--iters 40";;
    c3) pat=rk4_1d_resident; extra="--iters 200 --batch 8192"; skip=3;;
  esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -o gpurun_out/r2_$wl \
      python bench.py --workload $wl --also none --steps 1 --warmup 3 --no-cpu $extra > gpurun_out/ncu_$wl.log 2>&1
  tail -2 gpurun_out/ncu_$wl.log | cut -c1-300
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/r2_launches_c4_default.csv \
    python bench.py --also none --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | cut -c1-300
fi
