#!/bin/bash
# Round-end validation on one B200: GPU tests, smoke, the default bench line (both arms), launch list + full ncu
# captures of the dominant kernels, auxiliary kernel throughput.
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/final_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
timeout 900 python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/final_bench_reference.json
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/final_bench.json
for wl in c1 c3 c4; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/final_$wl.json
done
timeout 900 python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/final_c5.json
timeout 600 python tools/bench_aux.py 2>&1 | grep '^{' > gpurun_out/r1_aux_kernels.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r1_launches_c2_default.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --iters 256 > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk4_step_fused -s 300 -c 1 -o gpurun_out/r1_tile_c2_default \
    python bench.py --steps 1 --warmup 3 --no-cpu --iters 128 > gpurun_out/ncu_tile.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk4_stream -s 14 -c 1 -o gpurun_out/r1_stream_c4 \
    python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu --iters 4 > gpurun_out/ncu_stream.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rk4_1d_resident -s 3 -c 1 -o gpurun_out/r1_1d_c3 \
    python bench.py --workload c3 --batch 8192 --iters 50 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_1d.log 2>&1
tail -c 600 gpurun_out/final_bench.json
