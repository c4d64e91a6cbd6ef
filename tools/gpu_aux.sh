#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/engine_tests.log
timeout 600 python tools/bench_aux.py 2>&1 | grep '^{\|Error\|error' > gpurun_out/r1_aux_kernels.jsonl
cat gpurun_out/r1_aux_kernels.jsonl
