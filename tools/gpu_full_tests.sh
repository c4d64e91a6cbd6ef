#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -25 | tee gpurun_out/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
