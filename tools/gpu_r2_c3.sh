#!/bin/bash
# C3 (1D ensemble): the two flavours of the resident kernel after the live-node select was removed.
set -u
mkdir -p gpurun_out
for one in 0 1; do
  NLSB_1D_ONE_CTA=$one timeout 600 python bench.py --workload c3 --also none --steps 3 --warmup 3 --no-cpu --iters 2000 2>/dev/null | tail -1 > gpurun_out/r2_c3_one$one.json
  python tools/show_bench.py gpurun_out/r2_c3_one$one.json | head -4
done
timeout 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_parity.py -x -q -m gpu -k "1d or ensemble or hamiltonian" 2>&1 | tail -3
