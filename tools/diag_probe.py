#!/usr/bin/env python
"""Where the cost of `advance(iters, diagnostics=True)` goes: single steps with and without the
fused reduction, timed with CUDA events on the launching stream, on C4 (8192^2) and on the C3 ensemble; then chunks
of 50 / 200 steps.  One JSON line per measurement."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ORIG  # noqa: E402
from nls_b200.engine import Ensemble1D, Grid2D, device_pumping  # noqa: E402
from nls_b200.model import dimensionless_coefficients  # noqa: E402


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    c = dimensionless_coefficients(dict(ORIG))
    n = 8192
    P = device_pumping(2, "ring", n, 0.1, 20.0, 50.0, radius=200.0)
    grid = Grid2D(n, 0.1, 1e-3, pumping=P, coeffs=c, u0=0.1)
    for iters, reps in ((1, 20), (2, 20), (50, 4), (200, 2)):
        plain = timed(lambda: grid.advance(iters), reps)
        fused = timed(lambda: grid.advance(iters, diagnostics=True), reps)
        print(json.dumps({"workload": "c4 8192^2", "iters_per_call": iters, "plain_ms": plain, "with_diagnostics_ms": fused,
                          "extra_ms": fused - plain, "overhead": fused / plain - 1.0}), flush=True)
    del grid, P
    B, n1 = 65536, 1000
    P1 = device_pumping(1, "ring", n1, 0.1, np.linspace(1, 40, B), 3.14, radius=10.0)
    ens = Ensemble1D(n1, 0.1, 1e-3, batch=B, pumping=P1, coeffs=c, u0=0.1)
    for iters, reps in ((1, 10), (50, 4), (200, 2)):
        plain = timed(lambda: ens.advance(iters), reps)
        fused = timed(lambda: ens.advance(iters, diagnostics=True), reps)
        print(json.dumps({"workload": "c3 65536 x 1000", "iters_per_call": iters, "plain_ms": plain, "with_diagnostics_ms": fused,
                          "extra_ms": fused - plain, "overhead": fused / plain - 1.0}), flush=True)


if __name__ == "__main__":
    main()
