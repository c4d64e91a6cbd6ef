#!/usr/bin/env python
"""Throughput of the auxiliary device kernels against the measured HBM roofline:
diagnostics (24 algorithmic bytes per node: psi 16 + P 8) and pumping-profile generation (8 bytes per node).
Prints one JSON line per kernel; CUDA events, 3 warm-up calls, inputs larger than L2."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ORIG, measured_hbm_peak  # noqa: E402
from nls_b200.engine import Ensemble1D, Grid2D, device_pumping  # noqa: E402
from nls_b200.model import dimensionless_coefficients  # noqa: E402


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def main():
    peak, src = measured_hbm_peak()
    c = dimensionless_coefficients(dict(ORIG))
    n = 8192
    P = device_pumping(2, "ring", n, 0.1, 20.0, 50.0, radius=200.0)
    grid = Grid2D(n, 0.1, 1e-3, pumping=P, coeffs=c, u0=0.1)
    grid.advance(2)
    t = timed(grid.diagnostics)
    print(json.dumps({"kernel": "diagnostics_2d (8192^2, incl. the 64-byte D2H of the result)", "seconds": t,
                      "GB/s": 24.0 * n * n / t / 1e9, "frac_of_hbm_peak": 24.0 * n * n / t / 1e9 / peak, "peak": src}))
    t = timed(lambda: device_pumping(2, "ring", n, 0.1, 20.0, 50.0, radius=200.0))
    print(json.dumps({"kernel": "pumping_2d (8192^2 ring, incl. allocation and parameter upload)", "seconds": t,
                      "GB/s": 8.0 * n * n / t / 1e9, "frac_of_hbm_peak": 8.0 * n * n / t / 1e9 / peak}))
    B, n1 = 65536, 1000
    P1 = device_pumping(1, "ring", n1, 0.1, np.linspace(1, 40, B), 3.14, radius=10.0)
    ens = Ensemble1D(n1, 0.1, 1e-3, batch=B, pumping=P1, coeffs=c, u0=0.1)
    ens.advance(2)
    t = timed(ens.diagnostics)
    print(json.dumps({"kernel": "diagnostics_1d (65536 x 1000, incl. the 4 MB D2H of the result)", "seconds": t,
                      "GB/s": 24.0 * B * n1 / t / 1e9, "frac_of_hbm_peak": 24.0 * B * n1 / t / 1e9 / peak}))


if __name__ == "__main__":
    main()
