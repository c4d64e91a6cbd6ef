#!/usr/bin/env python
"""Throughput of the auxiliary device kernels against the measured HBM roofline, KERNEL time
only: every entry point is called through the C ABI on preallocated device buffers (no allocation, no host copy in
the timed region), CUDA events on the launching stream, 3 warm-up calls, inputs larger than L2.

* diagnostics_2d / diagnostics_1d: the stand-alone pass, 24 algorithmic bytes per node (psi 16 + P 8);
* pumping_2d: profile generation, 8 bytes per node written;
* fused diagnostics: what `advance(iters, diagnostics=True)` costs over `advance(iters)` on C4 (8192^2) and on a
  C3-shaped ensemble -- the reduction rides inside the last step's launch.
Prints one JSON line per measurement."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import ORIG, measured_hbm_peak  # noqa: E402
from nls_b200 import _lib  # noqa: E402
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


def ptr(t):
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def main():
    peak, src = measured_hbm_peak()
    c = dimensionless_coefficients(dict(ORIG))
    lib = _lib.load()
    n = 8192
    P = device_pumping(2, "ring", n, 0.1, 20.0, 50.0, radius=200.0)
    grid = Grid2D(n, 0.1, 1e-3, pumping=P, coeffs=c, u0=0.1)
    grid.advance(2)
    out8 = torch.empty((1, 8), dtype=torch.float64, device="cuda")
    scratch = torch.empty(lib.nlsb_dev_diagnostics_scratch(1), dtype=torch.uint8, device="cuda")
    wx, wy = grid.wx.ctypes.data_as(C.c_void_p), grid.wy.ctypes.data_as(C.c_void_p)
    t = timed(lambda: _lib.call("nlsb_dev_diagnostics_2d", 1, n, n, 5, 0.1, wx, wy, ptr(grid.pumping), ptr(grid.coeffs),
                                ptr(grid.psi), ptr(scratch), ptr(out8), stream()))
    print(json.dumps({"kernel": "diagnostics_2d + finish (8192^2, kernel time)", "seconds": t, "GB/s": 24.0 * n * n / t / 1e9,
                      "frac_of_hbm_peak": 24.0 * n * n / t / 1e9 / peak, "peak": src}))
    params = np.ascontiguousarray(np.array([[20.0, 0.0, 0.0, 50.0, 200.0]]))
    t = timed(lambda: _lib.call("nlsb_dev_pumping_profiles", 2, 1, 1, n, 0.1, params.ctypes.data_as(C.c_void_p), ptr(P), stream()))
    print(json.dumps({"kernel": "pumping_2d (8192^2 ring; 40-byte parameter upload + stream sync inside the entry point)",
                      "seconds": t, "GB/s": 8.0 * n * n / t / 1e9, "frac_of_hbm_peak": 8.0 * n * n / t / 1e9 / peak}))
    plain = timed(lambda: grid.advance(50), reps=3)
    fused = timed(lambda: grid.advance(50, diagnostics=True), reps=3)
    print(json.dumps({"measurement": "C4 8192^2, 50-step chunk: advance vs advance(diagnostics=True) incl. reading the 8 doubles back",
                      "plain_s": plain, "fused_s": fused, "overhead": fused / plain - 1.0,
                      "standalone_pass_s": timed(grid.diagnostics, reps=3)}))
    del grid, P
    B, n1 = 65536, 1000
    P1 = device_pumping(1, "ring", n1, 0.1, np.linspace(1, 40, B), 3.14, radius=10.0)
    ens = Ensemble1D(n1, 0.1, 1e-3, batch=B, pumping=P1, coeffs=c, u0=0.1)
    ens.advance(2)
    out8 = torch.empty((B, 8), dtype=torch.float64, device="cuda")
    scratch = torch.empty(lib.nlsb_dev_diagnostics_scratch(B), dtype=torch.uint8, device="cuda")
    t = timed(lambda: _lib.call("nlsb_dev_diagnostics_1d", B, n1, 5, 0.1, ptr(ens.taps), ptr(ens.pumping), ptr(ens.coeffs),
                                ptr(ens.psi), ptr(scratch), ptr(out8), stream()))
    print(json.dumps({"kernel": "diagnostics_1d (65536 x 1000, kernel time)", "seconds": t, "GB/s": 24.0 * B * n1 / t / 1e9,
                      "frac_of_hbm_peak": 24.0 * B * n1 / t / 1e9 / peak}))
    plain = timed(lambda: ens.advance(200), reps=3)
    fused = timed(lambda: ens.advance(200, diagnostics=True), reps=3)
    print(json.dumps({"measurement": "C3 65536 x 1000, 200-step chunk: advance vs advance(diagnostics=True) incl. reading 4 MB back",
                      "plain_s": plain, "fused_s": fused, "overhead": fused / plain - 1.0}))


if __name__ == "__main__":
    main()
