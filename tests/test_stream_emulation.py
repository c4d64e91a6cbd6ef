"""CPU check of the strip-marching 2D kernel's indexing: ``tests/emu/stream_emu.cu`` runs the kernel's own
per-thread body (``nls_b200/csrc/stream_2d_core.cuh``) thread by thread on the host -- TMA batches and barriers
replaced by their sequential meaning -- and the result is compared with the dp oracle.  The GPU tests
(``test_gpu_parity.py``) check the real kernel; this one keeps the register-window / ring arithmetic honest on a
machine without a GPU."""

import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "stream_emu.cu")
LIB = os.path.join(HERE, "emu", "libstream_emu.so")
CORE = os.path.join(HERE, "..", "nls_b200", "csrc", "stream_2d_core.cuh")

ORIG = dict(R=0.0242057488654, gamma=0.0242057488654, g=0.00162178517398, tilde_g=0.0169440242057,
            gamma_R=0.242057488654)


@pytest.fixture(scope="module")
def emu():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc is needed to compile the host emulation")
    deps = [SRC, CORE, os.path.join(os.path.dirname(CORE), "device_math.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call([nvcc, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-o", LIB, SRC],
                              stderr=subprocess.DEVNULL)
    return C.CDLL(LIB)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _weights(order, dx):
    from nls_b200 import _lib
    wx, wy = np.zeros(7), np.zeros(7)
    assert _lib.load().nlsb_cross_weights(order, C.c_double(dx), _p(wx), _p(wy)) == 0
    return wx, wy


def _steps(emu, order, threads, u0, P, coeffs, dt, dx, steps, chunk_rows=0, rows_window=None):
    n, m = u0.shape
    wx, wy = _weights(order, dx)
    a = np.array(u0, dtype=complex, order="C", copy=True)
    b = np.zeros_like(a)
    Pc = np.ascontiguousarray(P, dtype=float)
    coeffs = np.ascontiguousarray(coeffs, dtype=float)
    for _ in range(steps):
        rc = emu.emu_stream_step(order, threads, n, m, 0, n, 0, n, chunk_rows, _p(a), _p(Pc), _p(coeffs), _p(wx), _p(wy),
                                 C.c_double(dt), _p(b))
        assert rc == 0
        a, b = b, a
    return a


@pytest.mark.parametrize("order,n,threads,chunk_rows", [(5, 40, 64, 0), (5, 131, 64, 0), (5, 300, 256, 0), (5, 300, 128, 0),
                                                        (5, 97, 64, 28), (5, 7, 64, 0), (3, 131, 64, 0),
                                                        (3, 300, 256, 0), (7, 131, 64, 0), (7, 300, 192, 0)])
def test_emulated_stream_kernel_matches_oracle(emu, order, n, threads, chunk_rows):
    from nls_b200.model import dimensionless_coefficients
    coeffs = dimensionless_coefficients(dict(ORIG))
    rng = np.random.default_rng(order * 1000 + n)
    x = np.linspace(-1, 1, n)
    u0 = 0.1 + 0.05 * rng.standard_normal((n, n)) + 0.03j * rng.standard_normal((n, n))
    P = 20 * np.exp(-((x[None, :] * 3) ** 2 + (x[:, None] * 2 - 0.3) ** 2)) + rng.random((n, n))
    got = _steps(emu, order, threads, u0, P, coeffs, 1e-3, 0.1, 2, chunk_rows)
    want = O.dp.solve_nls_2d(1e-3, 0.1, order, 2, P, coeffs, u0)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-13


def test_emulated_slab_rows_match_the_single_domain(emu):
    """Local arrays that are a row window of the global grid (slab decomposition): same bits as the full run."""
    from nls_b200.model import dimensionless_coefficients
    coeffs = np.ascontiguousarray(dimensionless_coefficients(dict(ORIG)))
    n, order, halo = 96, 5, 8
    rng = np.random.default_rng(5)
    u0 = np.ascontiguousarray(0.1 + 0.05 * rng.standard_normal((n, n)) + 0.03j * rng.standard_normal((n, n)))
    P = np.ascontiguousarray(5.0 * rng.random((n, n)))
    full = _steps(emu, order, 64, u0, P, coeffs, 1e-3, 0.1, 1)
    wx, wy = _weights(order, 0.1)
    lo, hi = 40, 75                                   # owned rows of the slab
    a0, a1 = lo - halo, min(hi + halo, n)
    local_u = np.ascontiguousarray(u0[a0:a1])
    local_P = np.ascontiguousarray(P[a0:a1])
    out = np.zeros_like(local_u)
    rc = emu.emu_stream_step(order, 64, a1 - a0, n, a0, n, lo - a0, hi - a0, 0, _p(local_u), _p(local_P), _p(coeffs),
                             _p(wx), _p(wy), C.c_double(1e-3), _p(out))
    assert rc == 0
    assert np.array_equal(out[lo - a0:hi - a0], full[lo:hi])
