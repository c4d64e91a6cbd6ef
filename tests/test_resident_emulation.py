"""CPU check of the register-resident 2D kernel: ``tests/emu/resident_emu.cu`` runs the kernel's own per-thread
body (``nls_b200/csrc/resident_2d_core.cuh``) on the host, with the CTAs scheduled as coroutines in a seeded
random order -- a CTA advances only when every halo packet it needs carries the sequence number it waits for,
exactly like the spin-wait on the device.  This checks the patch layout, the halo-cell mapping and the two-parity
mailbox protocol (a packet overwritten before it was consumed, or a deadlock, is reported) without a GPU."""

import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "resident_emu.cu")
LIB = os.path.join(HERE, "emu", "libresident_emu.so")
CSRC = os.path.join(HERE, "..", "nls_b200", "csrc")

ORIG = dict(R=0.0242057488654, gamma=0.0242057488654, g=0.00162178517398, tilde_g=0.0169440242057,
            gamma_R=0.242057488654)


@pytest.fixture(scope="module")
def emu():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc is needed to compile the host emulation")
    deps = [SRC, os.path.join(CSRC, "resident_2d_core.cuh"), os.path.join(CSRC, "device_math.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call([nvcc, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-o", LIB, SRC],
                              stderr=subprocess.DEVNULL)
    return C.CDLL(LIB)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _run(emu, order, u0, P, coeffs, steps, capacity, seed, batch=1):
    from nls_b200 import _lib
    wx, wy = np.zeros(7), np.zeros(7)
    assert _lib.load().nlsb_cross_weights(order, C.c_double(0.1), _p(wx), _p(wy)) == 0
    a = np.array(u0, dtype=complex, order="C", copy=True)
    out = np.zeros_like(a)
    Pc = np.ascontiguousarray(P, dtype=float)
    cc = np.ascontiguousarray(coeffs, dtype=float)
    layout = (C.c_int * 4)()
    rows, cols = a.shape[-2:]
    rc = emu.emu_resident(order, batch, rows, cols, C.c_longlong(capacity), steps, seed, _p(a), _p(Pc), _p(cc), _p(wx),
                          _p(wy), C.c_double(1e-3), _p(out), layout)
    return rc, out, list(layout)


def _inputs(n, seed, rows=None):
    rows = rows or n
    rng = np.random.default_rng(seed)
    x, y = np.linspace(-1, 1, n), np.linspace(-1, 1, rows)
    u0 = 0.1 + 0.05 * rng.standard_normal((rows, n)) + 0.03j * rng.standard_normal((rows, n))
    P = 20 * np.exp(-((x[None, :] * 3) ** 2 + (y[:, None] * 2 - 0.3) ** 2)) + rng.random((rows, n))
    return u0, P


@pytest.mark.parametrize("order,n,capacity", [(5, 40, 148), (5, 131, 148), (5, 300, 148), (5, 97, 20), (5, 7, 148),
                                              (5, 512, 148), (3, 131, 148), (3, 3, 148), (7, 131, 148), (7, 97, 20),
                                              (7, 7, 148)])
def test_emulated_resident_kernel_matches_oracle(emu, order, n, capacity):
    from nls_b200.model import dimensionless_coefficients
    coeffs = dimensionless_coefficients(dict(ORIG))
    u0, P = _inputs(n, order * 1000 + n)
    rc, got, layout = _run(emu, order, u0, P, coeffs, 3, capacity, seed=n)
    assert rc == 0, "emulation status %d (1 = does not fit, 3 = packet overwritten early, 4 = deadlock)" % rc
    npx, npy, pw, ph = layout
    assert npx * npy <= capacity and pw <= 128 and ph <= 15
    want = O.dp.solve_nls_2d(1e-3, 0.1, order, 3, P, coeffs, u0)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-13


def test_schedule_does_not_change_a_bit(emu):
    from nls_b200.model import dimensionless_coefficients
    coeffs = dimensionless_coefficients(dict(ORIG))
    u0, P = _inputs(150, 9)
    runs = [_run(emu, 5, u0, P, coeffs, 4, 148, seed)[1] for seed in (1, 2, 3)]
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])


def test_batch_members_are_independent(emu):
    from nls_b200.model import dimensionless_coefficients
    c = np.array([dimensionless_coefficients(dict(ORIG, gamma_R=g)) for g in (0.1, 0.7)])
    ins = [_inputs(60, s) for s in (1, 2)]
    u0 = np.array([i[0] for i in ins])
    P = np.array([i[1] for i in ins])
    rc, got, layout = _run(emu, 5, u0, P, c, 3, 148, seed=5, batch=2)
    assert rc == 0
    for b in range(2):
        want = O.dp.solve_nls_2d(1e-3, 0.1, 5, 3, P[b], c[b], u0[b])
        assert np.linalg.norm(got[b] - want) / np.linalg.norm(want) <= 1e-13


def test_grids_that_do_not_fit_are_refused(emu):
    from nls_b200.model import dimensionless_coefficients
    coeffs = dimensionless_coefficients(dict(ORIG))
    u0, P = _inputs(600, 1)
    assert _run(emu, 5, u0, P, coeffs, 1, 148, seed=1)[0] == 1
