"""Host logic of the overlapped host-buffer solve (csrc/api.cu::host_rk4_2d_pipelined): the split points / head starts
nlsb_solve_nls_2d_plan reports, and a numpy emulation of the two-range schedule on a generic radius-4k update rule --
two row ranges running ahead of each other in two ping-pong buffers must reproduce the plain time loop bit for bit."""

import ctypes as C

import numpy as np
import pytest

from nls_b200 import _lib


def _plan(n, order, iters):
    out = [C.c_int() for _ in range(5)]
    assert _lib.load().nlsb_solve_nls_2d_plan(n, order, iters, *[C.byref(v) for v in out]) == 0
    return [v.value for v in out]


@pytest.mark.parametrize("order", [3, 5, 7])
def test_plan_invariants(order):
    halo = 2 * (order - 1)
    for n in (2048, 2050, 3000, 4096, 8192, 16384):
        for iters in (160, 164, 200, 333, 1000, 5000, 100000):
            on, r_top, s_up, r_dn, s_dn = _plan(n, order, iters)
            assert on == 1
            assert s_up % 2 == 0 and s_dn % 2 == 0 and s_up >= 0 and s_dn >= 0          # the middle part starts / ends in psi
            assert s_up + s_dn <= 0.8 * iters + 1e-9
            assert 0 < r_top <= n // 2 + 1 and r_top + halo * s_up <= 3 * n // 4         # the first transfer is the smaller one
            assert r_dn == n - r_top and r_dn + halo * s_dn <= n - halo                  # head start ends inside the grid
            assert r_top * n >= 1 << 20                                                  # every range is a large launch
    assert _plan(2047, order, 1000)[0] == 0 and _plan(1024, order, 1000)[0] == 0 and _plan(8192, order, 100)[0] == 0


def _step(src, dst, lo, hi, radius):
    """Rows [lo, hi) of a nonlinear update that reads `radius` rows either side (zeros outside the array)."""
    n = src.shape[0]
    pad = np.zeros((n + 2 * radius,) + src.shape[1:])
    pad[radius:radius + n] = src
    acc = np.zeros((hi - lo,) + src.shape[1:])
    for d in range(-radius, radius + 1):
        acc += np.cos(0.3 * d) * pad[radius + lo + d:radius + hi + d]
    dst[lo:hi] = np.tanh(0.2 * acc) + 0.5 * src[lo:hi]


@pytest.mark.parametrize("n,radius,s_up,s_dn,r_top,iters", [(96, 4, 6, 4, 24, 20), (97, 2, 8, 8, 30, 16), (64, 8, 2, 2, 16, 7)])
def test_two_ranges_running_ahead_reproduce_the_plain_loop(n, radius, s_up, s_dn, r_top, iters):
    rng = np.random.default_rng(n)
    u0 = rng.standard_normal((n, 5))
    a, b = u0.copy(), np.zeros_like(u0)
    for _ in range(iters):                       # the plain loop
        _step(a, b, 0, n, radius)
        a, b = b, a
    want = a
    r_dn = n - r_top
    bufs = [np.full_like(u0, np.nan), np.full_like(u0, np.nan)]
    upto = r_top + radius * s_up                 # rows uploaded first
    bufs[0][:upto] = u0[:upto]

    def range_step(j, lo, hi):
        _step(bufs[(j - 1) % 2], bufs[j % 2], lo, hi, radius)

    for j in range(1, s_up + 1):                 # top range ahead; rows >= upto are still NaN ("on the bus")
        range_step(j, 0, r_top + radius * (s_up - j))
    assert not np.isnan(bufs[0][:r_top]).any()
    bufs[0][upto:] = u0[upto:]                   # the rest arrives
    for j in range(1, s_up + 1):
        range_step(j, r_top + radius * (s_up - j), n)
    for j in range(1, iters - s_up - s_dn + 1):  # middle: whole grid, starting in buffer 0 (s_up is even)
        _step(bufs[(j - 1) % 2], bufs[j % 2], 0, n, radius)
    if (iters - s_up - s_dn) % 2:
        bufs[0][:] = bufs[1]                     # enqueue_rk4_2d leaves the state in psi
    for j in range(1, s_dn + 1):
        range_step(j, 0, r_dn + radius * (s_dn - j))
    top = bufs[0][:r_dn].copy()                  # leaves for the host now
    for j in range(1, s_dn + 1):
        range_step(j, r_dn + radius * (s_dn - j), n)
    got = np.concatenate([top, bufs[0][r_dn:]])
    assert np.array_equal(got, want)
