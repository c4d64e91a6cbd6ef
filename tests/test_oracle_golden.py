"""Pins the CPU oracle against every golden vector the reference's own tests hold for the hot path
 and against independent checks: SciPy's BLAS gbmv, polynomial exactness of the
radial operator, the truncated cross stencil in 2D, and the README figure."""

import numpy as np
import pytest
from scipy.linalg import blas

from oracle import oracle as O

KINDS = [O.sp, O.dp]

# test/test_nls.f95:50-59 -- make_banded_matrix(n=7, m=5, row=(1..5)); columns of the (5, 7) band
BAND_GOLDEN = np.array([
    [0, 0, 3, 4, 5],
    [0, 2, 3, 4, 5],
    [1, 2, 3, 4, 5],
    [1, 2, 3, 4, 5],
    [1, 2, 3, 4, 5],
    [1, 2, 3, 4, 0],
    [1, 2, 3, 0, 0]], dtype=float).T

# test/test_nls.f95:176-184 -- rgbmv of that band on the identity; row i of z is A e_i ... laid out as
# the reference does: y(i, :) = A x(i, :), z = transpose(reshape(...))
RGBMV_GOLDEN = np.array([
    [3, 4, 5, 0, 0, 0, 0],
    [2, 3, 4, 5, 0, 0, 0],
    [1, 2, 3, 4, 5, 0, 0],
    [0, 1, 2, 3, 4, 5, 0],
    [0, 0, 1, 2, 3, 4, 5],
    [0, 0, 0, 1, 2, 3, 4],
    [0, 0, 0, 0, 1, 2, 3]], dtype=float)

# test/test_nls.f95:102-108 -- make_laplacian_o5(n=7, h=0.01): STALE, pins the 24 h^2 revision
O5_STALE_GOLDEN = np.array([
    [0.00000000, 0.00000000, -1666.66666667, -1250.00000000, -833.33333333, -694.44444444, -625.00000000],
    [0.00000000, 26666.66666667, 13333.33333333, 10000.00000000, 8888.88888889, 8333.33333333, 8000.00000000],
    [-25000.00000000, -12083.33333333, -12500.00000000, -12500.00000000, -12500.00000000, -12500.00000000, -12500.00000000],
    [0.000000000, 3333.33333333, 4444.44444444, 5000.00000000, 5333.33333333, 5555.55555555, 0.00000000],
    [0.000000000, -138.88888888, -208.33333333, -250.00000000, -277.77777777, 0.00000000, 0.00000000]])


@pytest.mark.parametrize("kind", KINDS, ids=["sp", "dp"])
def test_make_banded_matrix_golden(kind):
    mat = kind.make_banded_matrix(7, [1, 2, 3, 4, 5])
    assert mat.shape == (5, 7)
    assert np.array_equal(mat, BAND_GOLDEN)          # reference tolerance is 1e-6; we are exact


@pytest.mark.parametrize("kind", KINDS, ids=["sp", "dp"])
def test_rgbmv_golden(kind):
    op = kind.make_banded_matrix(7, [1, 2, 3, 4, 5])
    eye = np.eye(7)
    for i in range(7):
        y = kind.rgbmv(eye[i], np.zeros(7), 1.0, op)   # the reference leaves y uninitialised; zero is intended
        assert np.array_equal(y, RGBMV_GOLDEN[i])


@pytest.mark.parametrize("kind", KINDS, ids=["sp", "dp"])
def test_make_laplacian_o5_stale_table_matches_24h2_revision(kind):
    legacy = kind.make_laplacian(7, 5, 0.01, legacy24=True)
    current = kind.make_laplacian(7, 5, 0.01)
    # with the 24 h^2 denominator the reference's table is reproduced far inside its own tolerance (1.0) ...
    assert np.abs(legacy - O5_STALE_GOLDEN).max() < 5e-3
    # ... while the shipped 12 h^2 code (nls.f90:177) is 2.7e4 away: the table is stale, not the oracle
    assert np.abs(current - O5_STALE_GOLDEN).max() > 2.0e4


@pytest.mark.parametrize("order", [3, 5, 7])
@pytest.mark.parametrize("kind,gbmv,tol", [(O.sp, blas.sgbmv, 2e-5), (O.dp, blas.dgbmv, 1e-13)], ids=["sp", "dp"])
def test_rgbmv_matches_scipy_blas(kind, gbmv, tol, order):
    n, k = 64, (order - 1) // 2
    rng = np.random.default_rng(0)
    op = kind.make_laplacian(n, order, 0.1)
    x = rng.standard_normal(n).astype(kind.real)
    y0 = rng.standard_normal(n).astype(kind.real)
    ours = kind.rgbmv(x, y0, -1.0, op)
    ref = gbmv(n, n, k, k, -1.0, op, x, beta=1.0, y=y0.copy())
    assert np.abs(ours - ref).max() <= tol * np.abs(ref).max()


@pytest.mark.parametrize("order", [3, 5, 7])
def test_radial_operator_is_exact_on_r_squared(order):
    # laplacian of r^2 in polar coordinates is 4; rows whose stencil is truncated excepted
    n, h, k = 40, 0.1, (order - 1) // 2
    op = O.dp.make_laplacian(n, order, h)
    r = np.arange(n) * h
    y = O.dp.rgbmv(r ** 2, np.zeros(n), 1.0, op)
    assert np.abs(y[:n - k] - 4.0).max() < 1e-11
    if order >= 5:
        y4 = O.dp.rgbmv(r ** 4, np.zeros(n), 1.0, op)
        assert np.abs(y4[:n - k] - 16.0 * r[:n - k] ** 2).max() < 1e-10


def test_radial_operator_first_rows():
    # first rows as exact rationals (times the scale shown there)
    h = 0.1
    op = O.dp.make_laplacian(12, 5, h)

    def row(i):
        return np.array([op[2 + i - j, j] if 0 <= 2 + i - j < 5 else 0.0 for j in range(6)])

    assert np.allclose(row(0)[:3] * 12 * h * h, [-60, 64, -4], rtol=1e-14)
    assert np.allclose(row(1)[:4] * 12 * h * h, [8, -30, 24, -2], rtol=1e-14)
    assert np.allclose(row(2)[:5] * 12 * h * h, [-0.5, 12, -30, 20, -1.5], rtol=1e-14)


def _cross_reference(u, order, h):
    """Dense truncated cross stencil (zero outside the square), SURVEY.md App. A.4."""
    w = {3: ([1, -4, 1], 1.0), 5: ([-1, 16, -60, 16, -1], 12.0), 7: ([2, -27, 270, -980, 270, -27, 2], 180.0)}[order]
    num, den = np.array(w[0], float), w[1] * h * h
    k = (order - 1) // 2
    n = u.shape[0]
    pad = np.zeros((n + 2 * k, n + 2 * k))
    pad[k:k + n, k:k + n] = u
    out = num[k] / den * u
    for s in range(1, k + 1):
        c = num[k + s] / den
        out = out + c * (pad[k + s:k + s + n, k:k + n] + pad[k - s:k - s + n, k:k + n]
                         + pad[k:k + n, k + s:k + s + n] + pad[k:k + n, k - s:k - s + n])
    return out


@pytest.mark.parametrize("order", [3, 5, 7])
def test_rbbmv_is_truncated_cross_stencil(order):
    n, h = 23, 0.2
    rng = np.random.default_rng(1)
    u = rng.standard_normal((n, n))
    blocks, orders = O.dp.make_laplacian_2d(n, order, h)
    y = O.dp.rbbmv(u, np.zeros((n, n)), 1.0, blocks, orders, n).reshape((n, n), order="F")
    ref = _cross_reference(u, order, h)
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()


def test_unsupported_order_is_an_error():
    with pytest.raises(ValueError):
        O.dp.make_laplacian(16, 4, 0.1)
    with pytest.raises(ValueError):
        O.dp.make_laplacian_2d(16, 9, 0.1)


def _c1_inputs(n=400):
    from nls_b200.model import Problem
    from nls_b200.pumping import GaussianRingPumping1D
    m = Problem().model(model="1d", dx=0.1, dt=1e-3, u0=0.1, order=5, num_nodes=n, num_iters=10000,
                        pumping=GaussianRingPumping1D(power=20.0, radius=10.0, variation=3.14))
    return m


def test_readme_figure_is_reproduced():
    # doc/pics/gaussian-ring-pumping.png (README.md:22): |psi(0)|^2 ~ 3.6, reservoir peak ~ 1.48 near r ~ 10
    m = _c1_inputs()
    P, c = m.getPumping(), m.getCoefficients()
    u = O.dp.solve_nls(m.dt, m.dx, 5, 10000, P, c, m.getInitialSolution())
    dens = np.abs(u) ** 2
    assert abs(dens[0] - 3.58) < 0.02 and dens.argmax() == 0
    res = c[11] * P / (c[12] + c[13] * dens)
    assert abs(res.max() - 1.48) < 0.01 and abs(res.argmax() * 0.1 - 10.0) < 0.5
    # single precision (what the reference ships) agrees with its promotion to ~6e-5 (SURVEY.md finding 1)
    us = O.sp.solve_nls(m.dt, m.dx, 5, 10000, P, c, m.getInitialSolution())
    assert np.linalg.norm(us - u) / np.linalg.norm(u) < 5e-4


def test_chemical_potential_consistency():
    m = _c1_inputs(200)
    P, c = m.getPumping(), m.getCoefficients()
    u = O.dp.solve_nls(m.dt, m.dx, 5, 300, P, c, m.getInitialSolution())
    mu = O.dp.chemical_potential_1d(m.dx, P, c, u)
    op = O.dp.make_laplacian(200, 5, m.dx)
    v = O.dp.hamiltonian(P, c, u, op)
    r = np.arange(200) * m.dx
    ref = 1j * np.vdot(u, v * r) / np.vdot(u, u * r)
    assert abs(mu - ref) <= 1e-12 * abs(ref)
    mus = O.sp.chemical_potential_1d(m.dx, P, c, u)
    assert abs(mus - ref) <= 1e-3 * abs(ref)


def _dense_from_band(op):
    """Dense matrix of a BLAS band (ku = kl = k): A(i, j) = op(k + i - j, j)."""
    m, n = op.shape
    k = (m - 1) // 2
    A = np.zeros((n, n))
    for j in range(n):
        for b in range(m):
            i = j + b - k
            if 0 <= i < n:
                A[i, j] = op[b, j]
    return A


@pytest.mark.parametrize("order", [3, 5, 7])
def test_solver_level_second_opinion_1d(order):
    """Nothing in the reference pins hamiltonian / runge_kutta / solve_nls ("parity unpinned").  A second,
    independently written restatement -- dense numpy algebra straight from Appendix A.1 / A.5 (v = (a - i b) u +
    i A u; classical RK4), sharing only the operator table -- must agree with the oracle's C code."""
    n, dx, dt, iters = 60, 0.1, 1e-3, 40
    rng = np.random.default_rng(order)
    P = 10.0 * rng.random(n)
    c = np.zeros(23)
    c[[2, 3, 4, 5, 11, 12, 13]] = [1.0, 1.0, 1.0, 2.8, 0.13658959, 1.0, 0.74626866]
    u = 0.1 + 0.05 * rng.standard_normal(n) + 0.05j * rng.standard_normal(n)
    A = _dense_from_band(O.dp.make_laplacian(n, order, dx))

    def rhs(y):
        usq = np.abs(y) ** 2
        res = c[11] * P / (c[12] + c[13] * usq)
        return ((c[2] * res - c[3]) - 1j * (c[4] * usq + c[5] * res)) * y + 1j * (A @ y)

    assert np.allclose(O.dp.hamiltonian(P, c, u, O.dp.make_laplacian(n, order, dx)), rhs(u), rtol=1e-13, atol=1e-13)
    y = u.copy()
    for _ in range(iters):
        k1 = rhs(y)
        k2 = rhs(y + k1 * dt / 2)
        k3 = rhs(y + k2 * dt / 2)
        k4 = rhs(y + k3 * dt)
        y = y + (k1 + 2 * k2 + 2 * k3 + k4) * dt / 6
    got = O.dp.solve_nls(dt, dx, order, iters, P, c, u)
    assert np.linalg.norm(got - y) / np.linalg.norm(y) <= 1e-13
