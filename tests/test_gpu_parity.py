"""GPU parity: the CUDA engine (through the C ABI) against the dp oracle on identical inputs.

Bar (BASELINE.json north_star): relative L2 error <= 1e-10 in complex128 after a fixed horizon;
single right-hand-side evaluations and matvecs are held to 1e-13.  The engine uses fused
multiply-adds and its own summation order, the oracle follows nls.f90's order without contraction,
so bitwise equality is not expected -- the tolerances are written next to each assertion.
"""

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

ORIG = dict(R=0.0242057488654, gamma=0.0242057488654, g=0.00162178517398, tilde_g=0.0169440242057,
            gamma_R=0.242057488654)


def rel_l2(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


def model_1d(n, iters=100, order=5, power=20.0):
    from nls_b200.model import Problem
    from nls_b200.pumping import GaussianRingPumping1D
    return Problem().model(model="1d", dx=0.1, dt=1e-3, u0=0.1, order=order, num_nodes=n, num_iters=iters,
                           pumping=GaussianRingPumping1D(power=power, radius=10.0, variation=3.14),
                           original_params=dict(ORIG))


def model_2d(n, iters=100, order=5, radius=10.0, dx=0.1):
    from nls_b200.model import Problem
    from nls_b200.pumping import GaussianRingPumping2D
    return Problem().model(model="2d", dx=dx, dt=1e-3, u0=0.1, order=order, num_nodes=n, num_iters=iters,
                           pumping=GaussianRingPumping2D(power=20.0, radius=radius, variation=3.14),
                           original_params=dict(ORIG))


def rough_field(shape, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) * 0.3


@pytest.fixture(scope="module")
def nls():
    from nls_b200.native import nls as native
    from nls_b200 import _lib
    assert _lib.device_available(), "the GPU tests need a CUDA device"
    return native


# ---- matvecs, reservoir, right-hand side ------------------------------------------------------------

@pytest.mark.parametrize("order", [3, 5, 7])
@pytest.mark.parametrize("n", [7, 33, 400, 1000])
def test_rgbmv(nls, order, n):
    rng = np.random.default_rng(order * 100 + n)
    op = O.dp.make_laplacian(n, order, 0.1)
    x, u = rng.standard_normal(n), rng.standard_normal(n)
    want = O.dp.rgbmv(x, u, -1.0, op)
    nls.rgbmv(x, u, -1.0, op)
    assert np.abs(u - want).max() <= 1e-13 * np.abs(want).max()


@pytest.mark.parametrize("order", [3, 5, 7])
def test_rbbmv(nls, order):
    n = 37
    rng = np.random.default_rng(order)
    blocks, orders = O.dp.make_laplacian_2d(n, order, 0.2)
    x, y = rng.standard_normal(n * n), rng.standard_normal(n * n)
    want = O.dp.rbbmv(x, y, 1.0, blocks, orders, n)
    nls.rbbmv(x, y, 1.0, blocks, orders, n)
    assert np.abs(y - want).max() <= 1e-13 * np.abs(want).max()


@pytest.mark.parametrize("order", [3, 5, 7])
def test_general_block_band_operators(nls, order):
    """rbbmv / hamiltonian_2d / runge_kutta_2d accept ANY blocks memory of the make_laplacian_2d layout
    (nls.f90:408-527): weights that vary along the line -- every entry of the reference's operator scaled by its own
    random factor here -- take the general kernels (kernels_2d.cu) instead of being refused; both orientations (flat
    vectors of rbbmv: line-major; (n, n) arrays: the reference's lines are Fortran columns) against the oracle."""
    n = 29
    rng = np.random.default_rng(40 + order)
    blocks, orders = O.dp.make_laplacian_2d(n, order, 0.2)
    blocks = np.asfortranarray(blocks * (0.5 + rng.random(blocks.shape)))
    wx, wy = np.zeros(7), np.zeros(7)
    from nls_b200 import _lib
    import ctypes as C
    assert _lib.load().nlsb_blocks_to_weights(n, order, blocks.ctypes.data_as(C.c_void_p), orders.ctypes.data_as(C.c_void_p),
                                              wx.ctypes.data_as(C.c_void_p), wy.ctypes.data_as(C.c_void_p)) == -5
    x, y = rng.standard_normal(n * n), rng.standard_normal(n * n)
    want = O.dp.rbbmv(x, y, -1.0, blocks, orders, n)
    nls.rbbmv(x, y, -1.0, blocks, orders, n)
    assert np.abs(y - want).max() <= 1e-13 * np.abs(want).max()
    m = model_2d(n, order=order, radius=0.8)
    c, P = m.getCoefficients(), m.getPumping() * (0.5 + rng.random((n, n)))
    u = rough_field((n, n), order) * 0.3 + 0.2 + 0.1j * rng.standard_normal((n, n))       # no symmetry at all
    assert rel_l2(nls.hamiltonian_2d(P, c, u, blocks, orders), O.dp.hamiltonian_2d(P, c, u, blocks, orders)) <= 1e-13
    got = nls.runge_kutta_2d(1e-4, 0.0, u, blocks, orders, 7, P, c)
    assert rel_l2(got, O.dp.runge_kutta_2d(1e-4, 0.0, u, blocks, orders, 7, P, c)) <= 1e-12
    # a transposed field must NOT give the transposed result: the operator is no longer symmetric under x <-> y
    assert rel_l2(nls.hamiltonian_2d(P.T, c, u.T, blocks, orders).T, nls.hamiltonian_2d(P, c, u, blocks, orders)) > 1e-6


def test_revervoir(nls):
    rng = np.random.default_rng(5)
    c = model_1d(16).getCoefficients()
    p, q = rng.random(1000) * 20, rng.random(1000) * 4
    assert np.array_equal(nls.revervoir(p, c, q), O.dp.revervoir(p, c, q))       # one mul, one fma-free divide
    p2, q2 = rng.random((31, 31)) * 20, rng.random((31, 31)) * 4
    got = nls.revervoir_2d(p2, c, q2)
    assert np.abs(got - O.dp.revervoir(p2, c, q2)).max() <= 1e-15 * 20


@pytest.mark.parametrize("order", [3, 5, 7])
@pytest.mark.parametrize("n", [7, 50, 401])
def test_hamiltonian_1d(nls, order, n):
    m = model_1d(n, order=order)
    u = rough_field(n, n + order)
    op = O.dp.make_laplacian(n, order, m.dx)
    want = O.dp.hamiltonian(m.getPumping(), m.getCoefficients(), u, op)
    got = nls.hamiltonian(m.getPumping(), m.getCoefficients(), u, op)
    assert rel_l2(got, want) <= 1e-13


@pytest.mark.parametrize("order", [3, 5, 7])
@pytest.mark.parametrize("n", [7, 64, 101])
def test_hamiltonian_2d(nls, order, n):
    m = model_2d(n, order=order)
    u = rough_field((n, n), n + order)
    blocks, orders = O.dp.make_laplacian_2d(n, order, m.dx)
    want = O.dp.hamiltonian_2d(m.getPumping(), m.getCoefficients(), u, blocks, orders)
    got = nls.hamiltonian_2d(m.getPumping(), m.getCoefficients(), u, blocks, orders)
    assert got.shape == (n, n) and rel_l2(got, want) <= 1e-13
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


# ---- time stepping ---------------------------------------------------------------------------------

def test_solve_1d_reference_example_full_horizon(nls):
    """BASELINE config 1 = examples/solve1d.py: n=400, 10 000 steps, order 5, ring pump."""
    m = model_1d(400, 10000)
    args = (m.dt, m.dx, 5, 10000, m.getPumping(), m.getCoefficients(), m.getInitialSolution())
    want = O.dp.solve_nls(*args)
    got = nls.solve_nls(*args)
    assert got.dtype == np.complex128 and got.shape == (400,)
    assert rel_l2(got, want) <= 1e-10
    # the shipped single-precision algorithm agrees to f32 drift: both restate the same algorithm
    assert rel_l2(got, O.sp.solve_nls(*args)) <= 5e-4


@pytest.mark.parametrize("order", [3, 5, 7])
@pytest.mark.parametrize("n", [7, 100, 1000, 1100, 2048])
def test_solve_1d_sizes_and_orders(nls, order, n):
    m = model_1d(n, 200, order=order)
    args = (m.dt, m.dx, order, 200, m.getPumping(), m.getCoefficients(), rough_field(n, n) * 0.1 + 0.1)
    assert rel_l2(nls.solve_nls(*args), O.dp.solve_nls(*args)) <= 1e-10
    assert rel_l2(nls.solve_nls_1d(*args), O.dp.solve_nls(*args)) <= 1e-10


def test_solve_1d_larger_than_one_cta(nls):
    n = 3000      # beyond the CTA-resident kernel: staged fallback through global memory
    m = model_1d(n, 50)
    args = (m.dt, m.dx, 5, 50, m.getPumping(), m.getCoefficients(), m.getInitialSolution())
    assert rel_l2(nls.solve_nls(*args), O.dp.solve_nls(*args)) <= 1e-10


def test_runge_kutta_with_caller_supplied_band(nls):
    n = 300
    m = model_1d(n, 100, order=7)
    op = O.dp.make_laplacian(n, 7, m.dx)
    u0 = rough_field(n, 9) * 0.2
    want = O.dp.runge_kutta(m.dt, 0.0, u0, op, 100, m.getPumping(), m.getCoefficients())
    got = nls.runge_kutta(m.dt, 0.0, u0, op, 100, m.getPumping(), m.getCoefficients())
    assert rel_l2(got, want) <= 1e-10


@pytest.fixture(params=["tma32", "tma64", "fused32", "stream", "staged"])
def path_2d(request):
    """Every 2D kernel family the library chooses by itself (TMA tile kernel with 32x32 / 32x64 tiles, plain-load tile
    kernel for odd widths and thin slab strips, strip-marching kernel, per-stage kernels) is held to the same bar.
    The experimental whole-loop kernels get one smoke test each (test_experimental_paths_smoke)."""
    from nls_b200.engine import set_2d_path
    set_2d_path(request.param)
    yield request.param
    set_2d_path("auto")


@pytest.mark.parametrize("order,n,iters", [(3, 40, 300), (5, 7, 50), (5, 96, 501), (5, 129, 200), (7, 64, 300),
                                           (3, 3, 20), (7, 7, 33), (5, 31, 65), (7, 100, 64), (3, 257, 40)])
def test_solve_2d(nls, path_2d, order, n, iters):
    m = model_2d(n, iters, order=order, radius=min(10.0, n * 0.1 / 4))
    args = (m.dt, m.dx, order, iters, m.getPumping(), m.getCoefficients(), rough_field((n, n), n) * 0.05 + 0.1)
    want = O.dp.solve_nls_2d(*args)
    got = nls.solve_nls_2d(*args)
    assert got.shape == (n, n) and got.dtype == np.complex128
    assert rel_l2(got, want) <= 1e-10


@pytest.mark.parametrize("path,order,n,iters", [("tma32_persistent", 5, 96, 120), ("tma64_persistent", 5, 129, 60),
                                                ("resident", 5, 400, 25), ("resident", 7, 260, 20), ("fused64", 5, 129, 100)])
def test_experimental_paths_smoke(nls, path, order, n, iters):
    """Kernel families that are measured slower and never chosen automatically: still correct."""
    from nls_b200.engine import set_2d_path
    m = model_2d(n, iters, order=order, radius=min(10.0, n * 0.1 / 4))
    args = (m.dt, m.dx, order, iters, m.getPumping(), m.getCoefficients(), rough_field((n, n), n) * 0.05 + 0.1)
    try:
        set_2d_path(path)
        got = nls.solve_nls_2d(*args)
    finally:
        set_2d_path("auto")
    assert rel_l2(got, O.dp.solve_nls_2d(*args)) <= 1e-10


def test_fused_and_staged_paths_agree(nls):
    from nls_b200.engine import set_2d_path
    m = model_2d(200, 300)
    args = (m.dt, m.dx, 5, 300, m.getPumping(), m.getCoefficients(), m.getInitialSolution())
    try:
        set_2d_path("staged")
        a = nls.solve_nls_2d(*args)
        set_2d_path("fused")
        b = nls.solve_nls_2d(*args)
    finally:
        set_2d_path("auto")
    assert rel_l2(a, b) <= 1e-13


@pytest.mark.parametrize("order,n,iters", [(5, 700, 3), (5, 513, 4), (3, 300, 5), (7, 260, 3), (5, 241, 2)])
def test_stream_kernel_is_bitwise_equal_to_tile_kernel(nls, order, n, iters):
    """The strip-marching kernel (several strips and chunks, ragged edges) performs each node's arithmetic in the
    same order as the tile kernel: identical bits, and both within 1e-10 of the oracle."""
    from nls_b200.engine import set_2d_path
    m = model_2d(n, iters, order=order)
    rng = np.random.default_rng(n)
    P = m.getPumping() * (1.0 + 0.5 * rng.random((n, n)))
    args = (m.dt, m.dx, order, iters, P, m.getCoefficients(), rough_field((n, n), n + 1) * 0.05 + 0.1)
    try:
        set_2d_path("stream")
        a = nls.solve_nls_2d(*args)
        set_2d_path("fused32")
        b = nls.solve_nls_2d(*args)
    finally:
        set_2d_path("auto")
    assert np.array_equal(a, b)
    assert rel_l2(a, O.dp.solve_nls_2d(*args)) <= 1e-10


def test_stream_kernel_random_geometries_are_bitwise_equal_to_tile_kernel(nls):
    """Random grid sizes x strip widths x rows per CTA x barrier cadence (nlsb_set_stream_tuning): last strips of one
    to eight warps, single-chunk and many-chunk strips, grids narrower than one strip -- always the tile kernel's bits."""
    from nls_b200 import _lib
    from nls_b200.engine import set_2d_path
    rng = np.random.default_rng(2024)
    try:
        for case in range(14):
            order = int(rng.choice([3, 5, 5, 5, 7]))
            n = int(rng.integers(order, 620))
            width = int(rng.choice([0, 128, 256])) if order != 7 else 0
            rows_per_cta = int(rng.choice([0, 0, 40, 90]))
            sync = int(rng.choice([-1, 0, 1]))
            m = model_2d(n, 3, order=order, radius=min(10.0, n * 0.1 / 4))
            P = m.getPumping() * (1.0 + 0.5 * rng.random((n, n)))
            args = (m.dt, m.dx, order, 3, P, m.getCoefficients(), rough_field((n, n), case) * 0.05 + 0.1)
            set_2d_path("fused32")
            want = nls.solve_nls_2d(*args)
            set_2d_path("stream")
            _lib.call("nlsb_set_stream_tuning", sync, width, rows_per_cta)
            got = nls.solve_nls_2d(*args)
            _lib.call("nlsb_set_stream_tuning", -1, 0, 0)
            assert np.array_equal(got, want), (case, order, n, width, rows_per_cta, sync)
    finally:
        _lib.call("nlsb_set_stream_tuning", -1, 0, 0)
        set_2d_path("auto")


@pytest.mark.parametrize("order,n,iters", [(5, 512, 40), (3, 300, 30)])
def test_resident_kernel_is_bitwise_equal_to_tile_kernel(nls, order, n, iters):
    """The register-resident kernel (one patch per CTA, edge nodes exchanged through L2 mailboxes every RK stage)
    performs each node's arithmetic in the order of the tile kernel: identical bits after `iters` steps."""
    from nls_b200.engine import set_2d_path
    m = model_2d(n, iters, order=order)
    rng = np.random.default_rng(n)
    P = m.getPumping() * (1.0 + 0.5 * rng.random((n, n)))
    args = (m.dt, m.dx, order, iters, P, m.getCoefficients(), rough_field((n, n), n + 1) * 0.05 + 0.1)
    try:
        set_2d_path("resident")
        a = nls.solve_nls_2d(*args)
        set_2d_path("fused32")
        b = nls.solve_nls_2d(*args)
    finally:
        set_2d_path("auto")
    assert np.array_equal(a, b)


def test_stream_kernel_batch_with_member_coefficients(nls):
    from nls_b200.engine import Grid2D, set_2d_path
    from nls_b200.model import dimensionless_coefficients
    n, iters = 300, 6
    ms = [model_2d(n, iters, radius=r) for r in (3.0, 5.0, 7.5)]
    P = np.array([m.getPumping() for m in ms])
    c = np.array([dimensionless_coefficients(dict(ORIG, gamma_R=g)) for g in (0.1, 0.242057488654, 0.7)])
    out = {}
    try:
        for path in ("stream", "fused32"):
            set_2d_path(path)
            out[path] = Grid2D(n, 0.1, 1e-3, order=5, batch=3, pumping=P, coeffs=c, u0=0.1).advance(iters).solution()
    finally:
        set_2d_path("auto")
    assert np.array_equal(out["stream"], out["fused32"])
    for b, m in enumerate(ms):
        want = O.dp.solve_nls_2d(m.dt, m.dx, 5, iters, P[b], c[b], m.getInitialSolution())
        assert rel_l2(out["stream"][b], want) <= 1e-10


def test_solve_2d_nonsymmetric_input_keeps_index_convention(nls):
    # a[i, j] <-> Fortran a(i+1, j+1): an asymmetric pump/initial field must come back un-transposed
    n = 48
    m = model_2d(n, 100)
    rng = np.random.default_rng(2)
    P = m.getPumping() * (1.0 + 0.5 * rng.random((n, n)))
    u0 = rough_field((n, n), 3) * 0.2
    args = (m.dt, m.dx, 5, 100, P, m.getCoefficients(), u0)
    assert rel_l2(nls.solve_nls_2d(*args), O.dp.solve_nls_2d(*args)) <= 1e-10
    blocks, orders = O.dp.make_laplacian_2d(n, 5, m.dx)
    got = nls.runge_kutta_2d(m.dt, 0.0, u0, blocks, orders, 100, P, m.getCoefficients())
    assert rel_l2(got, O.dp.solve_nls_2d(*args)) <= 1e-10


def test_solve_2d_c2_size_short_horizon(nls):
    """BASELINE config 2 grid (512 x 512, ring pump) against the oracle for 20 steps."""
    m = model_2d(512, 20)
    args = (m.dt, m.dx, 5, 20, m.getPumping(), m.getCoefficients(), m.getInitialSolution())
    assert rel_l2(nls.solve_nls_2d(*args), O.dp.solve_nls_2d(*args)) <= 1e-10


def test_chemical_potentials(nls):
    m = model_1d(400, 500)
    P, c = m.getPumping(), m.getCoefficients()
    u = O.dp.solve_nls(m.dt, m.dx, 5, 500, P, c, m.getInitialSolution())
    want = O.dp.chemical_potential_1d(m.dx, P, c, u)
    got = nls.chemical_potential_1d(m.dx, P, c, u)
    assert isinstance(got, complex) and abs(got - want) <= 1e-12 * abs(want)
    m2 = model_2d(64, 100)
    P2 = m2.getPumping()
    u2 = O.dp.solve_nls_2d(m2.dt, m2.dx, 5, 100, P2, c, m2.getInitialSolution())
    want2 = O.dp.chemical_potential_2d(m2.dx, P2, c, u2)
    got2 = nls.chemical_potential_2d(m2.dx, P2, c, u2)
    assert isinstance(got2, float) and abs(got2 - want2) <= 1e-12 * abs(want2)


def test_object_api_end_to_end(nls, capsys):
    m = model_1d(400, 1000)
    sol = m.solve()
    sol.report()
    assert "1000 iteration on 400 grid nodes" in capsys.readouterr().out
    want = O.dp.solve_nls(m.dt, m.dx, 5, 1000, m.getPumping(), m.getCoefficients(), m.getInitialSolution())
    assert rel_l2(sol.getSolution(), want) <= 1e-10
    mu = m.getChemicalPotential(sol)
    assert abs(mu - O.dp.chemical_potential_1d(m.dx, m.getPumping(), m.getCoefficients(), want)) <= 1e-9 * abs(mu)
    m2 = model_2d(64, 100)
    sol2 = m2.solve()
    want2 = O.dp.solve_nls_2d(m2.dt, m2.dx, 5, 100, m2.getPumping(), m2.getCoefficients(), m2.getInitialSolution())
    assert rel_l2(sol2.getSolution(), want2) <= 1e-10


def test_errors(nls):
    from nls_b200.native import error
    with pytest.raises(error):
        nls.solve_nls(1e-3, 0.1, 4, 1, np.ones(16), np.ones(23), np.ones(16))     # bad order
    with pytest.raises(error):
        nls.solve_nls_2d(1e-3, 0.1, 7, 1, np.ones((5, 5)), np.ones(23), np.ones((5, 5)))   # n < order
    with pytest.raises(ValueError):
        nls.solve_nls(1e-3, 0.1, 5, 1, np.ones(16), np.ones(22), np.ones(16))     # coeffs must be 23
    with pytest.raises(ValueError):
        nls.solve_nls(1e-3, 0.1, 5, 1, np.ones(16), np.ones(23), np.ones(17))
    # zero iterations returns the input
    u0 = rough_field(32, 1)
    assert np.array_equal(nls.solve_nls(1e-3, 0.1, 5, 0, np.ones(32), np.ones(23), u0), u0)
