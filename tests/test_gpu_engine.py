"""Device-resident API: ensembles, continuation, and size-independent properties at full sizes."""

import numpy as np
import pytest

from oracle import oracle as O
from test_gpu_parity import model_1d, model_2d, rel_l2, rough_field

pytestmark = pytest.mark.gpu


def test_ensemble_1d_members_match_oracle():
    """BASELINE config 3 in miniature: pumping power x reservoir rate lattice, one launch."""
    from nls_b200.engine import Ensemble1D
    from nls_b200.model import dimensionless_coefficients, DEFAULT_ORIGINAL_PARAMS
    n, iters = 1000, 300
    base = model_1d(n, iters, power=1.0)
    unit = base.getPumping()
    powers = np.linspace(1.0, 40.0, 6)
    gammas = np.geomspace(0.05, 1.0, 5)
    P, Cs = [], []
    for pw in powers:
        for gr in gammas:
            P.append(pw * unit)
            Cs.append(dimensionless_coefficients(dict(DEFAULT_ORIGINAL_PARAMS, gamma_R=gr)))
    P, Cs = np.array(P), np.array(Cs)
    ens = Ensemble1D(n, base.dx, base.dt, order=5, batch=len(P), pumping=P, coeffs=Cs, u0=0.1)
    got = ens.advance(iters).solution()
    assert got.shape == (30, n)
    for b in range(len(P)):
        want = O.dp.solve_nls(base.dt, base.dx, 5, iters, P[b], Cs[b], 0.1 * np.ones(n))
        assert rel_l2(got[b], want) <= 1e-10, b


def test_continuation_is_bitwise_additive_1d():
    from nls_b200.engine import Ensemble1D
    m = model_1d(400, 0)
    a = Ensemble1D(400, m.dx, m.dt, pumping=m.getPumping(), coeffs=m.getCoefficients(), u0=0.1)
    b = Ensemble1D(400, m.dx, m.dt, pumping=m.getPumping(), coeffs=m.getCoefficients(), u0=0.1)
    a.advance(300)
    b.advance(100).advance(150).advance(50)
    assert np.array_equal(a.solution(), b.solution())


def test_batch_2d_members_match_oracle_and_single_runs():
    from nls_b200.engine import Grid2D
    n, iters = 64, 150
    ms = [model_2d(n, iters, radius=r) for r in (1.0, 2.0, 2.5)]
    P = np.array([m.getPumping() for m in ms])
    c = ms[0].getCoefficients()
    grid = Grid2D(n, 0.1, 1e-3, order=5, batch=3, pumping=P, coeffs=c, u0=0.1)
    got = grid.advance(iters).solution()
    for b, m in enumerate(ms):
        want = O.dp.solve_nls_2d(m.dt, m.dx, 5, iters, P[b], c, m.getInitialSolution())
        assert rel_l2(got[b], want) <= 1e-10
        single = Grid2D(n, 0.1, 1e-3, order=5, batch=1, pumping=P[b], coeffs=c, u0=0.1).advance(iters).solution()[0]
        assert np.array_equal(single, got[b])          # batching never changes a member's arithmetic


def test_c2_full_size_properties():
    """512 x 512 ring pump (BASELINE config 2): properties that need no oracle run.

    * U(1) covariance: H(e^{i a} u) = e^{i a} H(u), so rotating u0 rotates the solution;
    * mirror / transpose symmetry of the ring problem;
    * continuation additivity, bitwise.
    """
    from nls_b200.engine import Grid2D
    n, iters = 512, 200
    m = model_2d(n, iters)
    P, c = m.getPumping(), m.getCoefficients()
    u0 = 0.1 * np.ones((n, n), dtype=complex)
    base = Grid2D(n, m.dx, m.dt, pumping=P, coeffs=c, u0=u0).advance(iters).solution()[0]
    assert np.isfinite(base).all()
    phase = np.exp(0.7j)
    rot = Grid2D(n, m.dx, m.dt, pumping=P, coeffs=c, u0=u0 * phase).advance(iters).solution()[0]
    assert rel_l2(rot, base * phase) <= 1e-12
    assert rel_l2(base.T, base) <= 1e-12 and rel_l2(base[::-1, :], base) <= 1e-12 and rel_l2(base[:, ::-1], base) <= 1e-12
    split = Grid2D(n, m.dx, m.dt, pumping=P, coeffs=c, u0=u0).advance(120).advance(80).solution()[0]
    assert np.array_equal(split, base)


def test_large_grid_short_horizon_vs_oracle():
    """2048 x 2048 (the parity size SURVEY 8d names for config 4), 3 steps against the oracle."""
    from nls_b200.engine import Grid2D
    from nls_b200.pumping import GaussianRingPumping2D
    from nls_b200.model import Problem
    n, iters = 2048, 3
    m = Problem().model(model="2d", dx=0.1, dt=1e-3, u0=0.1, order=5, num_nodes=n, num_iters=iters,
                        pumping=GaussianRingPumping2D(power=20.0, radius=50.0, variation=12.5))
    P, c = m.getPumping(), m.getCoefficients()
    u0 = 0.1 + 0.05 * rough_field((n, n), 11)
    got = Grid2D(n, m.dx, m.dt, pumping=P, coeffs=c, u0=u0).advance(iters).solution()[0]
    want = O.dp.solve_nls_2d(m.dt, m.dx, 5, iters, P, c, u0)
    assert rel_l2(got, want) <= 1e-10


def test_hamiltonian_device_api():
    from nls_b200.engine import Ensemble1D, Grid2D
    m = model_1d(400)
    u = rough_field(400, 4)
    e = Ensemble1D(400, m.dx, m.dt, pumping=m.getPumping(), coeffs=m.getCoefficients(), u0=u)
    want = O.dp.hamiltonian(m.getPumping(), m.getCoefficients(), u, O.dp.make_laplacian(400, 5, m.dx))
    assert rel_l2(e.hamiltonian().cpu().numpy()[0], want) <= 1e-13
    m2 = model_2d(80)
    u2 = rough_field((80, 80), 6)
    g = Grid2D(80, m2.dx, m2.dt, pumping=m2.getPumping(), coeffs=m2.getCoefficients(), u0=u2)
    blocks, orders = O.dp.make_laplacian_2d(80, 5, m2.dx)
    want2 = O.dp.hamiltonian_2d(m2.getPumping(), m2.getCoefficients(), u2, blocks, orders)
    assert rel_l2(g.hamiltonian().cpu().numpy()[0], want2) <= 1e-13


def test_hamiltonian_of_an_ensemble_larger_than_the_grid_y_limit():
    """BASELINE config 3 has 65 536 members: one more than gridDim.y allows, so members ride on gridDim.x."""
    from nls_b200.engine import Ensemble1D
    n, B = 64, 65536 + 3
    m = model_1d(n)
    u = rough_field(n, 9)
    scale = np.linspace(0.5, 1.5, B)
    e = Ensemble1D(n, m.dx, m.dt, batch=B, pumping=m.getPumping(), coeffs=m.getCoefficients(), u0=scale[:, None] * u[None, :])
    got = e.hamiltonian().cpu().numpy()
    op = O.dp.make_laplacian(n, 5, m.dx)
    for b in (0, 1, 65534, 65535, 65536, B - 1):
        want = O.dp.hamiltonian(m.getPumping(), m.getCoefficients(), scale[b] * u, op)
        assert rel_l2(got[b], want) <= 1e-13, b


def test_staged_1d_path_for_systems_larger_than_one_cta_in_a_batch():
    """n > 2048 takes one launch per stage through global memory (members on gridDim.x there too)."""
    from nls_b200.engine import Ensemble1D
    n, B, iters = 2304, 3, 20
    m = model_1d(n, iters)
    P = np.array([m.getPumping() * s for s in (0.5, 1.0, 1.5)])
    e = Ensemble1D(n, m.dx, m.dt, batch=B, pumping=P, coeffs=m.getCoefficients(), u0=0.1)
    got = e.advance(iters).solution()
    for b in range(B):
        want = O.dp.solve_nls(m.dt, m.dx, 5, iters, P[b], m.getCoefficients(), 0.1 * np.ones(n))
        assert rel_l2(got[b], want) <= 1e-10


def _host_diagnostics(m, u, order, dim):
    """The reference's own host-side formulas (model.py mirror) + the oracle's chemical potential."""
    from nls_b200.model import Solution
    sol = Solution(m, u)
    P, c = m.getPumping(), m.getCoefficients()
    if dim == 1:
        op = O.dp.make_laplacian(u.shape[0], order, m.dx)
        v = O.dp.hamiltonian(P, c, u, op)
        w = np.arange(u.shape[0]) * m.dx
    else:
        blocks, orders = O.dp.make_laplacian_2d(u.shape[0], order, m.dx)
        v = O.dp.hamiltonian_2d(P, c, u, blocks, orders)
        w = 1.0
    mu = 1j * np.sum(w * np.conj(u) * v) / np.sum(w * np.conj(u) * u)
    dens = sol.getDensity()
    if dim == 1:
        r = np.linspace(0, u.shape[0] * m.dx, u.shape[0])
        particles = 2 * np.pi * np.sum(dens * r * m.dx)
    else:
        particles = np.sum(dens) * m.dx ** 2
    return mu, sol.getDampingIntegral(), particles, dens.max(), sol.getReservoir().max()


@pytest.mark.parametrize("order", [3, 5, 7])
def test_device_diagnostics_match_host_formulas(order):
    """chemical potential, damping integral, particle number, peak density / reservoir reduced on
    the device against the reference's host formulas (nls/model.py:350-380) and the oracle's H(u); 1e-12."""
    from nls_b200.engine import Ensemble1D, Grid2D
    m1 = model_1d(400, order=order)
    u1 = np.array([rough_field(400, s) * 0.3 + 0.2 for s in (1, 2, 3)])
    P1 = np.array([m1.getPumping() * f for f in (0.5, 1.0, 2.0)])
    e = Ensemble1D(400, m1.dx, m1.dt, order=order, batch=3, pumping=P1, coeffs=m1.getCoefficients(), u0=u1)
    d = e.diagnostics()
    for b in range(3):
        m1.setPumping(lambda *grid, profile=P1[b]: profile)     # the model samples a functor
        mu, damp, part, dmax, rmax = _host_diagnostics(m1, u1[b], order, 1)
        got = (d["chemical_potential"][b], d["damping_integral"][b], d["particles"][b], d["max_density"][b], d["max_reservoir"][b])
        for g, w_ in zip(got, (mu, damp, part, dmax, rmax)):
            assert abs(g - w_) <= 1e-12 * max(abs(w_), 1.0), (order, b, g, w_)
    n = 96
    m2 = model_2d(n, order=order, radius=2.0)
    u2 = np.array([rough_field((n, n), s) * 0.3 + 0.2 for s in (4, 5)])
    g2 = Grid2D(n, m2.dx, m2.dt, order=order, batch=2, pumping=m2.getPumping(), coeffs=m2.getCoefficients(), u0=u2)
    d = g2.diagnostics()
    for b in range(2):
        mu, damp, part, dmax, rmax = _host_diagnostics(m2, u2[b], order, 2)
        got = (d["chemical_potential"][b], d["damping_integral"][b], d["particles"][b], d["max_density"][b], d["max_reservoir"][b])
        for g, w_ in zip(got, (mu, damp, part, dmax, rmax)):
            assert abs(g - w_) <= 1e-12 * max(abs(w_), 1.0), (order, b, g, w_)
    # the reference's entry point (order 5 hard-wired) agrees with the device path
    if order == 5:
        from nls_b200.native import nls
        assert abs(nls.chemical_potential_2d(m2.dx, m2.getPumping(), m2.getCoefficients(), u2[0]) - d["chemical_potential"][0].real) <= 1e-12
    again = g2.diagnostics()
    assert all(np.array_equal(again[k], d[k]) for k in d)           # fixed reduction tree: reproducible


@pytest.mark.parametrize("order,n", [(5, 100), (3, 37), (7, 333)])
def test_device_diagnostics_of_a_large_ensemble_use_a_warp_per_member(order, n):
    """Ensembles of >= 4736 short systems take the warp-per-member kernel (diagnostics.cu): sampled members against
    the host formulas (1e-12), every member against the CTA-per-member kernel through a small ensemble of copies, and
    reproducible bits.  n = 37 leaves lanes without a node in the last round, n = 333 is not a multiple of 32."""
    from nls_b200.engine import Ensemble1D
    B = 5000
    m = model_1d(n, order=order)
    rng = np.random.default_rng(n)
    scale = 0.5 + rng.random(B)
    u = (rough_field(n, 7) * 0.3 + 0.2)[None, :] * scale[:, None]
    P = m.getPumping()[None, :] * (0.5 + rng.random(B))[:, None]
    e = Ensemble1D(n, m.dx, m.dt, order=order, batch=B, pumping=P, coeffs=m.getCoefficients(), u0=u)
    d = e.diagnostics()
    keys = ("chemical_potential", "damping_integral", "particles", "max_density", "max_reservoir")
    for b in (0, 1, 7, 8, 2499, B - 1):
        m.setPumping(lambda *grid, profile=P[b]: profile)
        want = _host_diagnostics(m, u[b], order, 1)
        for k, w_ in zip(keys, want):
            assert abs(d[k][b] - w_) <= 1e-12 * max(abs(w_), 1.0), (order, b, k)
    pick = np.arange(0, B, 97)
    small = Ensemble1D(n, m.dx, m.dt, order=order, batch=len(pick), pumping=P[pick], coeffs=m.getCoefficients(), u0=u[pick])
    ds = small.diagnostics()                                   # CTA-per-member kernel: another summation order
    for k in keys:
        assert np.all(np.abs(d[k][pick] - ds[k]) <= 1e-12 * np.maximum(np.abs(ds[k]), 1.0)), k
    again = e.diagnostics()
    assert all(np.array_equal(again[k], d[k]) for k in d)


@pytest.mark.parametrize("order", [5, 7])
def test_host_buffer_solve_of_a_large_grid_overlaps_its_transfers_and_keeps_the_bits(order):
    """solve_nls_2d from host buffers on grids of >= 2048^2 with >= 160 steps advances two row ranges that run ahead
    of each other while the other range's bytes are on the bus (csrc/api.cu::host_rk4_2d_pipelined): same bits as the
    device-resident time loop, and within 1e-10 of a few oracle-checked rows is implied by the other parity tests."""
    from nls_b200.engine import Grid2D
    from nls_b200.native import nls
    n, iters = 2048, 164
    m = model_2d(n, iters, order=order, radius=30.0)
    rng = np.random.default_rng(order)
    P = np.ascontiguousarray(m.getPumping() * (1.0 + 0.5 * rng.random((n, n))))
    u0 = rough_field((n, n), 3) * 0.05 + 0.1
    got = nls.solve_nls_2d(m.dt, m.dx, order, iters, P, m.getCoefficients(), u0)
    want = Grid2D(n, m.dx, m.dt, order=order, pumping=P, coeffs=m.getCoefficients(), u0=u0).advance(iters).solution()[0]
    assert np.array_equal(got, want)
    again = nls.solve_nls_2d(m.dt, m.dx, order, iters, P, m.getCoefficients(), u0)     # cached graphs, reused pool
    assert np.array_equal(again, want)


def test_continuation_with_changing_pumping_equals_fresh_solves():
    """psi stays on the device across chunks while the pump changes (animation / check loops)."""
    from nls_b200.engine import Grid2D
    n = 64
    m = model_2d(n, radius=1.5)
    P, c = m.getPumping(), m.getCoefficients()
    grid = Grid2D(n, m.dx, m.dt, pumping=P, coeffs=c, u0=0.1)
    grid.advance(50).set_pumping(0.5 * P).advance(30).set_pumping(P).advance(20)
    u = O.dp.solve_nls_2d(m.dt, m.dx, 5, 50, P, c, 0.1 * np.ones((n, n), dtype=complex))
    u = O.dp.solve_nls_2d(m.dt, m.dx, 5, 30, 0.5 * P, c, u)
    u = O.dp.solve_nls_2d(m.dt, m.dx, 5, 20, P, c, u)
    assert rel_l2(grid.solution()[0], u) <= 1e-10


def _ulps(got, want):
    return np.max(np.abs(got - want) / (np.spacing(np.abs(want)) + 1e-300))


def test_device_pumping_profiles_match_host_classes_to_4_ulp():
    """ensemble pumping profiles generated on the device against the host classes (bit-exact to
    the reference's nls/pumping.py); policy: grid and arithmetic identical, exp() within the last place."""
    from nls_b200.engine import device_pumping
    from nls_b200.model import Problem
    from nls_b200 import pumping as H
    n1, n2, dx = 400, 96, 0.1
    powers, radii, vars_ = np.array([1.0, 20.0, 37.5]), np.array([2.0, 10.0, 17.0]), np.array([3.14, 1.0, 5.0])
    m1 = Problem().model(model="1d", dx=dx, dt=1e-3, u0=0.1, order=5, num_nodes=n1, num_iters=1, pumping=H.GaussianPumping1D())
    m2 = Problem().model(model="2d", dx=dx, dt=1e-3, u0=0.1, order=5, num_nodes=n2, num_iters=1, pumping=H.GaussianPumping2D())
    got = device_pumping(1, "ring", n1, dx, powers, vars_, radius=radii).cpu().numpy()
    for b in range(3):
        m1.setPumping(H.GaussianRingPumping1D(power=powers[b], radius=radii[b], variation=vars_[b]))
        assert _ulps(got[b], m1.getPumping()) <= 4
    got = device_pumping(1, "gaussian", n1, dx, powers, vars_, x0=radii).cpu().numpy()
    for b in range(3):
        m1.setPumping(H.GaussianPumping1D(power=powers[b], x0=radii[b], variation=vars_[b]))
        assert _ulps(got[b], m1.getPumping()) <= 4
    got = device_pumping(2, "ring", n2, dx, powers, vars_, radius=radii / 4, x0=0.3, y0=-0.2).cpu().numpy()
    for b in range(3):
        m2.setPumping(H.GaussianRingPumping2D(power=powers[b], x0=0.3, y0=-0.2, variation=vars_[b], radius=radii[b] / 4))
        assert _ulps(got[b], m2.getPumping()) <= 4
    got = device_pumping(2, "gaussian", n2, dx, powers, vars_, x0=0.3, y0=-0.2).cpu().numpy()
    for b in range(3):
        m2.setPumping(H.GaussianPumping2D(power=powers[b], x0=0.3, y0=-0.2, variation=vars_[b]))
        assert _ulps(got[b], m2.getPumping()) <= 4


def test_ensemble_from_device_generated_pumping():
    """An ensemble whose profiles never exist on the host: members match solves fed with the host profiles (the
    last-place differences of exp() stay far below the 1e-10 bar)."""
    from nls_b200.engine import Grid2D, device_pumping
    n, iters = 128, 60
    radii = np.array([1.0, 2.0, 3.0, 4.0])
    P = device_pumping(2, "ring", n, 0.1, 20.0, 3.14 / 4, radius=radii)
    m = model_2d(n, iters)
    grid = Grid2D(n, 0.1, 1e-3, batch=4, pumping=P, coeffs=m.getCoefficients(), u0=0.1).advance(iters)
    from nls_b200.pumping import GaussianRingPumping2D
    for b, r in enumerate(radii):
        m.setPumping(GaussianRingPumping2D(power=20.0, radius=float(r), variation=3.14 / 4))
        want = O.dp.solve_nls_2d(m.dt, m.dx, 5, iters, m.getPumping(), m.getCoefficients(), m.getInitialSolution())
        assert rel_l2(grid.solution()[b], want) <= 1e-10


def test_sweep_front_end_single_gpu():
    """a parameter scan as one ensemble launch; every point agrees with its own oracle solve."""
    from nls_b200.sweep import run_sweep, SweepPoint
    from nls_b200.model import Problem, Solution
    from nls_b200.pumping import GaussianRingPumping2D
    n, iters = 64, 80
    points = [dict(power=p, radius=1.5, variation=0.8, gamma_R=g) for p in (5.0, 20.0) for g in (0.1, 0.242057488654, 0.6)]
    table = run_sweep(points, model="2d", kind="ring", num_nodes=n, num_iters=iters, keep_fields=True)
    assert table["solution"].shape == (6, n, n)
    for i, pt in enumerate(points):
        m = Problem().model(model="2d", dx=0.1, dt=1e-3, u0=0.1, order=5, num_nodes=n, num_iters=iters,
                            pumping=GaussianRingPumping2D(power=pt["power"], radius=1.5, variation=0.8))
        c = SweepPoint(pt).coefficients()
        want = O.dp.solve_nls_2d(m.dt, m.dx, 5, iters, m.getPumping(), c, m.getInitialSolution())
        assert rel_l2(table["solution"][i], want) <= 1e-10
        m.coeffs = c
        assert abs(table["damping_integral"][i] - Solution(m, want).getDampingIntegral()) <= 1e-9 * abs(table["particles"][i])


def test_c4_full_size_properties():
    """8192 x 8192 ring pump (BASELINE config 4, strip-marching kernel), properties that need no oracle run:
    mirror / transpose symmetry of the ring problem, U(1) covariance, continuation additivity (bitwise)."""
    import torch
    from nls_b200.engine import Grid2D, device_pumping
    n = 8192
    m = model_2d(64)
    c = m.getCoefficients()
    P = device_pumping(2, "ring", n, 0.1, 20.0, 50.0, radius=200.0)

    def dev_rel(a, b):
        return float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))

    base = Grid2D(n, 0.1, 1e-3, pumping=P, coeffs=c, u0=0.1).advance(3).psi[0].clone()
    assert bool(torch.isfinite(torch.view_as_real(base)).all())
    assert dev_rel(base.T, base) <= 1e-12 and dev_rel(base.flip(0), base) <= 1e-12 and dev_rel(base.flip(1), base) <= 1e-12
    phase = complex(np.exp(0.7j))
    rot = Grid2D(n, 0.1, 1e-3, pumping=P, coeffs=c, u0=0.1 * phase).advance(3).psi[0]
    assert dev_rel(rot, base * phase) <= 1e-12
    del rot
    split = Grid2D(n, 0.1, 1e-3, pumping=P, coeffs=c, u0=0.1).advance(2).advance(1).psi[0]
    assert bool(torch.equal(split, base))


def test_c5_and_c3_shapes_members_are_independent_of_the_batch():
    """Ensemble configs at their member sizes (1024 x 1024 grids; 1000-node radial systems): a member's result does
    not depend on which batch it runs in, and matches the oracle on a short horizon."""
    from nls_b200.engine import Ensemble1D, Grid2D, device_pumping
    n, iters = 1024, 3
    radii = np.linspace(2.0, 40.0, 8)
    m = model_2d(64)
    c = m.getCoefficients()
    P = device_pumping(2, "ring", n, 0.1, 20.0, 3.14, radius=radii)
    many = Grid2D(n, 0.1, 1e-3, batch=8, pumping=P, coeffs=c, u0=0.1).advance(iters).solution()
    one = Grid2D(n, 0.1, 1e-3, batch=1, pumping=P[5:6], coeffs=c, u0=0.1).advance(iters).solution()[0]
    assert np.array_equal(many[5], one)
    want = O.dp.solve_nls_2d(1e-3, 0.1, 5, iters, P[5].cpu().numpy(), c, 0.1 * np.ones((n, n), dtype=complex))
    assert rel_l2(one, want) <= 1e-10
    n1 = 1000
    m1 = model_1d(n1, power=1.0)
    powers = np.linspace(1.0, 40.0, 64)
    P1 = device_pumping(1, "ring", n1, 0.1, powers, 3.14, radius=10.0)
    ens = Ensemble1D(n1, 0.1, 1e-3, batch=64, pumping=P1, coeffs=m1.getCoefficients(), u0=0.1).advance(500).solution()
    solo = Ensemble1D(n1, 0.1, 1e-3, batch=1, pumping=P1[17:18], coeffs=m1.getCoefficients(), u0=0.1).advance(500).solution()[0]
    assert np.array_equal(ens[17], solo)
    want = O.dp.solve_nls(1e-3, 0.1, 5, 500, P1[17].cpu().numpy(), m1.getCoefficients(), 0.1 * np.ones(n1, dtype=complex))
    assert rel_l2(solo, want) <= 1e-10


def test_advance_until_stops_on_a_device_side_criterion():
    """Steady-state stopping test built on the device diagnostics: same state as plain advance."""
    from nls_b200.engine import Ensemble1D
    m = model_1d(200)
    a = Ensemble1D(200, m.dx, m.dt, batch=2, pumping=np.array([m.getPumping(), 1.5 * m.getPumping()]),
                   coeffs=m.getCoefficients(), u0=0.1)
    steps, converged, history = a.advance_until(rel_tol=1e-3, check_every=250, max_iters=20000)
    assert converged and steps % 250 == 0 and 500 <= steps < 20000 and len(history) == steps // 250
    last, prev = history[-1], history[-2]
    assert np.max(np.abs(last - prev) / np.abs(last)) <= 1e-3
    b = Ensemble1D(200, m.dx, m.dt, batch=2, pumping=np.array([m.getPumping(), 1.5 * m.getPumping()]),
                   coeffs=m.getCoefficients(), u0=0.1).advance(steps)
    assert np.array_equal(a.solution(), b.solution())
    steps, converged, _ = b.advance_until(rel_tol=0.0, check_every=7, max_iters=20)
    assert steps == 20 and not converged


# ---- long / large solves against committed oracle fixtures (tests/golden/make_golden_2d.py) ----------------------
def _against_fixture(case, path="auto"):
    """Run the engine on the fixture's inputs and compare with the dp oracle's reduced field: the strided sub-sample,
    the complex sum and the |psi|^2 sum of EVERY row and column, and the norm -- all to 1e-10 (BASELINE tolerance)."""
    import importlib.util
    import os
    import torch
    from nls_b200.engine import Grid2D
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden_2d", os.path.join(here, "golden", "make_golden_2d.py"))
    gold = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gold)
    fix = np.load(os.path.join(here, "golden", "solve2d_%s.npz" % case))
    P, coeffs, u0, dx, dt, order = gold.inputs(case)
    grid = Grid2D(u0.shape[0], dx, dt, order=order, pumping=P, coeffs=coeffs, u0=u0)
    del P, u0
    psi = grid.advance(int(fix["iters"])).psi[0]
    s = int(fix["stride"])
    a2 = psi.real ** 2 + psi.imag ** 2
    got = dict(sub=psi[::s, ::s], row_sum=psi.sum(dim=1), col_sum=psi.sum(dim=0), row_abs2=a2.sum(dim=1),
               col_abs2=a2.sum(dim=0))
    errs = {}
    for key, value in got.items():
        want = torch.from_numpy(fix[key]).to(value.device)
        errs[key] = float(torch.linalg.vector_norm((value - want).reshape(-1)) / torch.linalg.vector_norm(want.reshape(-1)))
    errs["norm"] = abs(float(torch.sqrt(a2.sum())) - float(fix["norm"])) / float(fix["norm"])
    return errs


def test_c2_full_horizon_against_the_oracle():
    """examples/solve2d.py at 512^2 for its FULL 5000 steps (BASELINE config 2), engine vs dp oracle <= 1e-10."""
    errs = _against_fixture("c2_full")
    assert max(errs.values()) <= 1e-10, errs


def test_c4_size_three_steps_against_the_oracle():
    """8192^2 (BASELINE config 4 size, strip-marching kernel), rough initial field, 3 steps, vs dp oracle <= 1e-10."""
    errs = _against_fixture("c4_3steps")
    assert max(errs.values()) <= 1e-10, errs


def test_c5_member_200_steps_against_the_oracle():
    """One member of BASELINE config 5 (1024^2, pump radius 16.9), rough initial field, 200 steps, <= 1e-10."""
    errs = _against_fixture("c5_member")
    assert max(errs.values()) <= 1e-10, errs


def test_fast_divide_against_the_ieee_divide_on_its_domain():
    """div_fast (csrc/device_math.cuh: MUFU.RCP64H seed + Newton step + residual correction) is on every hot kernel;
    here 2^24 operand pairs from the reservoir's domain (denominator = c13 + c14 |psi|^2 >= c13, numerator c12 P of
    either sign and any magnitude) plus structured hard cases go through it and through __ddiv_rn on the device.
    Contract: never more than ONE ulp from the correctly rounded quotient, and equal to it except for about one pair
    in 10^7 (the corrected quotient is a/b (1 + ~2^-76): it rounds differently only when a/b lies that close to a
    rounding boundary; measured on B200: 1 pair of 2^24)."""
    import ctypes as C
    import torch
    from nls_b200 import _lib
    n = 1 << 24
    g = torch.Generator(device="cuda").manual_seed(7)
    b = 1.0 + torch.rand(n, generator=g, device="cuda", dtype=torch.float64) * 10.0 ** (torch.rand(n, generator=g, device="cuda", dtype=torch.float64) * 8 - 4)
    a = (torch.rand(n, generator=g, device="cuda", dtype=torch.float64) - 0.3) * 10.0 ** (torch.rand(n, generator=g, device="cuda", dtype=torch.float64) * 12 - 8)
    # structured: denominators one ulp around powers of two and around 1, numerators that make ties likely
    k = torch.arange(1 << 16, device="cuda", dtype=torch.float64)
    b[: 1 << 16] = 1.0 + k * 2.0 ** -52
    b[1 << 16: 2 << 16] = 2.0 - (k + 1) * 2.0 ** -52
    b[2 << 16: 3 << 16] = (1.0 + k * 2.0 ** -30) * 2.0 ** 40
    a[: 3 << 16] = torch.where(k.repeat(3) % 2 == 0, 1.0 + k.repeat(3) * 2.0 ** -51, 3.0 - k.repeat(3) * 2.0 ** -50)
    fast, exact = torch.empty_like(a), torch.empty_like(a)
    _lib.call("nlsb_dev_divide_check", C.c_size_t(n), C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()),
              C.c_void_p(fast.data_ptr()), C.c_void_p(exact.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert bool(torch.equal(exact, a / b))                       # torch's divide is the IEEE one as well
    ulps = (fast.view(torch.int64) - exact.view(torch.int64)).abs()
    mismatches = int((ulps != 0).sum())
    assert int(ulps.max()) <= 1 and mismatches <= 16, (mismatches, int(ulps.max()))


# ---- diagnostics fused into the last step's launch ---------------------------------------------
def _diag_close(got, want, tol=1e-12):
    for key in ("chemical_potential", "damping_integral", "particles", "max_density", "max_reservoir"):
        scale = np.maximum(np.abs(want[key]), np.abs(want["particles"]) * 1e-3 if key == "damping_integral" else 1e-300)
        assert np.all(np.abs(got[key] - want[key]) <= tol * scale), (key, got[key], want[key])


def test_fused_diagnostics_1d_equal_the_standalone_pass():
    from nls_b200.engine import Ensemble1D
    m = model_1d(400)
    P = np.array([m.getPumping() * s for s in (0.5, 1.0, 2.0)])
    for order in (3, 5, 7):
        mk = lambda: Ensemble1D(400, m.dx, m.dt, order=order, batch=3, pumping=P, coeffs=m.getCoefficients(), u0=0.1)
        a, b = mk(), mk()
        want = a.advance(36).diagnostics()          # state after 36 steps = the state entering step 37
        a.advance(1)
        got = b.advance(37, diagnostics=True)
        assert got["step"] == 36
        _diag_close(got, want)
        assert np.array_equal(a.solution(), b.solution())            # the fused reduction never changes the field
        again = mk().advance(37, diagnostics=True)
        assert all(np.array_equal(again[k], got[k]) for k in want)   # fixed summation order: reproducible bits


@pytest.mark.parametrize("order,n,batch", [(5, 1024, 1), (5, 1024, 2), (3, 1024, 1), (7, 1056, 1), (5, 96, 2), (5, 1025, 1)])
def test_fused_diagnostics_2d_equal_the_standalone_pass(order, n, batch):
    """1024-wide grids take the strip-marching kernel (reduction inside the last launch); 96 / 513 the tile kernel
    (stand-alone pass before the last step): one API, one meaning."""
    from nls_b200.engine import Grid2D
    from nls_b200.model import dimensionless_coefficients, DEFAULT_ORIGINAL_PARAMS
    m = model_2d(n, radius=min(10.0, n * 0.1 / 4))
    u0 = 0.1 + 0.05 * rough_field((n, n), n)
    c = np.array([dimensionless_coefficients(dict(DEFAULT_ORIGINAL_PARAMS, gamma_R=g)) for g in (0.242057488654, 0.5)][:batch])
    P = np.array([m.getPumping() * s for s in (1.0, 0.7)][:batch])
    mk = lambda: Grid2D(n, m.dx, m.dt, order=order, batch=batch, pumping=P, coeffs=c, u0=u0)
    a, b = mk(), mk()
    want = a.advance(4).diagnostics()
    a.advance(1)
    got = b.advance(5, diagnostics=True)
    assert got["step"] == 4
    _diag_close(got, want)
    assert np.array_equal(a.solution(), b.solution())
    again = mk().advance(5, diagnostics=True)
    assert all(np.array_equal(again[k], got[k]) for k in want)


def test_fused_diagnostics_cost_nothing_extra_on_the_large_grid():
    """A 4096^2 chunk of 50 steps with the reduction riding in its last launch (and its 8 doubles read back) is <= 5 %
    slower than the plain chunk -- the stand-alone pass alone costs 0.7 RK steps on such a grid."""
    import torch
    from nls_b200.engine import Grid2D
    n = 4096
    m = model_2d(n, radius=100.0)
    g = Grid2D(n, m.dx, m.dt, order=5, pumping=m.getPumping(), coeffs=m.getCoefficients(), u0=0.1)

    def timed(fn):
        best = 1e30
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    g.advance(50)
    g.advance(50, diagnostics=True)
    plain = timed(lambda: (g.advance(50), torch.cuda.synchronize()))
    fused = timed(lambda: g.advance(50, diagnostics=True))
    assert fused <= 1.05 * plain, (plain, fused)
