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
    assert converged and steps % 250 == 0 and 250 <= steps < 20000 and len(history) == steps // 250 + 1
    last, prev = history[-1], history[-2]
    assert np.max(np.abs(last - prev) / np.abs(last)) <= 1e-3
    b = Ensemble1D(200, m.dx, m.dt, batch=2, pumping=np.array([m.getPumping(), 1.5 * m.getPumping()]),
                   coeffs=m.getCoefficients(), u0=0.1).advance(steps)
    assert np.array_equal(a.solution(), b.solution())
    steps, converged, _ = b.advance_until(rel_tol=0.0, check_every=7, max_iters=20)
    assert steps == 20 and not converged
