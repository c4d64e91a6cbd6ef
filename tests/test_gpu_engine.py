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
