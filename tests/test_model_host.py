"""Host layer: Problem/Model/Solution mirror the reference's object API (nls/model.py, nls/solver.py)."""

import numpy as np
import pytest

import nls_b200
from nls_b200 import model as M
from nls_b200 import solver as S
from nls_b200.pumping import GaussianPumping, GaussianPumping1D, GaussianRingPumping1D, GaussianRingPumping2D, GridPumping
from oracle import oracle as O


def test_dimensionless_coefficients():
    c = M.dimensionless_coefficients(dict(M.DEFAULT_ORIGINAL_PARAMS))
    assert c.shape == (23,) and c.dtype == np.float64
    # coeffs[2..5] = (1, 1, 1, 2.8), coeffs[11..13] = (0.13658959..., 1, 0.74626866...)
    assert np.array_equal(c[[0, 1, 2, 3, 4]], np.ones(5))
    assert abs(c[5] - 2.8) < 1e-9 and c[5] == 4.0 * 0.0169440242057 / 0.0242057488654
    assert abs(c[11] - 0.13658959) < 1e-8 and c[12] == 1.0 and abs(c[13] - 0.74626866) < 1e-8
    assert c[10] == 0.0 and c[14] == 0.0 and not c[15:].any() and not c[6:10].any()
    # same float64 expression tree as ref model.py:142-159
    o = M.DEFAULT_ORIGINAL_PARAMS
    phi0 = np.sqrt(o["gamma"] / (2.0 * o["g"]))
    n0 = 2.0 / (o["R"] * phi0)
    assert c[11] == 1.0 / (n0 * o["gamma_R"]) and c[13] == o["R"] * phi0 ** 2 / o["gamma_R"]


def test_factory_defaults_follow_the_reference():
    m1 = M.Problem().model()
    assert isinstance(m1, M.Model1D) and isinstance(m1.solver, S.Solver1D)
    assert (m1.dx, m1.dt, m1.order, m1.num_nodes, m1.num_iters) == (0.1, 1e-3, 5, 1000, 100000)
    assert m1.getInitialSolution().shape == (1000,) and np.all(m1.getInitialSolution() == 0.1)
    assert isinstance(m1.pumping, GaussianPumping1D)      # documented divergence, App. B #7
    m2 = M.Problem().model(model="2d")
    assert isinstance(m2, M.Model2D) and isinstance(m2.solver, S.Solver2D)
    assert (m2.order, m2.num_nodes, m2.num_iters) == (3, 40, 1000)
    assert m2.getInitialSolution().shape == (40, 40)
    assert isinstance(m2.pumping, GaussianPumping)
    with pytest.raises(Exception):
        M.Problem().model(model="3d")


def test_original_params_are_completed_not_replaced():
    m = M.Problem().model(original_params={"R": 0.05})
    assert m.originals["R"] == 0.05 and m.originals["g"] == M.DEFAULT_ORIGINAL_PARAMS["g"]


def test_pumping_grid_spacing_quirk():
    # ref model.py:222-224: linspace(0, n*dx, n) -> spacing n*dx/(n-1), not dx
    m = M.Problem().model(model="1d", num_nodes=400, dx=0.1, pumping=GridPumping(None))
    m.setPumping(type("Echo", (), {"__call__": lambda self, x: x})())
    x = m.getPumping()
    assert x[0] == 0.0 and x[-1] == 400 * 0.1 and abs((x[1] - x[0]) - 0.1002506265664) < 1e-12
    m2 = M.Problem().model(model="2d", num_nodes=6, dx=0.5)
    m2.setPumping(type("Echo2", (), {"__call__": lambda self, x, y: x + 10 * y})())
    g = m2.getPumping()
    xs = np.linspace(-1.5, 1.5, 6)
    assert np.array_equal(g, xs[None, :] + 10 * xs[:, None])   # meshgrid 'xy': gx[i, j] = x[j], gy[i, j] = x[i]


def test_callable_u0_in_1d():
    m = M.Problem().model(model="1d", num_nodes=50, dx=0.2, u0=lambda x: np.exp(-x))
    assert np.array_equal(m.getInitialSolution(), np.exp(-np.linspace(0.0, 0.2 * 50, 50)))


class _OracleNative(object):
    """Records the positional arguments the facade passes and answers with the dp oracle."""

    def __init__(self):
        self.calls = []

    def solve_nls(self, *args):
        self.calls.append(("solve_nls", args))
        return O.dp.solve_nls(*args)

    def solve_nls_2d(self, *args):
        self.calls.append(("solve_nls_2d", args))
        return O.dp.solve_nls_2d(*args)

    def chemical_potential_1d(self, *args):
        self.calls.append(("chemical_potential_1d", args))
        return O.dp.chemical_potential_1d(*args)

    def chemical_potential_2d(self, *args):
        self.calls.append(("chemical_potential_2d", args))
        return O.dp.chemical_potential_2d(*args)


def test_solver_facade_argument_order(monkeypatch, capsys):
    fake = _OracleNative()
    monkeypatch.setattr(S, "nls", fake)
    m = M.Problem().model(model="1d", num_nodes=64, num_iters=7, dt=2e-3, dx=0.2, order=3,
                          pumping=GaussianRingPumping1D(power=3.0, radius=4.0, variation=1.0))
    sol = m.solve()
    name, args = fake.calls[0]
    # ref solver.py:27-34: (dt, dx, order, iters, pumping, coeffs, u0)
    assert name == "solve_nls" and args[:4] == (2e-3, 0.2, 3, 7)
    assert np.array_equal(args[4], m.getPumping()) and args[5] is m.getCoefficients() and args[6] is m.getInitialSolution()
    assert isinstance(sol, M.Solution) and sol.getSolution().shape == (64,) and sol.getElapsedTime() > 0.0
    sol.report()
    assert "7 iteration on 64 grid nodes" in capsys.readouterr().out
    # solve(num_iters) overrides the model's iteration count (ref solver.py:22-23)
    m.solve(3)
    assert fake.calls[1][1][3] == 3 and m.getNumberOfIterations() == 3
    mu = m.getChemicalPotential(sol)
    name, args = fake.calls[2]
    assert name == "chemical_potential_1d" and args[0] == 0.2 and args[3] is sol.getSolution()   # ref :45-50
    assert isinstance(mu, complex) or np.iscomplexobj(mu)

    m2 = M.Problem().model(model="2d", num_nodes=16, num_iters=2, order=5,
                           pumping=GaussianRingPumping2D(power=3.0, radius=0.5, variation=0.3))
    sol2 = m2.solve()
    assert fake.calls[-1][0] == "solve_nls_2d" and sol2.getSolution().shape == (16, 16)
    res = sol2.getReservoir()
    c = m2.coeffs
    assert np.allclose(res, c[11] * m2.getPumping() / (c[12] + c[13] * np.abs(sol2.getSolution()) ** 2))
    assert np.isfinite(sol2.getDampingIntegral()) and np.isfinite(sol.getDampingIntegral())


def test_mat_store_restore_roundtrip(tmp_path, monkeypatch):
    monkeypatch.setattr(S, "nls", _OracleNative())
    m = M.Problem().model(model="1d", num_nodes=32, num_iters=4, order=5,
                          pumping=GaussianRingPumping1D(power=2.0, radius=1.0, variation=0.5))
    sol = m.solve()
    path = str(tmp_path / "run.mat")
    sol.store(path, label="t", desc="roundtrip")
    again = M.Problem().model(filename=path)
    assert isinstance(again, M.Model1D)
    assert again.num_nodes == 32 and again.num_iters == 4 and again.order == 5
    assert np.array_equal(again.getCoefficients(), m.getCoefficients())
    assert np.array_equal(again.getPumping(), m.getPumping())      # restored as GridPumping (ref model.py:300)
    assert again.originals == pytest.approx(m.originals)
    back = M.Solution(again).restore(path)
    assert np.array_equal(back.getSolution(), sol.getSolution())
    # continuation pattern of tools/check.py:34-36
    again.setInitialSolution(back.getSolution())
    more = again.solve(2)
    assert more.getSolution().shape == (32,)


def test_checkpoints_written_by_the_reference_restore(tmp_path, monkeypatch):
    """The reference stores str(type(model)) = "<class 'nls.model.Model2D'>" (ref model.py:265); same field layout."""
    from scipy.io import loadmat, savemat
    monkeypatch.setattr(S, "nls", _OracleNative())
    m = M.Problem().model(model="2d", num_nodes=12, num_iters=2, order=3,
                          pumping=GaussianRingPumping2D(power=2.0, radius=0.4, variation=0.2))
    path = str(tmp_path / "own.mat")
    m.solve().store(path, label="t", desc="d")
    mat = {k: v for k, v in loadmat(path).items() if not k.startswith("__")}
    for ref_name, cls in (("<class 'nls.model.Model2D'>", M.Model2D), ("nls.model.Model2D", M.Model2D)):
        mat["model"] = ref_name
        ref_path = str(tmp_path / "ref.mat")
        savemat(ref_path, mat)
        again = M.Problem().model(filename=ref_path)
        assert isinstance(again, cls) and again.num_nodes == 12
        assert np.array_equal(again.getPumping(), m.getPumping())


def test_version():
    assert nls_b200.version() == (0, 2, 0)
