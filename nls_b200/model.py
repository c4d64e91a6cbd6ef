"""Problem / Model / Solution -- the object API in front of the B200 engine.

Python-3 mirror of the reference's ``nls/model.py``: ``Problem().model(**kwargs)`` builds a
``Model1D`` or ``Model2D``; ``model.solve()`` runs the engine through ``nls_b200.solver`` and returns a
``Solution`` with ``report()``, ``getSolution()``, ``getDensity()``, ``getReservoir()``, ...

The parts that feed the hot path reproduce the reference's float64 arithmetic exactly:

* the non-dimensional coefficient vector ``coeffs[23]`` (ref ``model.py:138-160``),
* the pumping grid -- ``linspace(0, n*dx, n)`` in 1D, ``linspace(-n*dx/2, n*dx/2, n)`` + ``meshgrid``
  in 2D, i.e. a spacing of ``n*dx/(n-1)`` and not ``dx`` (ref ``model.py:220-232``),
* the scalar-``u0`` broadcast (ref ``:96-97``, ``:114-115``).

Plotting and ``.mat`` I/O import matplotlib / scipy.io lazily: they are outside the hot path.
"""

from __future__ import annotations

import copy
from datetime import datetime
from types import FunctionType

import numpy as np

from .pumping import GaussianPumping, GaussianPumping1D, GridPumping
from .version import version

__all__ = ["Problem", "AbstractModel", "Model1D", "Model2D", "Solution", "DEFAULT_ORIGINAL_PARAMS",
           "dimensionless_coefficients"]

# ref model.py:56-65
DEFAULT_ORIGINAL_PARAMS = {
    "R": 0.0242057488654,
    "gamma": 0.0242057488654,
    "g": 0.00162178517398,
    "tilde_g": 0.0169440242057,
    "gamma_R": 0.242057488654,
}

_HBAR = 6.61e-34
_M_E = 9.1e-31
_M_0 = 1.0e-5 * _M_E


def _scales(originals):
    # ref model.py:142-145 (and :179-182): phi0 = t0 = sqrt(gamma / 2g), x0, n0
    phi0 = np.sqrt(originals["gamma"] / (2.0 * originals["g"]))
    t0 = phi0
    x0 = np.sqrt(_HBAR * t0 / (2 * _M_0))
    n0 = 2.0 / (originals["R"] * t0)
    return {"x": x0, "t": t0, "n": n0, "phi": phi0}


def dimensionless_coefficients(originals):
    """The 23-entry coefficient vector of ``solve_nls`` (ref ``model.py:147-160``, doc ``nls.f90:770-786``).

    Only entries 2..5 and 11..13 are read by the time stepper.
    """
    s = _scales(originals)
    c = np.zeros(23)
    c[0] = 1.0
    c[1] = 1.0
    c[2] = 1.0
    c[3] = 1.0
    c[4] = 1.0
    c[5] = 4.0 * originals["tilde_g"] / originals["R"]
    c[10] = 0.0
    c[11] = 1.0 / (s["n"] * originals["gamma_R"])
    c[12] = 1.0
    c[13] = originals["R"] * s["phi"] ** 2 / originals["gamma_R"]
    c[14] = 0.0
    return c


class Problem(object):
    """Factory of models (ref ``model.py:27-117``)."""

    def model(self, *args, **kwargs):
        if "filename" in kwargs:
            return self.modelFromFile(kwargs["filename"])
        kwargs.pop("params", None)

        kwargs.setdefault("model", "default")
        originals = dict(kwargs.get("original_params") or {})
        for key, value in DEFAULT_ORIGINAL_PARAMS.items():
            originals.setdefault(key, value)
        kwargs["original_params"] = originals

        # .mat files store str(type(model)) (ref model.py:265): accept this package's class strings and the
        # reference's ("<class 'nls.model.Model1D'>") alike, so that the reference's own checkpoints restore
        kind = str(kwargs["model"]).strip()
        if kind in ("1d", "default") or kind.rstrip("'>").endswith("Model1D"):
            return self.fabricateModel1D(*args, **kwargs)
        if kind == "2d" or kind.rstrip("'>").endswith("Model2D"):
            return self.fabricateModel2D(*args, **kwargs)
        raise Exception("Unknown model passed!")

    def modelFromFile(self, filename):
        from scipy.io import loadmat
        mat = loadmat(filename)
        if "model" not in mat:
            return None
        return self.model(model=str(mat["model"][0])).restore(filename)

    @staticmethod
    def _defaults(kwargs, order, num_nodes, num_iters, pumping):
        kwargs.setdefault("dx", 1.0e-1)
        kwargs.setdefault("dt", 1.0e-3)
        kwargs.setdefault("t0", 0.0e+0)
        kwargs.setdefault("u0", 1.0e-1)
        kwargs.setdefault("order", order)
        kwargs.setdefault("num_nodes", num_nodes)
        kwargs.setdefault("num_iters", num_iters)
        if kwargs.get("pumping") is None:
            kwargs["pumping"] = pumping

    def fabricateModel1D(self, *args, **kwargs):
        # ref :86-102.  Divergence: the reference's default pumping is the 2D
        # ``GaussianPumping()`` whose call needs ``y`` and raises TypeError in 1D; the 1D facade with
        # the same parameters is used instead so that ``Problem().model().solve()`` runs.
        self._defaults(kwargs, 5, 1000, 100000, GaussianPumping1D())
        n = kwargs["num_nodes"]
        if type(kwargs["u0"]) in (int, float, complex):
            kwargs["u0"] = kwargs["u0"] * np.ones(n)
        elif isinstance(kwargs["u0"], FunctionType):
            grid = np.linspace(0.0, kwargs["dx"] * n, n)
            kwargs["u0"] = kwargs["u0"](grid)
        return Model1D(**kwargs)

    def fabricateModel2D(self, *args, **kwargs):
        # ref :104-117
        self._defaults(kwargs, 3, 40, 1000, GaussianPumping())
        n = kwargs["num_nodes"]
        if type(kwargs["u0"]) in (int, float, complex):
            kwargs["u0"] = kwargs["u0"] * np.ones((n, n))
        return Model2D(**kwargs)


class AbstractModel(object):
    """State of one problem instance plus the getters the solver facade reads (ref ``model.py:120-314``)."""

    def __init__(self, *args, **kwargs):
        self.dt = kwargs["dt"]
        self.dx = kwargs["dx"]
        self.order = kwargs["order"]
        self.num_nodes = kwargs["num_nodes"]
        self.num_iters = kwargs["num_iters"]
        self.pumping = kwargs["pumping"]
        self.init_sol = np.asarray(kwargs["u0"])
        self.originals = kwargs["original_params"]
        self.coeffs = dimensionless_coefficients(self.originals)
        self.verbose = bool(kwargs.get("verbose"))
        self.solver = None

    def __repr__(self):
        from pprint import pformat
        return pformat({
            "dt": self.dt, "dx": self.dx, "order": self.order, "num_nodes": self.num_nodes,
            "num_iters": self.num_iters, "pumping": self.pumping, "originals": self.originals,
        }) + "\n" + str(self.coeffs)

    # -- getters read by the solver facade (ref solver.py:27-34) ----------------------------------
    def getApproximationOrder(self):
        return self.order

    def getCharacteristicScale(self, scale):
        return _scales(self.originals).get(scale)

    def getChemicalPotential(self, solution):
        if isinstance(solution, Solution):
            solution = solution.getSolution()
        self.mu = self.solver.chemicalPotential(solution)
        return self.mu

    def getCoefficients(self):
        return self.coeffs

    def getInitialSolution(self):
        return self.init_sol

    def getNumberOfIterations(self):
        return self.num_iters

    def getNumberOfNodes(self):
        return self.num_nodes

    def getPumping(self):
        """Sample the pumping functor on the reference's grid (ref ``model.py:220-232``)."""
        n = self.num_nodes
        if self.init_sol.ndim == 1:
            x = np.linspace(0.0, n * self.dx, n)
            return self.pumping(*np.meshgrid(x))
        right = n * self.dx / 2
        x = np.linspace(-right, right, n)
        return self.pumping(*np.meshgrid(x, x))

    def getSpatialStep(self):
        return self.dx

    def getSolver(self):
        return self.solver

    def getTimeStep(self):
        return self.dt

    def setNumberOfIterations(self, num_iters):
        self.num_iters = num_iters

    def setPumping(self, pumping):
        self.pumping = pumping

    def setInitialSolution(self, solution):
        self.init_sol = np.asarray(solution)

    def solve(self, num_iters=None):
        return self.solver(num_iters)

    # -- .mat persistence (ref model.py:257-314); outside the hot path ----------------------------
    def _as_matfile(self, date):
        return {
            "model": str(type(self)),
            "date": date,
            "dim": self.init_sol.ndim,
            "dimlesses": self.coeffs,
            "init_solution": self.init_sol,
            "num_iters": self.num_iters,
            "num_nodes": self.num_nodes,
            "order": self.order,
            "originals": self.originals,
            "pumping": self.getPumping(),
            "spatial_step": self.dx,
            "time_step": self.dt,
        }

    def store(self, filename=None, label=None, desc=None, date=None):
        from scipy.io import savemat
        date = (date if date else datetime.now()).replace(microsecond=0).isoformat()
        matfile = self._as_matfile(date)
        if desc:
            matfile["desc"] = desc
        if label:
            matfile["label"] = label
        savemat(filename if filename else date + ".mat", matfile)

    def restore(self, filename):
        from scipy.io import loadmat
        mat = loadmat(filename)
        one_d = int(mat["dim"][0, 0]) == 1
        self.coeffs = np.array(mat["dimlesses"][0, :], dtype=float)
        self.init_sol = mat["init_solution"][0, :] if one_d else mat["init_solution"]
        self.num_nodes = int(mat["num_nodes"][0, 0])
        self.num_iters = int(mat["num_iters"][0, 0])
        self.order = int(mat["order"][0, 0])
        self.pumping = GridPumping(mat["pumping"][0, :] if one_d else mat["pumping"])
        self.dx = float(mat["spatial_step"][0, 0])
        self.dt = float(mat["time_step"][0, 0])
        record = mat["originals"][0, 0]
        self.originals = {name: float(record[name][0, 0]) for name in record.dtype.names}
        if "desc" in mat:
            self.desc = str(mat["desc"][0])
        if "label" in mat:
            self.label = str(mat["label"][0])
        return self


class Model1D(AbstractModel):
    """NLS with reservoir, axially symmetric (radial) geometry (ref ``model.py:317-324``)."""

    def __init__(self, *args, **kwargs):
        AbstractModel.__init__(self, *args, **kwargs)
        from .solver import Solver1D
        self.solver = Solver1D(self)


class Model2D(AbstractModel):
    """NLS with reservoir on an n x n Cartesian grid (ref ``model.py:327-334``)."""

    def __init__(self, *args, **kwargs):
        AbstractModel.__init__(self, *args, **kwargs)
        from .solver import Solver2D
        self.solver = Solver2D(self)


class Solution(object):
    """Result of ``model.solve()`` with the reference's diagnostics (ref ``model.py:337-536``)."""

    def __init__(self, model, solution=None, verbose=False):
        self.elapsed_time = 0.0
        self.model = model
        self.solution = solution
        self.verbose = verbose

    def getDampingIntegral(self):
        # ref :350-365 (rectangle rule; builtin sum in the reference, same value up to rounding)
        reservoir, density = self.getReservoir(), self.getDensity()
        length = self.model.getSpatialStep()
        if self.solution.ndim == 1:
            nodes = self.model.getNumberOfNodes()
            radius = np.linspace(0, nodes * length, nodes)
            return 2 * np.pi * np.sum((reservoir - 1.0) * density * radius * length)
        return np.sum((reservoir - 1.0) * density * length ** 2)

    def getDensity(self):
        return (self.solution.conj() * self.solution).real

    def getElapsedTime(self):
        return self.elapsed_time

    def getModel(self):
        return self.model

    def getReservoir(self):
        # ref :376-380 -- the algebraic reservoir n = c11 P / (c12 + c13 |psi|^2)
        c = self.model.coeffs
        return c[11] * self.model.getPumping() / (c[12] + c[13] * self.getDensity())

    def getSolution(self):
        return self.solution

    def setElapsedTime(self, seconds):
        self.elapsed_time = seconds

    def setSolution(self, solution):
        self.solution = solution

    def visualize(self, *args, **kwargs):
        # Plotting (ref :391-495) is outside the engine's scope; kept as a thin optional hook.
        try:
            import matplotlib.pyplot as plt
        except ImportError as exc:  # pragma: no cover - matplotlib is not in the build image
            raise RuntimeError("Solution.visualize needs matplotlib, which is not installed") from exc
        density, reservoir, pumping = self.getDensity(), self.getReservoir(), self.model.getPumping()
        fig = kwargs.get("figure") or plt.figure()
        panels = (("Pumping profile.", pumping), ("Density distribution of BEC.", density),
                  ("Density distribution of reservoir.", reservoir))
        for idx, (name, value) in enumerate(panels, 1):
            ax = fig.add_subplot(1, 3, idx)
            if value.ndim == 1:
                ax.plot(np.arange(0.0, self.model.dx * self.model.num_nodes, self.model.dx)[:value.size], value)
            else:
                ax.imshow(value)
            ax.set_title(name)
        if kwargs.get("filename"):
            fig.savefig(kwargs["filename"])
        return fig

    def show(self):
        import matplotlib.pyplot as plt
        plt.show()

    def store(self, filename=None, label=None, desc=None, date=None):
        from scipy.io import savemat
        date = datetime.now() if date is None else date
        stamp = date.replace(microsecond=0).isoformat()
        content = self.model._as_matfile(stamp)
        if desc:
            content["desc"] = desc
        if label:
            content["label"] = label
        content.update({"elapsed_time": self.elapsed_time, "solution": self.solution, "version": version()})
        savemat(filename if filename else stamp + ".mat", content)

    def restore(self, filename):
        from scipy.io import loadmat
        mat = loadmat(filename)
        self.elapsed_time = float(mat["elapsed_time"][0, 0])
        self.solution = mat["solution"][0, :] if int(mat["dim"][0, 0]) == 1 else mat["solution"]
        return self

    def report(self):
        message = "Elapsed in {0} seconds with {1} iteration on {2} grid nodes."
        print(message.format(self.elapsed_time, self.model.getNumberOfIterations(), self.model.getNumberOfNodes()))

    def copy(self):
        return copy.copy(self)
