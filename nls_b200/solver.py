"""Solver facade: marshals a model's getters into the native call and times it.

Mirror of the reference's ``nls/solver.py``: ``AbstractSolver.__call__`` passes
``(dt, dx, order, iters, pumping, coeffs, u0)`` positionally to ``nls.solve_nls`` /
``nls.solve_nls_2d`` (ref ``solver.py:21-38``, ``:65-66``, ``:79-80``) and
``(dx, pumping, coeffs, solution)`` to ``nls.chemical_potential_1d/_2d`` (ref ``:45-50``, ``:68-69``,
``:82-83``).  Here ``nls`` is ``nls_b200.native.nls`` -- the CUDA engine behind the f2py signatures.
"""

from __future__ import annotations

from time import time

from .native import nls

__all__ = ["AbstractSolver", "Solver1D", "Solver2D"]


class AbstractSolver(object):
    def __init__(self, model):
        self.model = model
        self.elapsed_time = 0.0
        self.solution = None

    def __call__(self, num_iters=None):
        from .model import Solution
        model = self.model
        if num_iters:
            model.setNumberOfIterations(num_iters)
        started = time()
        self.solution = Solution(model)
        self.solution.setSolution(self.solve(
            model.getTimeStep(),
            model.getSpatialStep(),
            model.getApproximationOrder(),
            model.getNumberOfIterations(),
            model.getPumping(),
            model.getCoefficients(),
            model.getInitialSolution()))
        self.elapsed_time = time() - started
        self.solution.setElapsedTime(self.elapsed_time)
        return self.solution

    def solve(self, *args, **kwargs):
        raise Exception("AbstractSolver: native solver routine is not passed!")

    def chemicalPotential(self, solution, *args, **kwargs):
        model = self.model
        return self.chemicalPotentialRoutine(
            model.getSpatialStep(), model.getPumping(), model.getCoefficients(), solution)

    def chemicalPotentialRoutine(self, *args, **kwargs):
        raise Exception("AbstractSolver: native solver routine is not passed!")


class Solver1D(AbstractSolver):
    """Radial (axially symmetric) solver -> ``solve_nls`` (ref ``solver.py:58-69``)."""

    def solve(self, *args, **kwargs):
        return nls.solve_nls(*args, **kwargs)

    def chemicalPotentialRoutine(self, *args, **kwargs):
        return nls.chemical_potential_1d(*args, **kwargs)


class Solver2D(AbstractSolver):
    """Square-grid solver -> ``solve_nls_2d`` (ref ``solver.py:72-83``)."""

    def solve(self, *args, **kwargs):
        return nls.solve_nls_2d(*args, **kwargs)

    def chemicalPotentialRoutine(self, *args, **kwargs):
        return nls.chemical_potential_2d(*args, **kwargs)
