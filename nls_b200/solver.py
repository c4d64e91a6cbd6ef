"""Solver facade between a model and the native module.

Keeps the public surface of the reference's ``nls/solver.py`` -- ``AbstractSolver`` callable with an optional step
count, ``solve`` / ``chemicalPotential`` / ``chemicalPotentialRoutine``, ``Solver1D`` and ``Solver2D`` -- and its
calling convention: the model's getters are passed positionally as ``(dt, dx, order, iters, pumping, coeffs, u0)``
to ``nls.solve_nls`` / ``nls.solve_nls_2d`` (ref ``solver.py:21-38``, ``:65-66``, ``:79-80``) and as
``(dx, pumping, coeffs, solution)`` to ``nls.chemical_potential_1d/_2d`` (ref ``:45-50``, ``:68-69``, ``:82-83``).
Here ``nls`` is ``nls_b200.native.nls``: the CUDA engine behind the f2py signatures.  The two concrete solvers only
name the native routines they dispatch to.
"""

from __future__ import annotations

import time

from .native import nls

__all__ = ["AbstractSolver", "Solver1D", "Solver2D"]

# getters of the model, in the positional order of the native entry points
_SOLVE_ARGUMENTS = ("getTimeStep", "getSpatialStep", "getApproximationOrder", "getNumberOfIterations", "getPumping",
                    "getCoefficients", "getInitialSolution")
_POTENTIAL_ARGUMENTS = ("getSpatialStep", "getPumping", "getCoefficients")


class AbstractSolver(object):
    solve_routine = None        # names of the attributes of ``nls`` a concrete solver dispatches to
    potential_routine = None

    def __init__(self, model):
        self.model, self.elapsed_time, self.solution = model, 0.0, None

    def _native(self, name):
        if name is None:
            raise Exception("AbstractSolver: native solver routine is not passed!")
        return getattr(nls, name)

    def __call__(self, num_iters=None):
        from .model import Solution
        if num_iters:
            self.model.setNumberOfIterations(num_iters)
        began = time.time()
        result = Solution(self.model)
        result.setSolution(self.solve(*[getattr(self.model, getter)() for getter in _SOLVE_ARGUMENTS]))
        self.elapsed_time = time.time() - began
        result.setElapsedTime(self.elapsed_time)
        self.solution = result
        return result

    def solve(self, *args, **kwargs):
        return self._native(self.solve_routine)(*args, **kwargs)

    def chemicalPotential(self, solution, *args, **kwargs):
        leading = [getattr(self.model, getter)() for getter in _POTENTIAL_ARGUMENTS]
        return self.chemicalPotentialRoutine(*(leading + [solution]))

    def chemicalPotentialRoutine(self, *args, **kwargs):
        return self._native(self.potential_routine)(*args, **kwargs)


class Solver1D(AbstractSolver):
    """Radial (axially symmetric) problems (ref ``solver.py:58-69``)."""
    solve_routine, potential_routine = "solve_nls", "chemical_potential_1d"


class Solver2D(AbstractSolver):
    """Problems on the square grid (ref ``solver.py:72-83``)."""
    solve_routine, potential_routine = "solve_nls_2d", "chemical_potential_2d"
