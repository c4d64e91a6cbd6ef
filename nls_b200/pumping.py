"""Pumping profiles P(x[, y]) sampled on the solver grid.

Host-side mirror of the reference's ``nls/pumping.py`` (class names, constructor arguments,
``setPower``, ``+``/``-`` combinators and call signatures are the reference's).  The profiles feed
the hot path as the float64 array ``pumping`` of ``solve_nls*``; ``BASELINE.json`` asks for the
sampled profile to be *bit-exact*, so every ``__call__`` below evaluates the same numpy expression
tree, in the same association order, as the reference line it cites -- including the reference's
quirks:

* the base-class constructor drops its ``power`` argument (ref ``pumping.py:15-16``): combinators
  always scale by 1.0 and ``GaussianElipticPumping2D`` ignores the ``power`` it is given;
* a ring is the *sum* of two full-power Gaussians centred at +R and -R (ref ``:156-159``).

Parity: ``tests/test_pumping.py`` compares against golden arrays produced by the reference module
itself (``tests/golden/make_golden.py``).
"""

from __future__ import annotations

import numpy as np

__all__ = [
    "AbstractPumping", "OpSumPumping", "OpSubPumping", "OpMulPumping", "GridPumping",
    "GaussianPumping", "GaussianPumping1D", "GaussianPumping2D", "GaussianRingPumping1D",
    "GaussianRingPumping2D", "GaussianElipticPumping2D", "RectangularPumping1D",
    "RectangularRingPumping1D",
]


def _gauss(power, dx, dy, variation):
    # ref pumping.py:126 -- power * exp(-((x-x0)**2 + (y-y0)**2) / (2.0 * variation**2))
    return power * np.exp(-(dx ** 2 + dy ** 2) / (2.0 * variation ** 2))


class AbstractPumping(object):
    """Root of the pumping tree (ref ``pumping.py:11-40``)."""

    def __init__(self, power=1.0):
        # ref :15-16 -- the argument is accepted and ignored; the local multiplier starts at 1.0
        self.power = 1.0

    def setPower(self, power):
        """Local multiplier for combinators, physical power for leaf profiles (ref ``:18-22``)."""
        self.power = power

    def __call__(self, *args, **kwargs):
        raise Exception("Nothing to call: abstract class could not represent pumping.")

    def __add__(self, other):
        return OpSumPumping(self, other)

    def __sub__(self, other):
        return OpSubPumping(self, other)

    def __repr__(self):
        return "AbstractPumping"

    __str__ = __repr__


class _BinaryPumping(AbstractPumping):
    symbol = "?"

    def __init__(self, lhs, rhs):
        AbstractPumping.__init__(self)
        self.lhs, self.rhs = lhs, rhs

    def __repr__(self):
        return "%r %s %r" % (self.lhs, self.symbol, self.rhs)

    __str__ = __repr__


class OpSumPumping(_BinaryPumping):
    """``power * (lhs + rhs)`` (ref ``:43-57``)."""
    symbol = "+"

    def __call__(self, *args, **kwargs):
        return self.power * (self.lhs(*args, **kwargs) + self.rhs(*args, **kwargs))


class OpSubPumping(_BinaryPumping):
    """``power * (lhs - rhs)`` (ref ``:60-74``)."""
    symbol = "-"

    def __call__(self, *args, **kwargs):
        return self.power * (self.lhs(*args, **kwargs) - self.rhs(*args, **kwargs))


class OpMulPumping(_BinaryPumping):
    """Cartesian product of two 1D profiles -- unsupported in the reference too (ref ``:77-92``)."""
    symbol = "x"

    def __call__(self, *args, **kwargs):
        raise Exception("Not supported yet!")


class GridPumping(AbstractPumping):
    """A profile that is already sampled (custom, or restored from a ``.mat`` file; ref ``:95-110``)."""

    def __init__(self, pumping, desciption=None):
        AbstractPumping.__init__(self)
        self.pumping = pumping
        self.desciption = desciption if desciption else "<class GridPumping>"

    def __call__(self, *args, **kwargs):
        return self.pumping

    def __repr__(self):
        return self.desciption

    __str__ = __repr__


class GaussianPumping(AbstractPumping):
    """Gaussian spot with origin, peak power and width (ref ``:113-131``)."""

    def __init__(self, power=1.0, x0=0.0, y0=0.0, variation=5.0):
        AbstractPumping.__init__(self)
        self.power, self.x0, self.y0, self.variation = power, x0, y0, variation

    def __call__(self, x, y, t=None):
        return _gauss(self.power, x - self.x0, y - self.y0, self.variation)

    def __repr__(self):
        return "{0} exp(-{1} ((x - {2})^2 - (y - {3})^2))".format(
            self.power, 1.0 / (2.0 * self.variation ** 2), self.x0, self.y0)

    __str__ = __repr__


class GaussianPumping1D(GaussianPumping):
    """Radial cut of :class:`GaussianPumping`: ``y`` is fixed to 0.0 (ref ``:134-142``)."""

    def __call__(self, x, t=None):
        return GaussianPumping.__call__(self, x, 0.0, t)


class GaussianPumping2D(GaussianPumping):
    """Alias of :class:`GaussianPumping` for the 2D model (ref ``:145-150``)."""


class GaussianRingPumping1D(OpSumPumping):
    """Ring of radius R in the radial model: G(x; +R) + G(x; -R) (ref ``:152-160``)."""

    def __init__(self, power=1.0, radius=0.0, variation=5.0):
        OpSumPumping.__init__(self,
                              GaussianPumping1D(power, +radius, 0.0, variation),
                              GaussianPumping1D(power, -radius, 0.0, variation))


class GaussianRingPumping2D(AbstractPumping):
    """Ring with arbitrary centre on the Cartesian grid (ref ``:163-178``)."""

    def __init__(self, power=1.0, x0=0.0, y0=0.0, variation=5.0, radius=1.0):
        AbstractPumping.__init__(self)
        self.x0, self.y0 = x0, y0
        self.pumping = GaussianRingPumping1D(power, radius, variation)

    def __call__(self, x, y, t=None):
        radii = np.sqrt((x - self.x0) ** 2 + (y - self.y0) ** 2)
        return self.pumping(radii, t)

    def __repr__(self):
        return "<class GaussianRingPumping2D(AbstractPumping)>"

    __str__ = __repr__


class GaussianElipticPumping2D(AbstractPumping):
    """Ring replaced by an ellipse with semi-axes a, b (ref ``:181-202``).

    Note the reference's own arithmetic: the width enters as ``2 * var`` (not squared), the polar
    radius is measured from the grid origin (not from ``x0, y0``), and ``power`` is dropped by the
    base constructor, so the amplitude is 1.0 until ``setPower`` is called.
    """

    def __init__(self, power=1.0, x0=0.0, y0=0.0, variation=5.0, a=1.0, b=1.0):
        AbstractPumping.__init__(self, power)
        self.x0, self.y0, self.var, self.a, self.b = x0, y0, variation, a, b

    def __call__(self, x, y, t=None):
        t = np.arctan2(x - self.x0, y - self.y0)
        re = np.sqrt((self.a * np.cos(t)) ** 2 + (self.b * np.sin(t)) ** 2)
        rp = np.sqrt(x ** 2 + y ** 2)
        return self.power * (np.exp(-(re - rp) ** 2 / (2 * self.var)) +
                             np.exp(-(re + rp) ** 2 / (2 * self.var)))

    def __repr__(self):
        return "<class GaussianElipticPumping2D(AbstractPumping)>"

    __str__ = __repr__


class RectangularPumping1D(AbstractPumping):
    """Top-hat of given width centred at x0 (ref ``:205-220``)."""

    def __init__(self, power=10.0, x0=0.0, width=1.0):
        AbstractPumping.__init__(self, power)
        self.x0, self.width, self.power = x0, width, power

    def __call__(self, x, t=None):
        lo, hi = self.x0 - self.width / 2.0, self.x0 + self.width / 2.0
        out = np.zeros(x.shape)
        out[(x >= lo) & (x <= hi)] = self.power
        return out

    def __repr__(self):
        return "{0} * (\\theta(x - {1}) - \\theta({2} - x))".format(
            self.power, self.x0 + self.width / 2.0, self.x0 - self.width / 2.0)

    __str__ = __repr__


class RectangularRingPumping1D(OpSumPumping):
    """Two top-hats at +R and -R (ref ``:223-229``)."""

    def __init__(self, power=10.0, radius=10.0, width=2.0):
        OpSumPumping.__init__(self,
                              RectangularPumping1D(power, +radius, width),
                              RectangularPumping1D(power, -radius, width))
