"""Pumping profiles P(x[, y]) sampled on the solver grid.

Host-side counterpart of the reference's ``nls/pumping.py``: the same public names, constructor parameters,
``setPower``, ``+`` / ``-`` combinators and call signatures, so that user scripts keep working.  The profiles feed
the hot path as the float64 array ``pumping`` of ``solve_nls*``, and ``BASELINE.json`` asks for the sampled profile to
be *bit-exact*: each formula below is therefore the numpy expression tree of the reference line it cites, in the same
association order -- including the reference's quirks:

* the root constructor ignores the ``power`` it is given (ref ``pumping.py:15-16``): combinators always scale by
  1.0 until ``setPower`` is called, and ``GaussianElipticPumping2D`` starts with amplitude 1.0;
* a ring is the *sum* of two full-power Gaussians centred at +R and -R (ref ``:156-159``).

Organisation (this file's own): the formulas are module-level functions, leaf profiles declare their parameters in a
``_signature`` table that one shared constructor binds, combinators carry the numpy operator they apply.

Parity: ``tests/test_pumping.py`` compares against golden arrays produced by the reference module itself
(``tests/golden/make_golden.py``).
"""

from __future__ import annotations

import operator

import numpy as np

__all__ = [
    "AbstractPumping", "OpSumPumping", "OpSubPumping", "OpMulPumping", "GridPumping",
    "GaussianPumping", "GaussianPumping1D", "GaussianPumping2D", "GaussianRingPumping1D",
    "GaussianRingPumping2D", "GaussianElipticPumping2D", "RectangularPumping1D",
    "RectangularRingPumping1D",
]


# ---- formulas (expression trees of the reference) ---------------------------------------------------------------
def _gauss(power, dx, dy, variation):
    # ref pumping.py:126
    return power * np.exp(-(dx ** 2 + dy ** 2) / (2.0 * variation ** 2))


def _radii(x, y, x0, y0):
    # ref pumping.py:174
    return np.sqrt((x - x0) ** 2 + (y - y0) ** 2)


def _ellipse(power, x, y, x0, y0, var, a, b):
    # ref pumping.py:194-199: width enters as 2 * var (not squared), polar radius measured from the grid origin
    angle = np.arctan2(x - x0, y - y0)
    re = np.sqrt((a * np.cos(angle)) ** 2 + (b * np.sin(angle)) ** 2)
    rp = np.sqrt(x ** 2 + y ** 2)
    return power * (np.exp(-(re - rp) ** 2 / (2 * var)) + np.exp(-(re + rp) ** 2 / (2 * var)))


def _tophat(power, x, x0, width):
    # ref pumping.py:213-217
    lo, hi = x0 - width / 2.0, x0 + width / 2.0
    out = np.zeros(x.shape)
    out[(x >= lo) & (x <= hi)] = power
    return out


# ---- the tree ---------------------------------------------------------------------------------------------------
class AbstractPumping(object):
    """Root of the pumping tree (ref ``pumping.py:11-40``)."""

    def __init__(self, power=1.0):
        self.power = 1.0            # ref :15-16 -- the argument is accepted and ignored

    def setPower(self, power):
        """Local multiplier of a combinator, physical power of a leaf (ref ``:18-22``)."""
        self.power = power

    def __call__(self, *args, **kwargs):
        raise Exception("Nothing to call: abstract class could not represent pumping.")

    def __add__(self, other):
        return OpSumPumping(self, other)

    def __sub__(self, other):
        return OpSubPumping(self, other)

    def __repr__(self):
        return type(self).__name__ if type(self) is AbstractPumping else "<class %s(AbstractPumping)>" % type(self).__name__

    def __str__(self):
        return repr(self)


class _Combination(AbstractPumping):
    """``power * (lhs <op> rhs)`` with the local multiplier starting at 1.0."""
    symbol, combine = "?", None

    def __init__(self, lhs, rhs):
        AbstractPumping.__init__(self)
        self.lhs, self.rhs = lhs, rhs

    def __call__(self, *args, **kwargs):
        if self.combine is None:
            raise Exception("Not supported yet!")
        return self.power * type(self).combine(self.lhs(*args, **kwargs), self.rhs(*args, **kwargs))

    def __repr__(self):
        return "%r %s %r" % (self.lhs, self.symbol, self.rhs)


class OpSumPumping(_Combination):
    """ref ``:43-57``"""
    symbol, combine = "+", operator.add


class OpSubPumping(_Combination):
    """ref ``:60-74``"""
    symbol, combine = "-", operator.sub


class OpMulPumping(_Combination):
    """Cartesian product of two 1D profiles -- unsupported in the reference too (ref ``:77-92``)."""
    symbol = "x"


class _Leaf(AbstractPumping):
    """A profile with named parameters: ``_signature`` = ((constructor name, default, attribute or None), ...);
    attribute None means the reference accepts the argument and drops it."""
    _signature = ()

    def __init__(self, *args, **kwargs):
        AbstractPumping.__init__(self)
        names = [entry[0] for entry in self._signature]
        if len(args) > len(names):
            raise TypeError("%s takes at most %d arguments" % (type(self).__name__, len(names)))
        given = dict(zip(names, args))
        for key, value in kwargs.items():
            if key not in names or key in given:
                raise TypeError("%s: unexpected or repeated argument %r" % (type(self).__name__, key))
            given[key] = value
        for name, default, attribute in self._signature:
            if attribute is not None:
                setattr(self, attribute, given.get(name, default))


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


class GaussianPumping(_Leaf):
    """Gaussian spot with origin, peak power and width (ref ``:113-131``)."""
    _signature = (("power", 1.0, "power"), ("x0", 0.0, "x0"), ("y0", 0.0, "y0"), ("variation", 5.0, "variation"))

    def __call__(self, x, y, t=None):
        return _gauss(self.power, x - self.x0, y - self.y0, self.variation)

    def __repr__(self):
        return "{0} exp(-{1} ((x - {2})^2 - (y - {3})^2))".format(
            self.power, 1.0 / (2.0 * self.variation ** 2), self.x0, self.y0)


class GaussianPumping1D(GaussianPumping):
    """Radial cut of :class:`GaussianPumping`: ``y`` is fixed to 0.0 (ref ``:134-142``)."""

    def __call__(self, x, t=None):
        return GaussianPumping.__call__(self, x, 0.0, t)


class GaussianPumping2D(GaussianPumping):
    """The same spot under its 2D name (ref ``:145-150``)."""


class GaussianRingPumping1D(OpSumPumping):
    """Ring of radius R in the radial model: G(x; +R) + G(x; -R) (ref ``:152-160``)."""

    def __init__(self, power=1.0, radius=0.0, variation=5.0):
        halves = [GaussianPumping1D(power, centre, 0.0, variation) for centre in (+radius, -radius)]
        OpSumPumping.__init__(self, *halves)


class GaussianRingPumping2D(_Leaf):
    """Ring with arbitrary centre on the Cartesian grid (ref ``:163-178``)."""
    _signature = (("power", 1.0, None), ("x0", 0.0, "x0"), ("y0", 0.0, "y0"), ("variation", 5.0, None),
                  ("radius", 1.0, None))

    def __init__(self, power=1.0, x0=0.0, y0=0.0, variation=5.0, radius=1.0):
        _Leaf.__init__(self, power, x0, y0, variation, radius)
        self.pumping = GaussianRingPumping1D(power, radius, variation)

    def __call__(self, x, y, t=None):
        return self.pumping(_radii(x, y, self.x0, self.y0), t)


class GaussianElipticPumping2D(_Leaf):
    """Ring replaced by an ellipse with semi-axes a, b (ref ``:181-202``); the ``power`` argument is dropped by the
    root constructor, so the amplitude is 1.0 until ``setPower`` is called."""
    _signature = (("power", 1.0, None), ("x0", 0.0, "x0"), ("y0", 0.0, "y0"), ("variation", 5.0, "var"),
                  ("a", 1.0, "a"), ("b", 1.0, "b"))

    def __call__(self, x, y, t=None):
        return _ellipse(self.power, x, y, self.x0, self.y0, self.var, self.a, self.b)


class RectangularPumping1D(_Leaf):
    """Top-hat of given width centred at x0 (ref ``:205-220``)."""
    _signature = (("power", 10.0, "power"), ("x0", 0.0, "x0"), ("width", 1.0, "width"))

    def __call__(self, x, t=None):
        return _tophat(self.power, x, self.x0, self.width)

    def __repr__(self):
        return "{0} * (\\theta(x - {1}) - \\theta({2} - x))".format(
            self.power, self.x0 + self.width / 2.0, self.x0 - self.width / 2.0)


class RectangularRingPumping1D(OpSumPumping):
    """Two top-hats at +R and -R (ref ``:223-229``)."""

    def __init__(self, power=10.0, radius=10.0, width=2.0):
        halves = [RectangularPumping1D(power, centre, width) for centre in (+radius, -radius)]
        OpSumPumping.__init__(self, *halves)
