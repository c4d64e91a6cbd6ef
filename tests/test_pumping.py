"""Bit-exactness of the pumping profiles and grids against arrays produced by the REFERENCE module
(tests/golden/make_golden.py imports /root/reference/nls/pumping.py; only the .npz is read here)."""

import os

import numpy as np
import pytest

from nls_b200 import pumping as P
from nls_b200.model import Problem

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "pumping_golden.npz"))

CASES = [
    ("ring1d_c1", P.GaussianRingPumping1D, dict(power=20.0, radius=10.0, variation=3.14), "1d", 400, 0.1),
    ("ring1d_c3", P.GaussianRingPumping1D, dict(power=7.5, radius=10.0, variation=3.14), "1d", 1000, 0.1),
    ("gauss1d", P.GaussianPumping1D, dict(power=3.0, x0=1.5, variation=2.5), "1d", 257, 0.07),
    ("rect1d", P.RectangularPumping1D, dict(power=10.0, x0=5.0, width=3.0), "1d", 200, 0.1),
    ("rectring1d", P.RectangularRingPumping1D, dict(power=4.0, radius=6.0, width=2.0), "1d", 200, 0.1),
    ("ring2d_c2_96", P.GaussianRingPumping2D, dict(power=20.0, radius=10.0, variation=3.14), "2d", 96, 0.1),
    ("ring2d_offset", P.GaussianRingPumping2D, dict(power=5.0, x0=0.7, y0=-1.1, variation=1.3, radius=2.0), "2d", 65, 0.13),
    ("gauss2d_bench", P.GaussianPumping2D, dict(power=15.0, variation=3.14), "2d", 50, 0.2),
    ("eliptic2d", P.GaussianElipticPumping2D, dict(power=9.0, x0=0.2, y0=0.1, variation=2.0, a=3.0, b=1.5), "2d", 48, 0.2),
]


@pytest.mark.parametrize("key,cls,kwargs,dim,n,dx", CASES, ids=[c[0] for c in CASES])
def test_profile_on_model_grid_is_bit_exact(key, cls, kwargs, dim, n, dx):
    # goes through Problem().model(...).getPumping(): exercises the grid construction of model.py:220-232 too
    model = Problem().model(model=dim, dx=dx, num_nodes=n, pumping=cls(**kwargs))
    got = model.getPumping()
    assert got.dtype == np.float64 and got.shape == GOLDEN[key].shape
    assert np.array_equal(got, GOLDEN[key])


def test_combinators_are_bit_exact():
    a = P.GaussianPumping1D(power=2.0, x0=1.0, variation=1.5)
    b = P.GaussianPumping1D(power=0.5, x0=4.0, variation=0.7)
    x = np.meshgrid(np.linspace(0.0, 120 * 0.1, 120))
    assert np.array_equal((a + b)(*x), GOLDEN["sum1d"])
    assert np.array_equal((a - b)(*x), GOLDEN["sub1d"])
    s = a + b
    s.setPower(3.0)
    assert np.array_equal(s(*x), GOLDEN["sum1d_power3"])


def test_reference_quirks_are_kept():
    assert P.AbstractPumping(power=7.0).power == 1.0                 # ref pumping.py:15-16
    assert P.GaussianElipticPumping2D(power=9.0).power == 1.0        # power dropped by the base ctor
    assert P.RectangularPumping1D(power=9.0).power == 9.0            # set explicitly, ref :211
    with pytest.raises(Exception):
        P.OpMulPumping(P.GaussianPumping1D(), P.GaussianPumping1D())(np.zeros(3))
    grid = P.GridPumping(np.arange(4.0))
    assert np.array_equal(grid(np.zeros(9)), np.arange(4.0))
    # 2D default pumping called with one argument fails exactly as in the reference (App. B #7)
    with pytest.raises(TypeError):
        P.GaussianPumping()(np.zeros(3))
