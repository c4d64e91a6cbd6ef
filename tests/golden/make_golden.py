#!/usr/bin/env python
"""Generate golden pumping profiles from the REFERENCE's own ``nls/pumping.py``.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

``nls/pumping.py`` is the only part of the reference that imports under Python 3.12 (SURVEY.md
finding 2); it is loaded by file path, evaluated on the grids of ``nls/model.py:220-232`` and the
float64 results are stored in ``tests/golden/pumping_golden.npz``.  The GPU box has no
/root/reference: tests read only the committed .npz.
"""

import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("NLS_REFERENCE", "/root/reference")


def load_reference_pumping():
    spec = importlib.util.spec_from_file_location("ref_pumping", os.path.join(REFERENCE, "nls", "pumping.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def grid_1d(n, dx):
    # ref model.py:221-226
    x = np.linspace(0.0, n * dx, n)
    return np.meshgrid(x)


def grid_2d(n, dx):
    # ref model.py:228-232
    right = n * dx / 2
    x = np.linspace(-right, right, n)
    return np.meshgrid(x, x)


# (key, class name, constructor kwargs, dim, n, dx[, power set afterwards])
CASES = [
    ("ring1d_c1", "GaussianRingPumping1D", dict(power=20.0, radius=10.0, variation=3.14), 1, 400, 0.1),
    ("ring1d_c3", "GaussianRingPumping1D", dict(power=7.5, radius=10.0, variation=3.14), 1, 1000, 0.1),
    ("gauss1d", "GaussianPumping1D", dict(power=3.0, x0=1.5, variation=2.5), 1, 257, 0.07),
    ("rect1d", "RectangularPumping1D", dict(power=10.0, x0=5.0, width=3.0), 1, 200, 0.1),
    ("rectring1d", "RectangularRingPumping1D", dict(power=4.0, radius=6.0, width=2.0), 1, 200, 0.1),
    ("ring2d_c2_96", "GaussianRingPumping2D", dict(power=20.0, radius=10.0, variation=3.14), 2, 96, 0.1),
    ("ring2d_offset", "GaussianRingPumping2D", dict(power=5.0, x0=0.7, y0=-1.1, variation=1.3, radius=2.0), 2, 65, 0.13),
    ("gauss2d_bench", "GaussianPumping2D", dict(power=15.0, variation=3.14), 2, 50, 0.2),
    ("eliptic2d", "GaussianElipticPumping2D", dict(power=9.0, x0=0.2, y0=0.1, variation=2.0, a=3.0, b=1.5), 2, 48, 0.2),
]


def main():
    ref = load_reference_pumping()
    out = {}
    for key, cls, kwargs, dim, n, dx in CASES:
        functor = getattr(ref, cls)(**kwargs)
        grid = grid_1d(n, dx) if dim == 1 else grid_2d(n, dx)
        out[key] = np.asarray(functor(*grid), dtype=np.float64)
    # combinators: sum and difference of leaf profiles, and setPower on a combinator
    a = ref.GaussianPumping1D(power=2.0, x0=1.0, variation=1.5)
    b = ref.GaussianPumping1D(power=0.5, x0=4.0, variation=0.7)
    g = grid_1d(120, 0.1)
    out["sum1d"] = (a + b)(*g)
    out["sub1d"] = (a - b)(*g)
    s = a + b
    s.setPower(3.0)
    out["sum1d_power3"] = s(*g)
    np.savez_compressed(os.path.join(HERE, "pumping_golden.npz"), **out)
    print("wrote %d arrays" % len(out))


if __name__ == "__main__":
    main()
