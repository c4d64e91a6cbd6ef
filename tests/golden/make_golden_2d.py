#!/usr/bin/env python
"""Generate compact golden fixtures of LONG / LARGE 2D solves with the dp oracle (oracle/: C restatement of
nls.f90, kind-promoted).  Run in the build container -- CPU minutes that must not be spent on the GPU box:

    python tests/golden/make_golden_2d.py [c2_full] [c4_3steps] [c5_member]

The oracle's full field is reduced to what a test needs to pin the engine's field to 1e-10 without shipping
megabytes: a strided sub-sample, the complex sums and |psi|^2 sums of EVERY row and column (a wrong node anywhere
shows up in its row and its column), and the global norm.  Inputs are rebuilt by the tests from the same
deterministic recipes (`inputs` below), never stored.
"""

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

ORIG = dict(R=0.0242057488654, gamma=0.0242057488654, g=0.00162178517398, tilde_g=0.0169440242057,
            gamma_R=0.242057488654)

CASES = {
    # name: n, RK steps, pump radius, pump variation, amplitude of the rough part of u0 (0: the example's constant 0.1)
    "c2_full": dict(n=512, iters=5000, radius=10.0, variation=3.14, rough=0.0),        # examples/solve2d.py at 512^2, full horizon
    "c4_3steps": dict(n=8192, iters=3, radius=200.0, variation=50.0, rough=0.05),       # BASELINE config 4 size
    "c5_member": dict(n=1024, iters=200, radius=16.901960784313726, variation=3.14, rough=0.05),   # member 100 of config 5
}


def inputs(case):
    """(pumping, coeffs, u0, dx, dt, order) of a case -- the tests call this too."""
    from nls_b200.model import Problem
    from nls_b200.pumping import GaussianRingPumping2D
    c = CASES[case]
    n = c["n"]
    m = Problem().model(model="2d", dx=0.1, dt=1e-3, u0=0.1, order=5, num_nodes=n, num_iters=c["iters"],
                        pumping=GaussianRingPumping2D(power=20.0, radius=c["radius"], variation=c["variation"]),
                        original_params=dict(ORIG))
    u0 = np.full((n, n), 0.1 + 0j)
    if c["rough"]:
        rng = np.random.default_rng(n)
        u0 = u0 + c["rough"] * 0.3 * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    return m.getPumping(), m.getCoefficients(), u0, m.dx, m.dt, 5


def reduce_field(u):
    n = u.shape[0]
    s = max(1, n // 128)
    a2 = u.real ** 2 + u.imag ** 2
    return dict(stride=np.int64(s), sub=u[::s, ::s].copy(), row_sum=u.sum(axis=1), col_sum=u.sum(axis=0),
                row_abs2=a2.sum(axis=1), col_abs2=a2.sum(axis=0), norm=np.float64(np.sqrt(a2.sum())))


def main(names):
    from oracle import oracle as O
    for name in names:
        P, coeffs, u0, dx, dt, order = inputs(name)
        t0 = time.time()
        u = O.dp.solve_nls_2d(dt, dx, order, CASES[name]["iters"], P, coeffs, u0)
        red = reduce_field(u)
        red["iters"] = np.int64(CASES[name]["iters"])
        np.savez_compressed(os.path.join(HERE, "solve2d_%s.npz" % name), **red)
        print("%s: n=%d, %d steps, %.0f s, norm %.15g" % (name, u.shape[0], CASES[name]["iters"], time.time() - t0, red["norm"]))


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
