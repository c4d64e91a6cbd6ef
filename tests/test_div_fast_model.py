"""The reciprocal-based divide of the kernels (``nls_b200/csrc/device_math.cuh::div_fast``): exact rational model
of its operation sequence -- seed of relative error <= 2^-20 truncated to its high 32 bits (MUFU.RCP64H is specified
to 2^-23), ONE Newton step, one FMA residual correction of the quotient -- against the true quotient.  The claim in
the header (correctly rounded for the operand ranges of the reservoir term) is checked here without a GPU; the
device code itself is held to the oracle by the right-hand-side parity tests (1e-13)."""

import math
import random
import struct
from fractions import Fraction as F


def _fma(a, b, c):
    return float(F(a) * F(b) + F(c))          # Fraction -> float rounds to nearest: an exact fused multiply-add


def _high_word(x):
    bits = struct.unpack("<Q", struct.pack("<d", x))[0] & 0xFFFFFFFF00000000
    return struct.unpack("<d", struct.pack("<Q", bits))[0]


def div_fast_model(a, b, seed_error):
    x = _high_word(float(F(1) / F(b) * (1 + F(seed_error))))
    e = _fma(-b, x, 1.0)
    x = _fma(x, e, x)
    q = float(F(a) * F(x))
    r = _fma(-b, q, a)
    return _fma(r, x, q)


def test_one_newton_step_plus_residual_correction_is_correctly_rounded():
    rng = random.Random(1)
    worst = F(0)
    for _ in range(20000):
        a = rng.uniform(0.0, 50.0) * rng.choice([1e-6, 1e-3, 1.0, 1e3])               # c12 * P
        b = 1.0 + rng.uniform(0.0, 1e4) * rng.choice([1e-8, 1e-4, 1.0])                # c13 + c14 |psi|^2 >= 1
        err = rng.choice([-1, 1]) * rng.uniform(0.5, 1.0) * 2.0 ** -20
        got, true = div_fast_model(a, b, err), F(a) / F(b)
        if a > 0:
            worst = max(worst, abs(F(got) - true) / F(math.ulp(float(true))))
    assert worst <= F(1, 2)
