"""ctypes front-end of the CPU oracle (``oracle/nls_oracle.c``).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module; the engine package ``nls_b200`` never does.

Two instances are exported, mirroring the f2py module ``nls.native.nls`` of the reference
(signatures from SURVEY.md 8b):

    ``sp`` -- float32 / complex64, the kind the reference ships (nls/nls.f90:10)
    ``dp`` -- float64 / complex128, the kind-promoted restatement (the engine's parity target)

Index convention: numpy ``a[i, j]`` is Fortran ``a(i+1, j+1)`` exactly as f2py presents the
reference's arrays; internally arrays are handed to C in Fortran order.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnls_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc -O3 -ffp-contract=off)."""
    src = [os.path.join(_HERE, f) for f in ("nls_oracle.c", "nls_oracle_impl.h")]
    stale = (not os.path.exists(_LIB_PATH)
             or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "libnls_oracle.so"])
    return _LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


class _Kind(object):
    def __init__(self, suffix, real, cplx, creal):
        self.suffix, self.real, self.cplx, self.creal = suffix, np.dtype(real), np.dtype(cplx), creal

    # -- helpers ----------------------------------------------------------------------------------
    def _fn(self, name):
        return getattr(_load(), "nlso_%s_%s" % (name, self.suffix))

    def _r(self, a):
        return np.asfortranarray(a, dtype=self.real)

    def _c(self, a):
        return np.asfortranarray(a, dtype=self.cplx)

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    @staticmethod
    def _check(rc, what):
        if rc != 0:
            raise ValueError("oracle %s failed with status %d" % (what, rc))

    def _coeffs(self, coeffs):
        c = self._r(coeffs)
        if c.shape != (23,):
            raise ValueError("coeffs must have 23 entries")
        return c

    # -- operator builders ------------------------------------------------------------------------
    def make_banded_matrix(self, n, row):
        row = self._r(row)
        m = row.shape[0]
        mat = np.zeros((m, n), dtype=self.real, order="F")
        self._fn("make_banded_matrix")(C.c_int(n), C.c_int(m), self._p(row), self._p(mat))
        return mat

    def make_laplacian(self, n, m, h, legacy24=False):
        op = np.zeros((m, n), dtype=self.real, order="F")
        if legacy24:
            if m != 5:
                raise ValueError("legacy24 only exists for order 5")
            rc = self._fn("make_laplacian_o5_legacy24")(C.c_int(n), self.creal(h), self._p(op))
        else:
            rc = self._fn("make_laplacian")(C.c_int(n), C.c_int(m), self.creal(h), self._p(op))
        self._check(rc, "make_laplacian")
        return op

    def make_laplacian_2d(self, n, m, h):
        blocks = np.zeros((n, 2 * m - 1), dtype=self.real, order="F")
        orders = np.zeros(m, dtype=np.int32)
        rc = self._fn("make_laplacian_2d")(C.c_int(n), C.c_int(m), self.creal(h), self._p(blocks), self._p(orders))
        self._check(rc, "make_laplacian_2d")
        return blocks, orders

    # -- matvecs ----------------------------------------------------------------------------------
    def rgbmv(self, x, u, sign, op):
        """Returns u + sign * A x (the reference updates ``u`` in place)."""
        op = self._r(op)
        klu = (op.shape[0] - 1) // 2
        x = self._r(x)
        out = self._r(u).copy()
        self._fn("rgbmv")(self._p(x), self._p(out), self.creal(sign), self._p(op), C.c_int(klu), C.c_int(x.shape[0]))
        return out

    def rbbmv(self, x, y, sign, blocks, ms, n):
        blocks = self._r(blocks)
        ms = np.ascontiguousarray(ms, dtype=np.int32)
        x = np.ascontiguousarray(np.asarray(x, dtype=self.real).ravel(order="F"))
        out = np.array(np.asarray(y, dtype=self.real).ravel(order="F"), copy=True)
        rc = self._fn("rbbmv")(self._p(x), self._p(out), self.creal(sign), self._p(blocks), self._p(ms),
                               C.c_int(ms.shape[0]), C.c_int(n))
        self._check(rc, "rbbmv")
        return out

    # -- right-hand side --------------------------------------------------------------------------
    def revervoir(self, pumping, coeffs, u_sqr):
        p, q = self._r(pumping), self._r(u_sqr)
        r = np.zeros(p.shape, dtype=self.real, order="F")
        self._fn("revervoir")(self._p(p), self._p(self._coeffs(coeffs)), self._p(q), self._p(r), C.c_size_t(p.size))
        return r

    def hamiltonian(self, pumping, coeffs, u, op):
        op, u, p = self._r(op), self._c(u), self._r(pumping)
        v = np.zeros(u.shape, dtype=self.cplx)
        rc = self._fn("hamiltonian")(self._p(p), self._p(self._coeffs(coeffs)), self._p(u), self._p(v), self._p(op),
                                     C.c_int((op.shape[0] - 1) // 2), C.c_int(u.shape[0]))
        self._check(rc, "hamiltonian")
        return v

    def hamiltonian_2d(self, pumping, coeffs, u, blocks, orders):
        u, p, blocks = self._c(u), self._r(pumping), self._r(blocks)
        orders = np.ascontiguousarray(orders, dtype=np.int32)
        v = np.zeros(u.shape, dtype=self.cplx, order="F")
        rc = self._fn("hamiltonian_2d")(self._p(p), self._p(self._coeffs(coeffs)), self._p(u), self._p(v),
                                        self._p(blocks), self._p(orders), C.c_int(orders.shape[0]), C.c_int(u.shape[0]))
        self._check(rc, "hamiltonian_2d")
        return v

    # -- time stepping ----------------------------------------------------------------------------
    def runge_kutta(self, dt, t0, u0, op, iters, pumping, coeffs):
        op, u0, p = self._r(op), self._c(u0), self._r(pumping)
        u = np.zeros(u0.shape, dtype=self.cplx)
        rc = self._fn("runge_kutta")(self.creal(dt), self.creal(t0), self._p(u0), self._p(op), C.c_int(u0.shape[0]),
                                     C.c_int(op.shape[0]), C.c_int(iters), self._p(u), self._p(p),
                                     self._p(self._coeffs(coeffs)))
        self._check(rc, "runge_kutta")
        return u

    def runge_kutta_2d(self, dt, t0, u0, blocks, orders, iters, pumping, coeffs):
        u0, p, blocks = self._c(u0), self._r(pumping), self._r(blocks)
        orders = np.ascontiguousarray(orders, dtype=np.int32)
        u = np.zeros(u0.shape, dtype=self.cplx, order="F")
        rc = self._fn("runge_kutta_2d")(self.creal(dt), self.creal(t0), self._p(u0), C.c_int(u0.shape[0]),
                                        self._p(blocks), self._p(orders), C.c_int(orders.shape[0]), C.c_int(iters),
                                        self._p(u), self._p(p), self._p(self._coeffs(coeffs)))
        self._check(rc, "runge_kutta_2d")
        return u

    def solve_nls(self, dt, dx, order, iters, pumping, coeffs, u0):
        u0, p = self._c(u0), self._r(pumping)
        if p.shape != u0.shape or u0.ndim != 1:
            raise ValueError("pumping and u0 must be 1D arrays of the same length")
        u = np.zeros(u0.shape, dtype=self.cplx)
        rc = self._fn("solve_nls")(self.creal(dt), self.creal(dx), C.c_int(u0.shape[0]), C.c_int(order), C.c_int(iters),
                                   self._p(p), self._p(self._coeffs(coeffs)), self._p(u0), self._p(u))
        self._check(rc, "solve_nls")
        return u

    solve_nls_1d = solve_nls

    def solve_nls_2d(self, dt, dx, order, iters, pumping, coeffs, u0):
        u0, p = self._c(u0), self._r(pumping)
        if p.shape != u0.shape or u0.ndim != 2 or u0.shape[0] != u0.shape[1]:
            raise ValueError("pumping and u0 must be square 2D arrays of the same shape")
        u = np.zeros(u0.shape, dtype=self.cplx, order="F")
        rc = self._fn("solve_nls_2d")(self.creal(dt), self.creal(dx), C.c_int(u0.shape[0]), C.c_int(order),
                                      C.c_int(iters), self._p(p), self._p(self._coeffs(coeffs)), self._p(u0), self._p(u))
        self._check(rc, "solve_nls_2d")
        return u

    # -- diagnostics ------------------------------------------------------------------------------
    def chemical_potential_1d(self, dx, pumping, coeffs, u0):
        u0, p = self._c(u0), self._r(pumping)
        mu = np.zeros(1, dtype=self.cplx)
        rc = self._fn("chemical_potential_1d")(self.creal(dx), C.c_int(u0.shape[0]), self._p(p),
                                               self._p(self._coeffs(coeffs)), self._p(u0), self._p(mu))
        self._check(rc, "chemical_potential_1d")
        return mu[0]

    def chemical_potential_2d(self, dx, pumping, coeffs, u0):
        u0, p = self._c(u0), self._r(pumping)
        mu = np.zeros(1, dtype=self.real)
        rc = self._fn("chemical_potential_2d")(self.creal(dx), C.c_int(u0.shape[0]), self._p(p),
                                               self._p(self._coeffs(coeffs)), self._p(u0), self._p(mu))
        self._check(rc, "chemical_potential_2d")
        return mu[0]

    @staticmethod
    def version():
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        _load().nlso_version(C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value


sp = _Kind("sp", np.float32, np.complex64, C.c_float)
dp = _Kind("dp", np.float64, np.complex128, C.c_double)
