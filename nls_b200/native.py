"""``nls_b200.native`` -- stand-in for the reference's f2py extension ``nls.native``.

The reference builds ``nls.native`` from ``nls/nls.f90`` with f2py (``setup.py:69-79``) and Python
reaches the numerical core as ``nls.native.nls.<routine>`` (``nls/solver.py:9``, ``README.md:44-45``).
This module exposes the same attribute ``nls`` with the same routine names and the f2py call
signatures (shape-inferable integers are trailing optionals), backed by the CUDA
engine through the C ABI of ``libnls_b200.so``.

Differences from the f2py module, all deliberate (DESIGN.md "Precision contract"):

* arithmetic and results are float64 / complex128 (the reference down-casts to float32/complex64);
* an unsupported ``order`` raises :class:`error` instead of returning uninitialised memory;
* 2D results are returned C-ordered; ``a[i, j]`` still addresses Fortran ``a(i+1, j+1)``.

There is no CPU path here: every routine that computes runs on the current CUDA device.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import NativeError as error

__all__ = ["nls", "error"]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _real(a, order="C"):
    return np.require(a, dtype=np.float64, requirements=[order, "ALIGNED"])


def _cplx(a, order="C"):
    return np.require(a, dtype=np.complex128, requirements=[order, "ALIGNED"])


def _coeffs(c):
    c = np.ascontiguousarray(c, dtype=np.float64)
    if c.shape != (23,):
        raise ValueError("coeffs must be a vector of 23 reals (got shape %r)" % (c.shape,))
    return c


def _same_shape(what, ref_name, ref, **arrays):
    """f2py checks every array extent against the inferred n (``check(shape(pumping,0)==n)``) and raises; the C ABI
    takes raw pointers, so the same check happens here before anything is handed over."""
    for name, a in arrays.items():
        if a.shape != ref.shape:
            raise ValueError("%s: %s has shape %r but %s has shape %r" % (what, name, a.shape, ref_name, ref.shape))


def _band(op, what, n, klu=None):
    """A band matrix (2 klu + 1, n) in BLAS GB storage; returns klu."""
    if op.ndim != 2 or op.shape[0] % 2 == 0 or op.shape[1] != n:
        raise ValueError("%s: op must have shape (2*klu+1, %d), got %r" % (what, n, op.shape))
    k = (op.shape[0] - 1) // 2
    _check_n(klu, k, what)
    return k


def _blocks(blocks, orders, what, n):
    m = orders.shape[0]
    if orders.ndim != 1 or blocks.ndim != 2 or blocks.shape != (n, 2 * m - 1):
        raise ValueError("%s: blocks must have shape (%d, 2*m-1) with m = len(orders) = %d, got %r"
                         % (what, n, m, blocks.shape))
    return m


_PINNED_FROM = 1 << 20   # results of at least this many bytes are returned in page-locked memory


def _result_like(a, order="C"):
    """Uninitialised result array of a's shape.  Large results live in page-locked host memory (from torch's caching
    host allocator, reused between calls) so that the device-to-host copy runs at the speed of the bus; the array
    is an ordinary numpy array to the caller (the f2py module returns a fresh numpy-owned array too)."""
    if a.nbytes >= _PINNED_FROM and a.ndim in (1, 2):
        try:
            import torch
            if torch.cuda.is_available():
                dt = {np.dtype(np.float64): torch.float64, np.dtype(np.complex128): torch.complex128}[a.dtype]
                shape = a.shape if order == "C" else a.shape[::-1]
                out = torch.empty(shape, dtype=dt, pin_memory=True).numpy()
                return out if order == "C" else out.T
        except Exception:       # no torch / no pinned memory: pageable result, same values
            pass
    return np.empty(a.shape, dtype=a.dtype, order=order)


def _check_n(n, inferred, what):
    if n is not None and int(n) != inferred:
        raise ValueError("%s: n = %d does not match the array extent %d" % (what, n, inferred))


def _square(a, what):
    if a.ndim != 2 or a.shape[0] != a.shape[1]:
        raise ValueError("%s must be a square 2D array (got shape %r)" % (what, a.shape))
    return a.shape[0]


class _Module(object):
    """The routines of ``module nls`` (nls/nls.f90:13-24) with their f2py signatures."""

    error = error

    # -- version (nls.f90:29-37) -------------------------------------------------------------------
    @staticmethod
    def version():
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        _lib.load().nlsb_version(C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    # -- operator builders (host side of the engine) ----------------------------------------------
    @staticmethod
    def make_banded_matrix(n, row, m=None):
        row = _real(row)
        _check_n(m, row.shape[0], "make_banded_matrix")
        mat = np.zeros((row.shape[0], n), dtype=np.float64, order="F")
        _lib.call("nlsb_make_banded_matrix", n, row.shape[0], _ptr(row), _ptr(mat))
        return mat

    @staticmethod
    def clear_first_row_of_derivative(L1):
        out = np.array(L1, dtype=np.float64, order="F")
        _lib.call("nlsb_clear_first_row_of_derivative", out.shape[1], out.shape[0], _ptr(out))
        return out

    @staticmethod
    def divide_derivative_on_radius(h, L1):
        out = np.array(L1, dtype=np.float64, order="F")
        _lib.call("nlsb_divide_derivative_on_radius", out.shape[1], out.shape[0], float(h), _ptr(out))
        return out

    @staticmethod
    def make_laplacian(n, m, h):
        op = np.zeros((m, n), dtype=np.float64, order="F")
        _lib.call("nlsb_make_laplacian", n, m, float(h), _ptr(op))
        return op

    def make_laplacian_o3(self, n, h):
        return self.make_laplacian(n, 3, h)

    def make_laplacian_o5(self, n, h):
        return self.make_laplacian(n, 5, h)

    def make_laplacian_o7(self, n, h):
        return self.make_laplacian(n, 7, h)

    @staticmethod
    def make_laplacian_2d(n, m, h):
        blocks = np.zeros((n, 2 * m - 1), dtype=np.float64, order="F")
        orders = np.zeros(m, dtype=np.int32)
        _lib.call("nlsb_make_laplacian_2d", n, m, float(h), _ptr(blocks), _ptr(orders))
        return blocks, orders

    def make_laplacian_2d_o3(self, n, h):
        return self.make_laplacian_2d(n, 3, h)

    def make_laplacian_2d_o5(self, n, h):
        return self.make_laplacian_2d(n, 5, h)

    def make_laplacian_2d_o7(self, n, h):
        return self.make_laplacian_2d(n, 7, h)

    # -- matvecs (in place on the second argument, as f2py's intent(inout)) ---------------------------
    @staticmethod
    def rgbmv(x, u, sign, op, klu=None, n=None):
        op = _real(op, "F")
        if not (isinstance(u, np.ndarray) and u.dtype == np.float64 and u.flags.c_contiguous and u.ndim == 1):
            raise TypeError("rgbmv: u is updated in place and must be a contiguous float64 vector")
        x = _real(x)
        _check_n(n, u.shape[0], "rgbmv")
        _same_shape("rgbmv", "u", u, x=x)
        k = _band(op, "rgbmv", u.shape[0], klu)
        _lib.call("nlsb_rgbmv", _ptr(x), _ptr(u), float(sign), _ptr(op), k, u.shape[0])

    @staticmethod
    def rbbmv(x, y, sign, blocks, ms, n, m=None):
        blocks = _real(blocks, "F")
        ms = np.ascontiguousarray(ms, dtype=np.int32)
        _check_n(m, ms.shape[0], "rbbmv")
        if not (isinstance(y, np.ndarray) and y.dtype == np.float64 and y.flags.c_contiguous and y.ndim == 1):
            raise TypeError("rbbmv: y is updated in place and must be a contiguous float64 vector of n*n")
        x = _real(x)
        if x.shape != (n * n,) or y.shape != (n * n,):
            raise ValueError("rbbmv: x and y must have n*n entries")
        _blocks(blocks, ms, "rbbmv", n)
        _lib.call("nlsb_rbbmv", _ptr(x), _ptr(y), float(sign), _ptr(blocks), _ptr(ms), ms.shape[0], n)

    def rbbmv_o3(self, x, y, sign, blocks, ms, n):
        return self.rbbmv(x, y, sign, blocks, ms, n, 3)

    def rbbmv_o5(self, x, y, sign, blocks, ms, n):
        return self.rbbmv(x, y, sign, blocks, ms, n, 5)

    def rbbmv_o7(self, x, y, sign, blocks, ms, n):
        return self.rbbmv(x, y, sign, blocks, ms, n, 7)

    # -- reservoir ---------------------------------------------------------------------------------
    @staticmethod
    def revervoir(pumping, coeffs, u_sqr, n=None):
        p, q = _real(pumping), _real(u_sqr)
        if p.ndim != 1:
            raise ValueError("revervoir: pumping must be a vector (got shape %r)" % (p.shape,))
        _check_n(n, p.shape[0], "revervoir")
        _same_shape("revervoir", "pumping", p, u_sqr=q)
        r = np.empty_like(p)
        _lib.call("nlsb_revervoir", _ptr(p), _ptr(_coeffs(coeffs)), _ptr(q), _ptr(r), p.shape[0])
        return r

    @staticmethod
    def revervoir_2d(pumping, coeffs, u_sqr, n=None):
        p, q = _real(pumping), _real(u_sqr)
        _check_n(n, _square(p, "pumping"), "revervoir_2d")
        _same_shape("revervoir_2d", "pumping", p, u_sqr=q)
        r = _result_like(p)
        _lib.call("nlsb_revervoir_2d", _ptr(p), _ptr(_coeffs(coeffs)), _ptr(q), _ptr(r), p.shape[0])
        return r

    # -- right-hand side ---------------------------------------------------------------------------
    @staticmethod
    def hamiltonian(pumping, coeffs, u, op, klu=None, n=None):
        op, u, p = _real(op, "F"), _cplx(u), _real(pumping)
        if u.ndim != 1:
            raise ValueError("hamiltonian: u must be a vector (got shape %r)" % (u.shape,))
        _check_n(n, u.shape[0], "hamiltonian")
        _same_shape("hamiltonian", "u", u, pumping=p)
        k = _band(op, "hamiltonian", u.shape[0], klu)
        v = np.empty_like(u)
        _lib.call("nlsb_hamiltonian", _ptr(p), _ptr(_coeffs(coeffs)), _ptr(u), _ptr(v), _ptr(op), k, u.shape[0])
        return v

    @staticmethod
    def hamiltonian_2d(pumping, coeffs, u, blocks, orders, order=None, n=None):
        # handed over in Fortran order: a user-supplied block operator need not be transpose-symmetric
        blocks, u, p = _real(blocks, "F"), _cplx(u, "F"), _real(pumping, "F")
        orders = np.ascontiguousarray(orders, dtype=np.int32)
        _check_n(n, _square(u, "u"), "hamiltonian_2d")
        _same_shape("hamiltonian_2d", "u", u, pumping=p)
        _check_n(order, _blocks(blocks, orders, "hamiltonian_2d", u.shape[0]), "hamiltonian_2d")
        v = _result_like(u, "F")
        _lib.call("nlsb_hamiltonian_2d", _ptr(p), _ptr(_coeffs(coeffs)), _ptr(u), _ptr(v), _ptr(blocks),
                  _ptr(orders), orders.shape[0], u.shape[0])
        return v

    # -- time stepping -----------------------------------------------------------------------------
    @staticmethod
    def runge_kutta(dt, t0, u0, op, iters, pumping, coeffs, n=None, order=None):
        op, u0, p = _real(op, "F"), _cplx(u0), _real(pumping)
        if u0.ndim != 1:
            raise ValueError("runge_kutta: u0 must be a vector (got shape %r)" % (u0.shape,))
        _check_n(n, u0.shape[0], "runge_kutta")
        _same_shape("runge_kutta", "u0", u0, pumping=p)
        _check_n(order, 2 * _band(op, "runge_kutta", u0.shape[0]) + 1, "runge_kutta")
        u = np.empty_like(u0)
        _lib.call("nlsb_runge_kutta", float(dt), float(t0), _ptr(u0), _ptr(op), u0.shape[0], op.shape[0],
                  int(iters), _ptr(u), _ptr(p), _ptr(_coeffs(coeffs)))
        return u

    @staticmethod
    def runge_kutta_2d(dt, t0, u0, blocks, orders, iters, pumping, coeffs, n=None, order=None):
        blocks, u0, p = _real(blocks, "F"), _cplx(u0, "F"), _real(pumping, "F")
        orders = np.ascontiguousarray(orders, dtype=np.int32)
        _check_n(n, _square(u0, "u0"), "runge_kutta_2d")
        _same_shape("runge_kutta_2d", "u0", u0, pumping=p)
        _check_n(order, _blocks(blocks, orders, "runge_kutta_2d", u0.shape[0]), "runge_kutta_2d")
        u = _result_like(u0, "F")
        _lib.call("nlsb_runge_kutta_2d", float(dt), float(t0), _ptr(u0), u0.shape[0], _ptr(blocks), _ptr(orders),
                  orders.shape[0], int(iters), _ptr(u), _ptr(p), _ptr(_coeffs(coeffs)))
        return u

    # -- entry points used by the solver facade (ref solver.py:65-69, :79-83) -----------------------
    @staticmethod
    def solve_nls(dt, dx, order, iters, pumping, coeffs, u0, n=None):
        u0, p = _cplx(u0), _real(pumping)
        if u0.ndim != 1 or p.shape != u0.shape:
            raise ValueError("solve_nls: pumping and u0 must be vectors of the same length")
        _check_n(n, p.shape[0], "solve_nls")
        u = np.empty_like(u0)
        _lib.call("nlsb_solve_nls", float(dt), float(dx), p.shape[0], int(order), int(iters), _ptr(p),
                  _ptr(_coeffs(coeffs)), _ptr(u0), _ptr(u))
        return u

    def solve_nls_1d(self, dt, dx, order, iters, pumping, coeffs, u0, n=None):
        return self.solve_nls(dt, dx, order, iters, pumping, coeffs, u0, n)

    @staticmethod
    def solve_nls_2d(dt, dx, order, iters, pumping, coeffs, u0, n=None):
        # The cross stencil built inside solve_nls_2d is transpose-symmetric and every other term is
        # pointwise, so a C-ordered (n, n) buffer can be handed over as is -- no Fortran-order copy
        # (f2py makes one on every call) -- and the C-ordered result indexes identically.
        u0, p = _cplx(u0), _real(pumping)
        nn = _square(u0, "u0")
        if p.shape != u0.shape:
            raise ValueError("solve_nls_2d: pumping and u0 must have the same shape")
        _check_n(n, nn, "solve_nls_2d")
        u = _result_like(u0)
        _lib.call("nlsb_solve_nls_2d", float(dt), float(dx), nn, int(order), int(iters), _ptr(p),
                  _ptr(_coeffs(coeffs)), _ptr(u0), _ptr(u))
        return u

    @staticmethod
    def chemical_potential_1d(dx, pumping, coeffs, u0, n=None):
        u0, p = _cplx(u0), _real(pumping)
        if u0.ndim != 1:
            raise ValueError("chemical_potential_1d: u0 must be a vector (got shape %r)" % (u0.shape,))
        _check_n(n, u0.shape[0], "chemical_potential_1d")
        _same_shape("chemical_potential_1d", "u0", u0, pumping=p)
        mu = np.zeros(1, dtype=np.complex128)
        _lib.call("nlsb_chemical_potential_1d", float(dx), u0.shape[0], _ptr(p), _ptr(_coeffs(coeffs)), _ptr(u0), _ptr(mu))
        return complex(mu[0])

    @staticmethod
    def chemical_potential_2d(dx, pumping, coeffs, u0, n=None):
        u0, p = _cplx(u0), _real(pumping)
        _check_n(n, _square(u0, "u0"), "chemical_potential_2d")
        _same_shape("chemical_potential_2d", "u0", u0, pumping=p)
        mu = np.zeros(1, dtype=np.float64)
        _lib.call("nlsb_chemical_potential_2d", float(dx), u0.shape[0], _ptr(p), _ptr(_coeffs(coeffs)), _ptr(u0), _ptr(mu))
        return float(mu[0])


nls = _Module()
