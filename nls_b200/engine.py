"""Device-resident front-end of the engine: torch tensors in HBM, kernels through the C ABI.

PyTorch is plumbing here -- it owns the complex128/float64 device buffers, the CUDA streams and (for
multi-GPU) ``torch.distributed``; every arithmetic kernel is in ``libnls_b200.so``.

What this adds on top of ``nls_b200.native`` (which copies host arrays in and out on every call like the
reference's f2py module does):

* state that stays on the GPU between calls (``Ensemble1D`` / ``Grid2D`` objects: psi, pumping,
  coefficient table and operator tables are uploaded once; ``advance(iters)`` can be called
  repeatedly -- the continuation pattern of ``tools/check.py:34-36`` and ``nls/animation.py:64-67``);
* the batched-ensemble mode: many independent parameter points (per-member ``coeffs[23]`` and
  pumping profile) advanced by one launch.

Reference semantics per member are those of ``solve_nls`` / ``solve_nls_2d`` (nls.f90:797-813,
:903-919).
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

__all__ = ["Ensemble1D", "Grid2D", "radial_taps", "cross_weights", "require_cuda", "device_pumping"]


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("nls_b200.engine needs a CUDA device: the engine has no CPU fallback")


def _dptr(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def set_2d_path(path):
    """Select the 2D kernels (C ABI ``nlsb_set_2d_path``): "auto" (strip-marching kernel from 2^20 nodes, TMA tile
    kernel below), "stream", "tma32" / "tma64" / "fused32" / "fused64" (tile kernel variants), "staged" (one launch
    per RK stage), "tma32_persistent" / "tma64_persistent" / "resident" (experimental whole-loop kernels)."""
    code = {"auto": 0, "staged": 1, "fused": 4, "fused32": 2, "fused64": 3, "tma32": 4, "tma64": 5,
            "tma32_persistent": 6, "tma64_persistent": 7, "stream": 8, "resident": 9}[path]
    _lib.call("nlsb_set_2d_path", code)


def radial_taps(n, order, dx):
    """Row-major tap table of the radial operator (host, float64, bit-identical to make_laplacian)."""
    taps = np.zeros((n, order), dtype=np.float64)
    _lib.call("nlsb_radial_taps", int(n), int(order), float(dx), taps.ctypes.data_as(C.c_void_p))
    return taps


def cross_weights(order, dx):
    """(wx, wy) weights of the 2D cross stencil (host, float64, bit-identical to make_laplacian_2d)."""
    wx, wy = np.zeros(order), np.zeros(order)
    _lib.call("nlsb_cross_weights", int(order), float(dx), wx.ctypes.data_as(C.c_void_p), wy.ctypes.data_as(C.c_void_p))
    return wx, wy


def _coeff_table(coeffs, batch):
    c = np.asarray(coeffs, dtype=np.float64)
    if c.ndim == 1:
        c = np.broadcast_to(c, (batch, 23))
    if c.shape != (batch, 23):
        raise ValueError("coeffs must have shape (23,) or (batch, 23); got %r" % (c.shape,))
    used = c[:, [2, 3, 4, 5, 11, 12, 13]]
    if not np.isfinite(used).all() or not (c[:, 12] > 0).all() or not (c[:, 13] >= 0).all():
        # precondition of the kernels' divide (csrc/device_math.cuh); the C ABI checks the same for host tables
        raise ValueError("coeffs: the entries read by the right-hand side must be finite, coeffs[12] > 0 and coeffs[13] >= 0 "
                         "(the reservoir denominator c13 + c14 |psi|^2 must stay positive)")
    return np.array(c, dtype=np.float64, order="C", copy=True)


def _to_device(a, dtype, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=dtype).contiguous()
    return torch.from_numpy(np.array(a, order="C", copy=True)).to(device=device, dtype=dtype)


def device_pumping(dim, kind, n, dx, power, variation, radius=0.0, x0=0.0, y0=0.0, device=None):
    """Pumping profiles of an ensemble sampled ON THE DEVICE on the reference's grid (``model.py:220-232``).

    kind: "gaussian" (``GaussianPumping1D/2D``) or "ring" (``GaussianRingPumping1D/2D``); power, variation, radius,
    x0, y0: scalars or arrays of one value per member (broadcast).  Returns a float64 tensor (batch, n) or
    (batch, n, n) that ``Ensemble1D`` / ``Grid2D`` accept as ``pumping``.  Grid and arithmetic are bit-identical to
    the host classes in ``nls_b200.pumping``; exp() may differ in the last place (C ABI nlsb_dev_pumping_profiles).
    """
    require_cuda()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    cols = np.broadcast_arrays(*[np.atleast_1d(np.asarray(v, dtype=np.float64)) for v in (power, x0, y0, variation, radius)])
    params = np.ascontiguousarray(np.stack(cols, axis=1))
    batch = params.shape[0]
    shape = (batch, int(n)) if int(dim) == 1 else (batch, int(n), int(n))
    out = torch.empty(shape, dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        _lib.call("nlsb_dev_pumping_profiles", int(dim), {"gaussian": 0, "ring": 1}[kind], batch, int(n), float(dx),
                  params.ctypes.data_as(C.c_void_p), _dptr(out), _stream())
    return out


_DIAG_HOST = {}


def _diagnostics_dict(out8):
    """Host view of the [batch][8] result of nlsb_dev_diagnostics_*: chemical potential mu = i E / M
    (nls.f90:946-947, :968-970) and the other scalars, one entry per member."""
    # page-locked staging buffer (cached per shape): a pageable copy of the 4 MB a 65 536-member ensemble returns takes
    # about a millisecond, as long as two of its RK steps
    key = (tuple(out8.shape), out8.device.index)
    host = _DIAG_HOST.get(key)
    if host is None:
        if len(_DIAG_HOST) > 8:
            _DIAG_HOST.clear()
        host = _DIAG_HOST[key] = torch.empty(out8.shape, dtype=out8.dtype).pin_memory()
    host.copy_(out8, non_blocking=True)
    torch.cuda.current_stream(out8.device).synchronize()
    d = host.numpy()
    M = d[:, 0] + 1j * d[:, 1]
    E = d[:, 2] + 1j * d[:, 3]
    return {"chemical_potential": 1j * E / M, "damping_integral": d[:, 4].copy(), "particles": d[:, 5].copy(),
            "max_density": d[:, 6].copy(), "max_reservoir": d[:, 7].copy()}


def _advance_until(engine, rel_tol, check_every, max_iters, quantity):
    """Shared body of ``advance_until``: chunks of ``check_every`` steps; the reduction rides inside the last launch of
    every chunk (``advance(chunk, diagnostics=True)``: the state one step before the chunk's end), so the field is
    read once per step and nothing but 8 doubles per member is copied back per chunk."""
    if check_every < 1 or max_iters < 0:
        raise ValueError("check_every must be positive and max_iters non-negative")
    previous, done, history = None, 0, []
    while done < max_iters:
        chunk = min(int(check_every), int(max_iters) - done)
        current = engine.advance(chunk, diagnostics=True)[quantity]
        done += chunk
        history.append(current)
        if previous is not None:
            change = np.max(np.abs(current - previous) / np.maximum(np.abs(current), np.finfo(float).tiny))
            if change <= rel_tol:
                return done, True, history
        previous = current
    return done, False, history


class Ensemble1D(object):
    """``batch`` independent radial systems of ``n`` nodes sharing dx, dt and the stencil order.

    pumping: (batch, n) or (n,); coeffs: (batch, 23) or (23,); u0: (batch, n), (n,) or scalar.
    """

    def __init__(self, n, dx, dt, order=5, batch=1, pumping=None, coeffs=None, u0=0.1, device=None):
        require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n, self.dx, self.dt, self.order, self.batch = int(n), float(dx), float(dt), int(order), int(batch)
        with torch.cuda.device(self.device):
            self.taps = _to_device(radial_taps(self.n, self.order, self.dx), torch.float64, self.device)
            self.coeffs = _to_device(_coeff_table(coeffs, self.batch), torch.float64, self.device)
            self.pumping = self._field(pumping, torch.float64)
            self.psi = self._field(u0, torch.complex128)
        self.steps_done = 0

    def _field(self, value, dtype):
        if isinstance(value, (int, float, complex)):
            return torch.full((self.batch, self.n), value, dtype=dtype, device=self.device)
        t = _to_device(value, dtype, self.device)
        if t.ndim == 1:
            t = t.unsqueeze(0).expand(self.batch, self.n).contiguous()
        if tuple(t.shape) != (self.batch, self.n):
            raise ValueError("expected shape (%d, %d), got %r" % (self.batch, self.n, tuple(t.shape)))
        return t

    def advance(self, iters, diagnostics=False):
        """``iters`` RK4 steps of every member, in place, asynchronously on the current stream; returns self.

        diagnostics=True (iters >= 1): the same steps, and the scalar diagnostics (as ``diagnostics()``) of the state
        ENTERING the last step -- step ``steps_done - 1`` after the call -- are reduced inside that step's launch
        (its first stage holds H(psi) of that state; C ABI nlsb_dev_rk4_1d_diag).  Returns the dictionary, with the
        step index under ``"step"``; reading it synchronises."""
        with torch.cuda.device(self.device):
            if not diagnostics:
                _lib.call("nlsb_dev_rk4_1d", self.batch, self.n, self.order, int(iters), self.dt, _dptr(self.taps),
                          _dptr(self.pumping), _dptr(self.coeffs), _dptr(self.psi), _stream())
                self.steps_done += int(iters)
                return self
            out = torch.empty((self.batch, 8), dtype=torch.float64, device=self.device)
            scratch = torch.empty(_lib.load().nlsb_dev_diagnostics_scratch(self.batch), dtype=torch.uint8, device=self.device)
            _lib.call("nlsb_dev_rk4_1d_diag", self.batch, self.n, self.order, int(iters), self.dt, self.dx, _dptr(self.taps),
                      _dptr(self.pumping), _dptr(self.coeffs), _dptr(self.psi), _dptr(scratch), _dptr(out), _stream())
        self.steps_done += int(iters)
        result = _diagnostics_dict(out)
        result["step"] = self.steps_done - 1
        return result

    def hamiltonian(self, u=None):
        u = self.psi if u is None else _to_device(u, torch.complex128, self.device)
        v = torch.empty_like(u)
        with torch.cuda.device(self.device):
            _lib.call("nlsb_dev_hamiltonian_1d", self.batch, self.n, self.order, _dptr(self.taps), _dptr(self.pumping),
                      _dptr(self.coeffs), _dptr(u), _dptr(v), _stream())
        return v

    def advance_until(self, rel_tol=1e-6, check_every=100, max_iters=100000, quantity="particles"):
        """Advance in chunks of ``check_every`` steps until ``quantity`` (a key of ``diagnostics()``) changes by at
        most ``rel_tol`` (relative, worst member) over a chunk, or ``max_iters`` steps were taken.  The field never
        leaves the device: per chunk 8 doubles per member cross the bus.  The reference has no stopping test
        (fixed ``iters``, nls.f90:705-734); ``tools/check.py:28-63`` does this loop by hand with a full
        solve -> copy -> chemical potential round trip per chunk.  Returns (steps taken, converged, history)."""
        return _advance_until(self, rel_tol, check_every, max_iters, quantity)

    def diagnostics(self):
        """Chemical potential, damping integral, particle number, peak density and peak reservoir of every member,
        reduced on the device in one pass (only 8 doubles per member cross the bus); see nlsb_dev_diagnostics_1d."""
        out = torch.empty((self.batch, 8), dtype=torch.float64, device=self.device)
        scratch = torch.empty(_lib.load().nlsb_dev_diagnostics_scratch(self.batch), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("nlsb_dev_diagnostics_1d", self.batch, self.n, self.order, self.dx, _dptr(self.taps),
                      _dptr(self.pumping), _dptr(self.coeffs), _dptr(self.psi), _dptr(scratch), _dptr(out), _stream())
        return _diagnostics_dict(out)

    def set_pumping(self, pumping):
        """Replace the pumping profile(s) between chunks of steps; psi and the tap table stay on the device
        (the continuation pattern of nls/animation.py:64-101 and tools/check.py:34-36)."""
        with torch.cuda.device(self.device):
            self.pumping.copy_(self._field(pumping, torch.float64))
        return self

    def set_coefficients(self, coeffs):
        with torch.cuda.device(self.device):
            self.coeffs.copy_(_to_device(_coeff_table(coeffs, self.batch), torch.float64, self.device))
        return self

    def solution(self):
        return self.psi.cpu().numpy()


class Grid2D(object):
    """``batch`` independent n x n (or rows x cols) Cartesian grids sharing dx, dt and the order.

    pumping: (batch, rows, cols) or (rows, cols); coeffs: (batch, 23) or (23,); u0 likewise or scalar.
    """

    def __init__(self, n, dx, dt, order=5, batch=1, pumping=None, coeffs=None, u0=0.1, device=None, cols=None):
        require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.rows, self.cols = int(n), int(n if cols is None else cols)
        self.dx, self.dt, self.order, self.batch = float(dx), float(dt), int(order), int(batch)
        self.wx, self.wy = cross_weights(self.order, self.dx)
        with torch.cuda.device(self.device):
            table = _coeff_table(coeffs, self.batch)
            # members that all share one coefficient set let the kernel keep it in its constant bank
            self.shared_coeffs = table[0].copy() if bool((table == table[0]).all()) else None
            self.coeffs = _to_device(table, torch.float64, self.device)
            self.pumping = self._field(pumping, torch.float64)
            self.psi = self._field(u0, torch.complex128)
            nbytes = _lib.load().nlsb_dev_rk4_2d_workspace(self.batch, self.rows, self.cols)
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self._diag_scratch = None
        self.steps_done = 0

    def _field(self, value, dtype):
        shape = (self.batch, self.rows, self.cols)
        if isinstance(value, (int, float, complex)):
            return torch.full(shape, value, dtype=dtype, device=self.device)
        t = _to_device(value, dtype, self.device)
        if t.ndim == 2:
            t = t.unsqueeze(0).expand(*shape).contiguous()
        if tuple(t.shape) != shape:
            raise ValueError("expected shape %r, got %r" % (shape, tuple(t.shape)))
        return t

    def _w(self, a):
        return a.ctypes.data_as(C.c_void_p)

    def advance(self, iters, diagnostics=False):
        """``iters`` RK4 steps of every member, in place, asynchronously on the current stream; returns self.

        diagnostics=True (iters >= 1): as ``Ensemble1D.advance`` -- the diagnostics of the state entering the last
        step, reduced inside that step's launch by the strip-marching kernel (C ABI nlsb_dev_rk4_2d_diag; kernels
        without the fused reduction run the stand-alone pass before their last step: same numbers)."""
        shared = self._w(self.shared_coeffs) if self.shared_coeffs is not None else None
        with torch.cuda.device(self.device):
            if not diagnostics:
                _lib.call("nlsb_dev_rk4_2d", self.batch, self.rows, self.cols, self.order, int(iters), self.dt,
                          self._w(self.wx), self._w(self.wy), _dptr(self.pumping), _dptr(self.coeffs), shared, _dptr(self.psi),
                          _dptr(self.workspace), C.c_size_t(self.workspace.numel()), _stream())
                self.steps_done += int(iters)
                return self
            out = torch.empty((self.batch, 8), dtype=torch.float64, device=self.device)
            if self._diag_scratch is None:
                nbytes = _lib.load().nlsb_dev_rk4_2d_diag_scratch(self.batch, self.rows, self.cols, self.order)
                self._diag_scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            _lib.call("nlsb_dev_rk4_2d_diag", self.batch, self.rows, self.cols, self.order, int(iters), self.dt, self.dx,
                      self._w(self.wx), self._w(self.wy), _dptr(self.pumping), _dptr(self.coeffs), shared, _dptr(self.psi),
                      _dptr(self.workspace), C.c_size_t(self.workspace.numel()), _dptr(self._diag_scratch), _dptr(out), _stream())
        self.steps_done += int(iters)
        result = _diagnostics_dict(out)
        result["step"] = self.steps_done - 1
        return result

    def hamiltonian(self, u=None):
        u = self.psi if u is None else _to_device(u, torch.complex128, self.device)
        v = torch.empty_like(u)
        with torch.cuda.device(self.device):
            _lib.call("nlsb_dev_hamiltonian_2d", self.batch, self.rows, self.cols, self.order, self._w(self.wx),
                      self._w(self.wy), _dptr(self.pumping), _dptr(self.coeffs), _dptr(u), _dptr(v), _stream())
        return v

    def advance_until(self, rel_tol=1e-6, check_every=100, max_iters=100000, quantity="particles"):
        """Advance in chunks of ``check_every`` steps until ``quantity`` (a key of ``diagnostics()``) changes by at
        most ``rel_tol`` (relative, worst member) over a chunk, or ``max_iters`` steps were taken.  The field never
        leaves the device: per chunk 8 doubles per member cross the bus.  The reference has no stopping test
        (fixed ``iters``, nls.f90:705-734); ``tools/check.py:28-63`` does this loop by hand with a full
        solve -> copy -> chemical potential round trip per chunk.  Returns (steps taken, converged, history)."""
        return _advance_until(self, rel_tol, check_every, max_iters, quantity)

    def diagnostics(self):
        """As ``Ensemble1D.diagnostics`` for 2D grids (area element dx^2, chemical potential weight 1)."""
        out = torch.empty((self.batch, 8), dtype=torch.float64, device=self.device)
        scratch = torch.empty(_lib.load().nlsb_dev_diagnostics_scratch(self.batch), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("nlsb_dev_diagnostics_2d", self.batch, self.rows, self.cols, self.order, self.dx,
                      self._w(self.wx), self._w(self.wy), _dptr(self.pumping), _dptr(self.coeffs), _dptr(self.psi),
                      _dptr(scratch), _dptr(out), _stream())
        return _diagnostics_dict(out)

    def set_pumping(self, pumping):
        """Replace the pumping profile(s) between chunks of steps; psi and the operator stay on the device
        (the continuation pattern of nls/animation.py:64-101 and tools/check.py:34-36)."""
        with torch.cuda.device(self.device):
            self.pumping.copy_(self._field(pumping, torch.float64))
        return self

    def set_coefficients(self, coeffs):
        table = _coeff_table(coeffs, self.batch)
        self.shared_coeffs = table[0].copy() if bool((table == table[0]).all()) else None
        with torch.cuda.device(self.device):
            self.coeffs.copy_(_to_device(table, torch.float64, self.device))
        return self

    def solution(self):
        return self.psi.cpu().numpy()
