"""Loader and ctypes prototypes of ``libnls_b200.so`` (the C ABI declared in ``include/nls_b200.h``).

The library is the product: there is no Python or CPU fallback.  If it is missing it is built with
nvcc (``nls_b200/_build.py``); if that is impossible the import fails loudly.
"""

from __future__ import annotations

import ctypes as C
import os

from . import _build

_D = C.POINTER(C.c_double)   # documentation only: pointers are passed as c_void_p
_P = C.c_void_p
_I = C.c_int
_F = C.c_double
_Z = C.c_size_t

# name -> (restype, argtypes); mirrors include/nls_b200.h declaration by declaration
PROTOTYPES = {
    "nlsb_last_error": (C.c_char_p, []),
    "nlsb_device_available": (_I, []),
    "nlsb_kernel_launches": (C.c_ulonglong, []),
    "nlsb_trim_memory": (_I, []),
    "nlsb_version": (None, [_P, _P, _P]),
    "nlsb_make_banded_matrix": (_I, [_I, _I, _P, _P]),
    "nlsb_clear_first_row_of_derivative": (_I, [_I, _I, _P]),
    "nlsb_divide_derivative_on_radius": (_I, [_I, _I, _F, _P]),
    "nlsb_make_laplacian": (_I, [_I, _I, _F, _P]),
    "nlsb_make_laplacian_o3": (_I, [_I, _F, _P]),
    "nlsb_make_laplacian_o5": (_I, [_I, _F, _P]),
    "nlsb_make_laplacian_o7": (_I, [_I, _F, _P]),
    "nlsb_make_laplacian_2d": (_I, [_I, _I, _F, _P, _P]),
    "nlsb_make_laplacian_2d_o3": (_I, [_I, _F, _P, _P]),
    "nlsb_make_laplacian_2d_o5": (_I, [_I, _F, _P, _P]),
    "nlsb_make_laplacian_2d_o7": (_I, [_I, _F, _P, _P]),
    "nlsb_rgbmv": (_I, [_P, _P, _F, _P, _I, _I]),
    "nlsb_rbbmv": (_I, [_P, _P, _F, _P, _P, _I, _I]),
    "nlsb_rbbmv_o3": (_I, [_P, _P, _F, _P, _P, _I]),
    "nlsb_rbbmv_o5": (_I, [_P, _P, _F, _P, _P, _I]),
    "nlsb_rbbmv_o7": (_I, [_P, _P, _F, _P, _P, _I]),
    "nlsb_revervoir": (_I, [_P, _P, _P, _P, _I]),
    "nlsb_revervoir_2d": (_I, [_P, _P, _P, _P, _I]),
    "nlsb_hamiltonian": (_I, [_P, _P, _P, _P, _P, _I, _I]),
    "nlsb_hamiltonian_2d": (_I, [_P, _P, _P, _P, _P, _P, _I, _I]),
    "nlsb_runge_kutta": (_I, [_F, _F, _P, _P, _I, _I, _I, _P, _P, _P]),
    "nlsb_runge_kutta_2d": (_I, [_F, _F, _P, _I, _P, _P, _I, _I, _P, _P, _P]),
    "nlsb_solve_nls": (_I, [_F, _F, _I, _I, _I, _P, _P, _P, _P]),
    "nlsb_solve_nls_1d": (_I, [_F, _F, _I, _I, _I, _P, _P, _P, _P]),
    "nlsb_solve_nls_2d": (_I, [_F, _F, _I, _I, _I, _P, _P, _P, _P]),
    "nlsb_chemical_potential_1d": (_I, [_F, _I, _P, _P, _P, _P]),
    "nlsb_chemical_potential_2d": (_I, [_F, _I, _P, _P, _P, _P]),
    "nlsb_radial_taps": (_I, [_I, _I, _F, _P]),
    "nlsb_band_to_taps": (_I, [_I, _I, _P, _P]),
    "nlsb_cross_weights": (_I, [_I, _F, _P, _P]),
    "nlsb_blocks_to_weights": (_I, [_I, _I, _P, _P, _P, _P]),
    "nlsb_dev_rk4_1d": (_I, [_I, _I, _I, _I, _F, _P, _P, _P, _P, _P]),
    "nlsb_dev_hamiltonian_1d": (_I, [_I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "nlsb_dev_band_matvec_1d": (_I, [_I, _I, _P, _P, _P, _F, _P]),
    "nlsb_dev_rk4_2d_workspace": (_Z, [_I, _I, _I]),
    "nlsb_set_2d_path": (_I, [_I]),
    "nlsb_set_stream_tuning": (_I, [_I, _I, _I]),
    "nlsb_solve_nls_2d_plan": (_I, [_I, _I, _I, _P, _P, _P, _P, _P]),
    "nlsb_dev_rk4_2d": (_I, [_I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    "nlsb_dev_rk4_step_2d_slab": (_I, [_I, _I, _I, _F, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "nlsb_planar_pitch": (_I, [_I]),
    "nlsb_dev_rk4_step_2d_slab_planar": (_I, [_I, _I, _I, _F, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "nlsb_dev_hamiltonian_2d": (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "nlsb_dev_cross_matvec_2d": (_I, [_I, _I, _I, _P, _P, _P, _P, _F, _P]),
    "nlsb_dev_divide_check": (_I, [_Z, _P, _P, _P, _P, _P]),
    "nlsb_dev_reservoir": (_I, [_Z, _P, _P, _P, _P, _P]),
    "nlsb_dev_rk4_2d_plan": (_I, [_I, _I, _I, _I, _P, _P, _P, _P]),
    "nlsb_dev_pumping_profiles": (_I, [_I, _I, _I, _I, _F, _P, _P, _P]),
    "nlsb_peer_alloc": (_I, [_Z, _P]),
    "nlsb_peer_free": (_I, [_P]),
    "nlsb_peer_export": (_I, [_P, _P]),
    "nlsb_peer_open": (_I, [_P, _P]),
    "nlsb_peer_close": (_I, [_P]),
    "nlsb_peer_enable_access": (_I, [_I]),
    "nlsb_dev_halo_exchange": (_I, [_P, _P, _P, _P, _Z, _P, _P, _P, _P, _F, _P]),
    "nlsb_dev_rk4_step_2d_slab_exchange": (_I, [_I, _I, _I, _F, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _P, _I, _P,
                                                _P, _P, _P, _P, _F, _P]),
    "nlsb_dev_halo_status": (_I, [_P, _P, _P]),
    "nlsb_add_kernel_launches": (None, [C.c_ulonglong]),
    "nlsb_dev_diagnostics_scratch": (_Z, [_I]),
    "nlsb_dev_rk4_2d_diag_scratch": (_Z, [_I, _I, _I, _I]),
    "nlsb_dev_rk4_2d_diag": (_I, [_I, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P, _P, _P, _P, _Z, _P, _P, _P]),
    "nlsb_dev_rk4_1d_diag": (_I, [_I, _I, _I, _I, _F, _F, _P, _P, _P, _P, _P, _P, _P]),
    "nlsb_dev_diagnostics_1d": (_I, [_I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P]),
    "nlsb_dev_diagnostics_2d": (_I, [_I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P, _P]),
}


class NativeError(RuntimeError):
    """Raised when an ``nlsb_*`` entry point returns a non-zero status (the f2py module's ``error``)."""

    def __init__(self, name, status, message):
        RuntimeError.__init__(self, "%s failed with status %d: %s" % (name, status, message))
        self.status = status


_handle = None


def library_path():
    return _build.LIB


def load():
    """Return the loaded library, building it first when the .so is absent (needs nvcc)."""
    global _handle
    if _handle is not None:
        return _handle
    path = _build.LIB
    if not os.path.exists(path):
        try:
            _build.build_library()
        except Exception as exc:
            raise ImportError("libnls_b200.so is not built and cannot be built here (%s); "
                              "the engine has no CPU fallback" % exc)
    lib = C.CDLL(path)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)   # AttributeError here means header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _handle = lib
    return lib


def call(name, *args):
    """Invoke an int-returning entry point and raise :class:`NativeError` on failure."""
    lib = load()
    status = getattr(lib, name)(*args)
    if status != 0:
        raise NativeError(name, status, lib.nlsb_last_error().decode("utf-8", "replace"))


def kernel_launches():
    return int(load().nlsb_kernel_launches())


def device_available():
    return bool(load().nlsb_device_available())
