"""The C-ABI library loads, exports every symbol include/nls_b200.h declares, its host-side operator
builders are bit-identical to the dp oracle, and compute entry points fail loudly without a GPU."""

import ctypes as C
import os
import re

import numpy as np
import pytest

from nls_b200 import _lib
from nls_b200.native import nls, error
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nls_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nlsb_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_prototyped():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 40
    for name in names:
        assert hasattr(lib, name), "libnls_b200.so does not export %s" % name
        assert name in _lib.PROTOTYPES, "no ctypes prototype for %s" % name
    assert sorted(_lib.PROTOTYPES) == names


def test_library_is_built_for_sm_100a():
    log = os.path.join(ROOT, "nls_b200", "build.log")
    if os.path.exists(log):
        assert "arch=compute_100a,code=sm_100a" in open(log).read()


def test_version_and_error_channel():
    assert nls.version() == (0, 2, 0)
    with pytest.raises(error) as info:
        nls.make_laplacian(32, 4, 0.1)                 # the reference would return garbage (nls.f90:293-294)
    assert info.value.status == -2 and "order" in str(info.value)
    with pytest.raises(error) as info:
        nls.make_laplacian_2d(4, 7, 0.1)
    assert info.value.status == -3


@pytest.mark.parametrize("order", [3, 5, 7])
@pytest.mark.parametrize("n,h", [(7, 0.01), (40, 0.1), (400, 0.1), (1000, 0.05)])
def test_operator_builders_bit_exact_vs_oracle(order, n, h):
    assert np.array_equal(nls.make_laplacian(n, order, h), O.dp.make_laplacian(n, order, h))
    blocks, orders = nls.make_laplacian_2d(n, order, h)
    oblocks, oorders = O.dp.make_laplacian_2d(n, order, h)
    assert np.array_equal(blocks, oblocks) and np.array_equal(orders, oorders)


def test_band_helpers_match_oracle():
    row = [1.0, 2.0, 3.0, 4.0, 5.0]
    assert np.array_equal(nls.make_banded_matrix(7, row), O.dp.make_banded_matrix(7, row))
    rng = np.random.default_rng(3)
    L1 = np.asfortranarray(rng.standard_normal((5, 9)))
    a = nls.clear_first_row_of_derivative(L1)
    b = L1.copy(order="F")
    O._load().nlso_clear_first_row_of_derivative_dp(C.c_int(9), C.c_int(5), b.ctypes.data_as(C.c_void_p))
    assert np.array_equal(a, b)
    a = nls.divide_derivative_on_radius(0.3, L1)
    b = L1.copy(order="F")
    O._load().nlso_divide_derivative_on_radius_dp(C.c_int(9), C.c_int(5), C.c_double(0.3), b.ctypes.data_as(C.c_void_p))
    assert np.array_equal(a, b)


def test_taps_and_weights_tables():
    lib = _lib.load()
    n, m, h = 50, 5, 0.1
    taps = np.zeros((n, m))
    assert lib.nlsb_radial_taps(n, m, h, taps.ctypes.data_as(C.c_void_p)) == 0
    op = O.dp.make_laplacian(n, m, h)
    dense = np.zeros((n, n))
    for j in range(n):
        for b in range(m):
            i = j + b - 2
            if 0 <= i < n:
                dense[i, j] = op[b, j]
    for i in range(n):
        for t in range(m):
            j = i + t - 2
            assert taps[i, t] == (dense[i, j] if 0 <= j < n else 0.0)
    wx, wy = np.zeros(m), np.zeros(m)
    blocks, orders = O.dp.make_laplacian_2d(n, m, h)
    assert lib.nlsb_blocks_to_weights(n, m, blocks.ctypes.data_as(C.c_void_p), orders.ctypes.data_as(C.c_void_p),
                                      wx.ctypes.data_as(C.c_void_p), wy.ctypes.data_as(C.c_void_p)) == 0
    dx2 = 12 * (h * h)
    assert np.array_equal(wx, np.array([-1, 16, -60, 16, -1]) / dx2)
    assert np.array_equal(wy, np.array([-1, 16, 0, 16, -1]) / dx2)
    blocks[3, 1] *= 2.0      # a line-dependent off-diagonal block is not a cross stencil
    assert lib.nlsb_blocks_to_weights(n, m, blocks.ctypes.data_as(C.c_void_p), orders.ctypes.data_as(C.c_void_p),
                                      wx.ctypes.data_as(C.c_void_p), wy.ctypes.data_as(C.c_void_p)) == -5


def test_compute_fails_loudly_without_a_device():
    if _lib.device_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(error) as info:
        nls.solve_nls(1e-3, 0.1, 5, 1, np.ones(16), np.ones(23), np.ones(16) * 0.1)
    assert info.value.status > 0          # a cudaError_t: no CPU fallback exists


def test_kernel_choice_and_stream_geometry_for_the_baseline_configs():
    """nlsb_dev_rk4_2d_plan (host arithmetic): which 2D kernel the automatic path takes and how the strip-marching
    kernel cuts the grid -- every SM busy, chunk rows + 6k a multiple of the unrolled march, strips covering the
    columns."""
    lib = _lib.load()

    def plan(batch, n, order=5):
        out = [C.c_int() for _ in range(4)]
        assert lib.nlsb_dev_rk4_2d_plan(batch, n, n, order, *[C.byref(v) for v in out]) == 0
        return [v.value for v in out]

    assert plan(1, 512)[:2] == [1, 512]            # C2: one wave of 32x64 tiles
    assert plan(1, 400)[0] in (0, 1) and plan(1, 256)[:2] == [0, 256] and plan(1, 768)[:2] == [0, 256]
    assert plan(1, 513)[0] in (0, 1) and plan(1, 1025)[0] == 2     # odd widths stream too (pitched pumping copy)
    for batch, n, order in ((1, 8192, 5), (256, 1024, 5), (1, 2048, 5), (1, 4096, 3), (1, 4096, 7), (8, 1024, 5)):
        kernel, threads, strips, chunk = plan(batch, n, order)
        k = (order - 1) // 2
        assert kernel == 2 and threads in (128, 192, 256)
        assert strips * (threads - 8 * k) >= n > (strips - 1) * (threads - 8 * k)
        assert (chunk + 6 * k) % (2 * (2 * k + 1)) == 0 and chunk >= 1
        ctas = strips * -(-n // chunk) * batch
        assert ctas >= 140                          # at least ~one CTA per SM
    # order 5: two 128-thread CTAs per SM are faster per swept column than one 256-thread CTA (measured), which
    # outweighs their wider relative halo on the 8192-column grid; a 1024-column member is cut into 4 x 240 + 64 columns
    # (the last strip keeps 3 of its 8 warps: 1120 swept columns) rather than 9 x 112 + 16 (1184) -- measured 6.49e10
    # against 6.17e10 node-steps/s; order 3 keeps 256, order 7 has a single shape
    assert plan(256, 1024)[1:3] == [256, 5] and plan(1, 8192)[1:] == [128, 74, 2048]      # 74 x 4 CTAs: one wave
    assert plan(1, 4096, 3)[1] == 256 and plan(1, 4096, 7)[1] == 192
    # a 1024-row slab of the 8192-column grid with its deep halo (multi-GPU): exactly one wave of 2 x 148 CTAs
    out = [C.c_int() for _ in range(4)]
    assert lib.nlsb_dev_rk4_2d_plan(1, 1072, 8192, 5, *[C.byref(v) for v in out]) == 0
    assert out[2].value * -(-1072 // out[3].value) == 296


def test_array_extents_are_checked_before_the_c_abi_sees_pointers():
    """f2py raises on mismatched extents (check(shape(pumping,0)==n)); the ctypes shim must too -- the C side
    would otherwise copy n elements out of a shorter host buffer."""
    c, z8, z88 = np.ones(23), np.ones(8, complex), np.ones((8, 8), complex)
    orders = np.array([0, 0, 2, 0, 0])
    cases = [
        (nls.hamiltonian, (np.ones(5), c, z8, np.ones((5, 8)))),
        (nls.hamiltonian, (np.ones(8), c, z8, np.ones((5, 7)))),
        (nls.hamiltonian_2d, (np.ones((8, 7)), c, z88, np.ones((8, 9)), orders)),
        (nls.hamiltonian_2d, (np.ones((8, 8)), c, z88, np.ones((8, 7)), orders)),
        (nls.runge_kutta, (1e-3, 0.0, z8, np.ones((5, 8)), 1, np.ones(7), c)),
        (nls.runge_kutta, (1e-3, 0.0, z8, np.ones((5, 9)), 1, np.ones(8), c)),
        (nls.runge_kutta_2d, (1e-3, 0.0, z88, np.ones((8, 9)), orders, 1, np.ones((8, 7)), c)),
        (nls.runge_kutta_2d, (1e-3, 0.0, z88, np.ones((7, 9)), orders, 1, np.ones((8, 8)), c)),
        (nls.chemical_potential_1d, (0.1, np.ones(9), c, z8)),
        (nls.chemical_potential_2d, (0.1, np.ones((8, 9)), c, z88)),
        (nls.revervoir, (np.ones(8), c, np.ones(7))),
        (nls.revervoir_2d, (np.ones((8, 8)), c, np.ones((8, 7)))),
        (nls.rgbmv, (np.ones(7), np.ones(8), 1.0, np.ones((5, 8)))),
        (nls.rgbmv, (np.ones(8), np.ones(8), 1.0, np.ones((5, 9)))),
        (nls.rbbmv, (np.ones(64), np.ones(64), 1.0, np.ones((7, 9)), orders, 8)),
        (nls.solve_nls, (1e-3, 0.1, 5, 1, np.ones(9), c, z8)),
        (nls.solve_nls_2d, (1e-3, 0.1, 5, 1, np.ones((8, 9)), c, z88)),
    ]
    for fn, args in cases:
        with pytest.raises(ValueError):
            fn(*args)


def test_coefficients_outside_the_divide_domain_are_refused():
    """The fused kernels' divide needs a positive, finite reservoir denominator c13 + c14 |psi|^2 (always true for
    model.py's coefficient sets: c13 = 1, c14 > 0); other sets are an argument error, not a silent NaN."""
    good = np.ones(23)
    for bad_index, bad_value in ((12, 0.0), (12, -1.0), (13, -0.5), (11, np.inf), (2, np.nan)):
        c = good.copy()
        c[bad_index] = bad_value
        with pytest.raises(error) as info:
            nls.solve_nls(1e-3, 0.1, 5, 1, np.ones(16), c, np.ones(16) * 0.1)
        assert info.value.status == -1
        with pytest.raises(error):
            nls.solve_nls_2d(1e-3, 0.1, 5, 1, np.ones((8, 8)), c, np.ones((8, 8)) * 0.1)
