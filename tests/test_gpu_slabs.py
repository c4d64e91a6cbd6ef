"""Slab decomposition on the GPU: partition invariance (bitwise) and parity with the oracle.

Several ranks are emulated inside one process on one GPU (``advance_emulated``): same kernel calls,
same halo rows moved, only the transport differs from the NCCL run (tests/run_slab_nccl.py covers that
one under torchrun on 2+ GPUs)."""

import numpy as np
import pytest

from oracle import oracle as O
from test_gpu_parity import model_2d, rel_l2, rough_field

pytestmark = pytest.mark.gpu


def _run(n, iters, world, order=5):
    from nls_b200.multigpu import SlabGrid2D, advance_emulated
    m = model_2d(n, iters, order=order, radius=min(10.0, n * 0.1 / 4))
    u0 = 0.1 + 0.05 * rough_field((n, n), n)
    slabs = [SlabGrid2D(n, m.dx, m.dt, order, m.getPumping(), m.getCoefficients(), u0, rank=r, world=world)
             for r in range(world)]
    advance_emulated(slabs, iters)
    full = np.concatenate([g.local_solution().cpu().numpy() for g in slabs], axis=0)
    return full, m, u0


@pytest.mark.parametrize("order,n,iters", [(5, 256, 25), (3, 96, 40), (7, 192, 15), (5, 130, 21)])
def test_slabs_bitwise_equal_to_single_domain(order, n, iters):
    from nls_b200.engine import Grid2D
    single, m, u0 = _run(n, iters, 1, order)
    for world in (2, 4, 8):
        if n // world < 4 * ((order - 1) // 2):
            continue
        full, _, _ = _run(n, iters, world, order)
        assert np.array_equal(full, single), (order, n, world)
    # ... and to the ordinary single-GPU engine (same kernel, different tiling origin)
    plain = Grid2D(n, m.dx, m.dt, order=order, pumping=m.getPumping(), coeffs=m.getCoefficients(), u0=u0)
    assert np.array_equal(plain.advance(iters).solution()[0], single)
    want = O.dp.solve_nls_2d(m.dt, m.dx, order, iters, m.getPumping(), m.getCoefficients(), u0)
    assert rel_l2(single, want) <= 1e-10


def test_slab_parity_size_of_config_4():
    """parity for the slab config at n=2048 on 1/2/4/8 partitions (short horizon)."""
    single, m, u0 = _run(2048, 3, 1)
    for world in (2, 8):
        full, _, _ = _run(2048, 3, world)
        assert np.array_equal(full, single)
    want = O.dp.solve_nls_2d(m.dt, m.dx, 5, 3, m.getPumping(), m.getCoefficients(), u0)
    assert rel_l2(single, want) <= 1e-10
