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


# ---- device-initiated halo exchange (csrc/peer.cu) ------------------------------------------------------------
def _run_peer(n, iters, world, order, halo_steps, graphs, expect_fused=None):
    """Slabs of ONE process linked by plain pointers, every slab on its own stream, halos moved by the product's
    exchange kernel (ready / data flags in device memory) -- what the ranks of a multi-GPU run do over NVLink."""
    from nls_b200.multigpu import (SlabGrid2D, _cuda_stepper_interleaved, advance_emulated_peer, link_local_peers)
    m = model_2d(n, iters, order=order, radius=min(10.0, n * 0.1 / 4))
    u0 = 0.1 + 0.05 * rough_field((n, n), n)
    slabs = [SlabGrid2D(n, m.dx, m.dt, order, m.getPumping(), m.getCoefficients(), u0, rank=r, world=world,
                        stepper=_cuda_stepper_interleaved, halo_steps=halo_steps, exchange="local", device="cuda")
             for r in range(world)]
    link_local_peers(slabs)
    for g in slabs:
        g.use_graphs = graphs
        if expect_fused is not None:
            assert g.fused_exchange == expect_fused
    advance_emulated_peer(slabs, iters)
    full = np.concatenate([g.local_solution().cpu().numpy() for g in slabs], axis=0)
    status = [g.peer.status() for g in slabs]
    for g in slabs:
        g.close()
    return full, status


@pytest.mark.parametrize("order,n,iters,world,halo_steps,graphs", [
    (5, 256, 24, 2, 1, False), (5, 256, 24, 2, 1, True), (5, 256, 25, 4, 2, True), (5, 256, 26, 4, 4, True),
    (3, 96, 20, 4, 2, True), (7, 192, 12, 2, 2, True), (5, 130, 21, 3, 1, True)])
def test_peer_exchange_kernel_matches_single_domain(order, n, iters, world, halo_steps, graphs):
    single, _, _ = _run(n, iters, 1, order)
    full, status = _run_peer(n, iters, world, order, halo_steps, graphs)
    assert np.array_equal(full, single)
    for epoch, timeouts in status:
        assert timeouts == 0 and epoch == iters // halo_steps


def test_peer_exchange_on_stream_kernel_slabs():
    """Slabs large enough for the strip-marching kernel (2048 x 2048 over 2 slabs, deep halo 4, graph replay): the last
    step of every cycle carries the exchange in its own launch (stores into the neighbour's halo rows + flags)."""
    single, _, _ = _run(2048, 9, 1)
    full, status = _run_peer(2048, 9, 2, 5, 4, True, expect_fused=True)
    assert np.array_equal(full, single)
    assert all(t == 0 and e == 2 for e, t in status)


@pytest.mark.parametrize("order,n,iters,world,halo_steps,graphs", [
    (5, 256, 24, 2, 1, False), (5, 256, 24, 2, 1, True), (5, 384, 26, 4, 4, True), (3, 160, 20, 4, 2, True),
    (7, 288, 12, 2, 2, True), (5, 260, 21, 3, 1, True), (5, 512, 17, 8, 2, True)])
def test_exchange_inside_the_step_launch_matches_single_domain(order, n, iters, world, halo_steps, graphs):
    """The exchange-carrying step (strip-marching kernel forced on small slabs): READY / DATA flags, peer stores from
    the store stage, epoch counter advanced by the launch itself -- ragged slabs, 2..8 ranks, every order."""
    from nls_b200.engine import set_2d_path
    single, _, _ = _run(n, iters, 1, order)
    try:
        set_2d_path("stream")
        full, status = _run_peer(n, iters, world, order, halo_steps, graphs, expect_fused=True)
    finally:
        set_2d_path("auto")
    assert np.array_equal(full, single)
    for epoch, timeouts in status:
        assert timeouts == 0 and epoch == iters // halo_steps


def test_peer_exchange_between_processes():
    """Two processes, two GPUs: IPC-mapped neighbours, exchange over NVLink, CUDA-graph replay (torchrun)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "run_slab_nccl.py")]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-3000:]
    assert "exchange peer" in proc.stdout
