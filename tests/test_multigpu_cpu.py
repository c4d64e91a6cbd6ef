"""Host-side logic of the multi-GPU paths on CPU: shard ranges, slab plans, and the slab time loop with
halo exchange over a world_size-2 and -3 gloo group (the kernel is replaced by a numpy stepper)."""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nls_b200.multigpu import SlabGrid2D, SlabPlan, advance_emulated, halo_rows, shard_range
from oracle import oracle as O

sys.path.insert(0, os.path.dirname(__file__))
from helpers_slab import numpy_stepper  # noqa: E402


def test_shard_range_partitions_exactly():
    for total in (1, 7, 256, 65536, 1000):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 3, 3)


def test_slab_plan_geometry():
    assert [halo_rows(o) for o in (3, 5, 7)] == [4, 8, 12]
    plans = [SlabPlan(8192, 5, r, 8) for r in range(8)]
    assert all(p.rows_local == 1024 and p.rows_alloc == 1040 for p in plans)
    assert plans[0].up is None and plans[0].down == 1 and plans[7].down is None
    assert plans[3].global_row0 == 3 * 1024 - 8
    top, bottom, interior = plans[3].strips()
    assert top == (8, 40) and bottom == (1000, 1032) and interior == (40, 1000)
    assert plans[3].send_up() == (8, 16) and plans[3].send_down() == (1024, 1032)
    assert plans[3].recv_from_up() == (0, 8) and plans[3].recv_from_down() == (1032, 1040)
    # strips cover the owned rows exactly once, whatever the slab height
    for n, world in ((100, 3), (64, 2), (40, 4), (129, 2)):
        for r in range(world):
            p = SlabPlan(n, 5, r, world)
            rows = []
            for part in p.strips():
                if part:
                    rows += list(range(*part))
            assert sorted(rows) == list(range(*p.owned))
    with pytest.raises(ValueError):
        SlabPlan(40, 5, 0, 8)          # 5-row slabs cannot feed an 8-row halo


def _problem(n):
    from nls_b200.model import Problem
    from nls_b200.pumping import GaussianRingPumping2D
    m = Problem().model(model="2d", dx=0.1, dt=1e-3, u0=0.1, order=5, num_nodes=n, num_iters=1,
                        pumping=GaussianRingPumping2D(power=20.0, radius=1.5, variation=0.8))
    rng = np.random.default_rng(7)
    u0 = 0.1 + 0.05 * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    return m, u0


def _single_domain(n, iters, order=5):
    m, u0 = _problem(n)
    g = SlabGrid2D(n, m.dx, m.dt, order, m.getPumping(), m.getCoefficients(), u0, stepper=numpy_stepper, rank=0, world=1)
    return g.advance(iters).gather(), m, u0


def test_numpy_stepper_matches_oracle():
    n, iters = 48, 12
    got, m, u0 = _single_domain(n, iters)
    want = O.dp.solve_nls_2d(m.dt, m.dx, 5, iters, m.getPumping(), m.getCoefficients(), u0)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-12


@pytest.mark.parametrize("world", [2, 3, 4])
def test_emulated_slabs_are_bitwise_partition_invariant(world):
    n, iters = 48, 9
    single, m, u0 = _single_domain(n, iters)
    slabs = [SlabGrid2D(n, m.dx, m.dt, 5, m.getPumping(), m.getCoefficients(), u0, stepper=numpy_stepper,
                        rank=r, world=world) for r in range(world)]
    advance_emulated(slabs, iters)
    full = np.concatenate([g.local_solution().numpy() for g in slabs], axis=0)
    assert np.array_equal(full, single)


@pytest.mark.parametrize("world,halo_steps,iters", [(2, 2, 9), (3, 2, 8), (2, 3, 7)])
def test_deep_halo_slabs_are_bitwise_partition_invariant(world, halo_steps, iters):
    """halo_steps = m: one exchange per m steps, the neighbours' rows recomputed on a shrinking region."""
    n = 96
    single, m, u0 = _single_domain(n, iters)
    slabs = [SlabGrid2D(n, m.dx, m.dt, 5, m.getPumping(), m.getCoefficients(), u0, stepper=numpy_stepper,
                        rank=r, world=world, halo_steps=halo_steps) for r in range(world)]
    assert slabs[0].plan.halo == 8 * halo_steps and slabs[0].plan.step_rows(0)[0] == 8 * halo_steps
    advance_emulated(slabs, iters)
    full = np.concatenate([g.local_solution().numpy() for g in slabs], axis=0)
    assert np.array_equal(full, single)
    # continuing after a partial macro step keeps working
    advance_emulated(slabs, 3)
    again, _, _ = _single_domain(n, iters + 3)
    assert np.array_equal(np.concatenate([g.local_solution().numpy() for g in slabs], axis=0), again)


def test_default_halo_steps():
    from nls_b200.multigpu import default_halo_steps
    assert default_halo_steps(8192, 5, 8) == 4 and default_halo_steps(8192, 5, 1) == 1
    assert default_halo_steps(512, 5, 2) == 4 and default_halo_steps(256, 5, 2) == 2 and default_halo_steps(96, 5, 4) == 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, iters, out_dir, halo_steps=1):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m, u0 = _problem(n)
        g = SlabGrid2D(n, m.dx, m.dt, 5, m.getPumping(), m.getCoefficients(), u0, stepper=numpy_stepper,
                       halo_steps=halo_steps)
        assert (g.rank, g.world) == (rank, world)
        full = g.advance(iters).gather()
        if rank == 0:
            np.save(os.path.join(out_dir, "full.npy"), full)
        # ensemble sharding needs no communication: every rank simply owns its range
        lo, hi = shard_range(10, rank, world)
        counts = torch.tensor([hi - lo])
        dist.all_reduce(counts)
        assert int(counts) == 10
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,halo_steps", [(2, 1), (3, 1), (2, 2)])
def test_gloo_slab_run_matches_single_domain_bitwise(tmp_path, world, halo_steps):
    n, iters = 48, 9
    mp.spawn(_worker, args=(world, _free_port(), n, iters, str(tmp_path), halo_steps), nprocs=world, join=True)
    full = np.load(tmp_path / "full.npy")
    single, m, u0 = _single_domain(n, iters)
    assert np.array_equal(full, single)
    want = O.dp.solve_nls_2d(m.dt, m.dx, 5, iters, m.getPumping(), m.getCoefficients(), u0)
    assert np.linalg.norm(full - want) / np.linalg.norm(want) <= 1e-12


# ---- sweep front-end (nls_b200/sweep.py): sharding and gathering over gloo, the engine replaced by a stub --------
def _stub_runner(model, kind, n, dx, dt, order, iters, u0, points, keep_fields):
    return {"tag": np.array([p.pump("power") * 2 + p.pump("radius") for p in points]),
            "c12": np.array([p.coefficients()[11] for p in points])}


def _sweep_points(count):
    return [dict(power=1.0 + i, radius=0.5 * i, gamma_R=0.1 + 0.01 * i) for i in range(count)]


def _sweep_worker(rank, world, port, count, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nls_b200.sweep import run_sweep
        table = run_sweep(_sweep_points(count), num_nodes=32, num_iters=5, runner=_stub_runner)
        assert (table is None) == (rank != 0)
        if rank == 0:
            np.savez(os.path.join(out_dir, "table.npz"), **table)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,count", [(2, 7), (3, 10), (3, 2)])
def test_sweep_shards_points_and_gathers_in_order(tmp_path, world, count):
    from nls_b200.sweep import SweepPoint, run_sweep
    mp.spawn(_sweep_worker, args=(world, _free_port(), count, str(tmp_path)), nprocs=world, join=True)
    table = np.load(tmp_path / "table.npz")
    single = run_sweep(_sweep_points(count), num_nodes=32, num_iters=5, runner=_stub_runner)     # no process group
    assert np.array_equal(table["tag"], single["tag"]) and np.array_equal(table["c12"], single["c12"])
    want = [SweepPoint(p).pump("power") * 2 + SweepPoint(p).pump("radius") for p in _sweep_points(count)]
    assert np.array_equal(table["tag"], np.array(want))
    # per-point coefficients follow nls/model.py:157 (c12 = 1 / (n0 gamma_R)) for the swept gamma_R
    assert len(set(table["c12"])) == count
