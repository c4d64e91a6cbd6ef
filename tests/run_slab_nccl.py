#!/usr/bin/env python
"""Distributed check of the slab path (peer-mapped exchange kernel and NCCL isend/irecv); run on a multi-GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/run_slab_nccl.py

Every rank advances its slab with halo exchange over NVLink; rank 0 compares the gathered field with the
single-GPU engine (bitwise) and with the dp oracle (<= 1e-10), and checks ensemble sharding."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from nls_b200.engine import Ensemble1D, Grid2D
    from nls_b200.multigpu import SlabGrid2D, shard_range
    from test_gpu_parity import model_1d, model_2d, rel_l2, rough_field

    ok = True
    # 512: planar slabs + TMA tile kernel; 2048: interleaved slabs + strip-marching kernel (>= 2^20 owned nodes)
    # 2048 twice: device-initiated exchange over peer-mapped memory (default), then host-issued NCCL
    # 4096 (more than two ranks): slabs of 8 ranks still take the strip-marching kernel, i.e. the exchange rides inside
    # the last step launch of every cycle; compared bitwise with one GPU only (the oracle would take minutes)
    cases = [(512, 40, None), (2048, 22, None), (2048, 6, "nccl")] + ([(4096, 13, None)] if world > 2 else [])
    for n, iters, exchange in cases:
        m = model_2d(n, iters, radius=min(10.0, n * 0.1 / 4))
        u0 = 0.1 + 0.05 * rough_field((n, n), 5)
        slab = SlabGrid2D(n, m.dx, m.dt, 5, m.getPumping(), m.getCoefficients(), u0, exchange=exchange)
        full = slab.advance(iters).gather()
        how = "exchange %s%s%s" % (slab.exchange, (" [" + slab.exchange_note + "]") if slab.exchange_note else "",
                                   " inside the step launch" if getattr(slab, "fused_exchange", False) else "")
        if slab.peer is not None:
            epoch, timeouts = slab.peer.status()
            how += ", %d exchanges, %d time-outs, halo steps %d" % (epoch, timeouts, slab.plan.halo_steps)
            if timeouts:
                ok = False
        slab.close()
        if rank == 0:
            from oracle import oracle as O
            single = Grid2D(n, m.dx, m.dt, order=5, pumping=m.getPumping(), coeffs=m.getCoefficients(), u0=u0)
            single = single.advance(iters).solution()[0]
            bitwise = bool(np.array_equal(full, single))
            err = 0.0 if n > 2048 else rel_l2(full, O.dp.solve_nls_2d(m.dt, m.dx, 5, iters, m.getPumping(), m.getCoefficients(), u0))
            print("slabs %d^2 over %d ranks (%s layout, %s): bitwise equal to 1 GPU: %s, rel-L2 vs oracle %.2e"
                  % (n, world, "planar" if slab.planar else "interleaved", how, bitwise, err))
            ok = ok and bitwise and err <= 1e-10

    # ensemble sharding: each rank advances its members; results gathered and compared with one launch
    B, n1 = 16, 400
    m1 = model_1d(n1, 200, power=1.0)
    powers = np.linspace(1.0, 40.0, B)
    P = powers[:, None] * m1.getPumping()[None, :]
    lo, hi = shard_range(B, rank, world)
    mine = Ensemble1D(n1, m1.dx, m1.dt, batch=hi - lo, pumping=P[lo:hi], coeffs=m1.getCoefficients(), u0=0.1)
    part = mine.advance(200).psi
    sizes = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    parts = [torch.empty((s, n1), dtype=part.dtype, device=part.device) for s in sizes]
    dist.all_gather(parts, part) if len(set(sizes)) == 1 else None
    if rank == 0 and len(set(sizes)) == 1:
        whole = Ensemble1D(n1, m1.dx, m1.dt, batch=B, pumping=P, coeffs=m1.getCoefficients(), u0=0.1).advance(200).psi
        same = bool(torch.equal(torch.cat(parts), whole))
        print("ensemble sharded over %d ranks: bitwise equal to one launch: %s" % (world, same))
        ok = ok and same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
