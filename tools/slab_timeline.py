#!/usr/bin/env python
"""Where the time of one multi-GPU slab step goes (run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \\
        tools/slab_timeline.py --out gpurun_out/r2_slab_timeline_8gpu.json

CUDA-event timings on every rank (max over ranks reported) of, per cycle of m RK steps + one halo exchange:
  cycle        the product path: the captured graph (m step launches + exchange kernel), back to back
  steps_only   the same m step launches without the exchange (timing only -- halos go stale)
  exchange     the exchange kernel alone, back to back (handshake + 2 x halo rows over NVLink)
  single_gpu   the same number of rows per step on ONE rank's share of a single-GPU launch (no redundancy, no halo)
so that cycle - steps_only = what the exchange and the neighbour skew cost, steps_only - m * single_gpu = redundant
halo rows + wave quantisation of the slab launch.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--cycles", type=int, default=200)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    from bench import build_inputs
    from nls_b200.engine import Grid2D
    from nls_b200.multigpu import SlabGrid2D
    w = build_inputs("c4", n=args.n)
    slab = SlabGrid2D(w["n"], w["dx"], w["dt"], w["order"], w["pumping"][0], w["coeffs"][0], w["u0"][0], device=dev)
    m = slab.plan.halo_steps
    cycle = slab._cycle_steps()

    def timed(fn, reps):
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    slab.advance(cycle * 20)                                      # warm-up, graph captured
    t_cycle = timed(lambda: slab.advance(cycle), args.cycles) * (m / cycle)

    def steps_only():
        p = slab.plan
        for s in range(m):
            src, dst = slab.psi[(slab.cur + s) % 2], slab.psi[(slab.cur + s + 1) % 2]
            slab.stepper(src, dst, slab.pumping, *p.step_rows(s))
    steps_only()
    t_steps = timed(steps_only, args.cycles)
    t_exch = timed(lambda: slab.peer.exchange(slab.cur), args.cycles) if slab.peer is not None else None
    epoch, timeouts = slab.peer.status() if slab.peer is not None else (None, None)
    rows_local, halo = slab.plan.rows_local, slab.plan.halo
    slab.close()

    # the same rows on a single-GPU launch: rank 0 alone advances a (rows_local x n) grid (no halo, no neighbours)
    t_single = None
    if rank == 0:
        g = Grid2D(rows_local, w["dx"], w["dt"], order=w["order"], pumping=w["pumping"][0][:rows_local], coeffs=w["coeffs"][0],
                   u0=w["u0"][0][:rows_local], cols=w["n"])
        g.advance(40)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.advance(400)
        b.record()
        torch.cuda.synchronize()
        t_single = a.elapsed_time(b) / 400
    if rank == 0:
        rec = {"n": args.n, "world": world, "halo_steps": m, "rows_per_rank": rows_local, "halo_rows": halo,
               "exchange": "peer" if t_exch is not None else "nccl",
               "us_per_step": {"cycle": 1e3 * t_cycle / m, "steps_only": 1e3 * t_steps / m,
                               "exchange_kernel_alone_per_step": (1e3 * t_exch / m) if t_exch is not None else None,
                               "single_gpu_same_rows": 1e3 * t_single},
               "us_per_exchange_kernel_alone": 1e3 * t_exch if t_exch is not None else None,
               "exchanges": epoch, "timeouts": timeouts,
               "note": "max over ranks, CUDA events; cycle = captured graph of m steps + exchange kernel"}
        print(json.dumps(rec))
        if args.out:
            with open(args.out, "w") as fh:
                fh.write(json.dumps(rec, indent=1) + "\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
