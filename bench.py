#!/usr/bin/env python
"""bench.py -- grid-point RK-steps per second of the daskol/nls hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference] [--workload c2]

A bench "step" is one pass of the hot path over the workload: one ``solve`` of `iters` RK4 steps.
Default workload at every N is BASELINE.json configs[1] (C2: 2D 512 x 512 ring pumping, 5000 RK
steps, order 5).  C2 does not shard ("replicas only"), so with --gpus N every rank
advances its own replica (weak scaling, no data-path collective).  Other workloads (--workload c1,
c3, c4, c5) are for DESIGN.md / profiles, not the driver's bench line.

Output: ONE JSON line on rank 0 (see the keys in `main`).  `value` is device-timed with inputs resident
in HBM; `e2e` goes through the f2py-signature entry point ``nls_b200.native.nls.solve_nls_2d`` with
pinned HOST buffers (H2D + D2H inside the timed region).  `--impl reference` times the CPU oracle
(``oracle/``: the C restatement of nls.f90 in the reference's own single precision -- the reference's
Fortran cannot be compiled in this image) on a bounded sample of the same workload.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "grid-point RK-steps/sec"
UNIT = "point-steps/s"
BYTES_PER_POINT_STEP = 40.0      # read psi 16 + read P 8 + write psi 16

# DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum) from the
# `ncu --set full` captures summarised under profiles/ (round 1, default kernels).  C2's working set is
# L2-resident: the figure is the cold-cache replay ncu measures, in steady state it is ~0.
NCU_TRAFFIC = {
    "c2": (6.32e6, "profiles/r1_ncu_tma64_pdl_c2_512_default.txt (cold L2 under ncu; L2-resident in steady state)"),
    "c4": (2.716e9, "profiles/r1_ncu_stream_c4_8192.txt"),
}
# FP64-pipe utilisation of the dominant kernel (sm__pipe_fp64_cycles_active, same ncu captures): the path is bound
# by FP64 issue + shared-memory traffic, not by HBM -- reported beside the HBM roofline fraction.
NCU_FP64_PIPE_PCT = {"c2": 39.6, "c3": 67.4, "c4": 66.1}
DOMINANT_KERNEL = {"c1": "rk4_1d_resident (whole time loop, one launch)", "c3": "rk4_1d_resident (whole time loop, one launch)",
                   "c2": "rk4_step_fused_kernel (TMA tile kernel, 32x64 tiles, one RK4 step per launch)",
                   "c4": "rk4_stream_kernel (strip-marching kernel, one RK4 step per launch)",
                   "c5": "rk4_stream_kernel (strip-marching kernel, one RK4 step per launch)"}

ORIG = dict(R=0.0242057488654, gamma=0.0242057488654, g=0.00162178517398, tilde_g=0.0169440242057,
            gamma_R=0.242057488654)

WORKLOADS = {
    # name: dim, n, batch, iters (per bench step), description
    "c1": dict(dim=1, n=400, batch=1, iters=10000, desc="examples/solve1d.py: 1D radial n=400, ring pump, 10000 RK steps, order 5"),
    "c2": dict(dim=2, n=512, batch=1, iters=5000, desc="examples/solve2d.py at 512x512: 2D ring pump, 5000 RK steps, order 5"),
    "c3": dict(dim=1, n=1000, batch=65536, iters=1000, desc="1D ensemble 65536 members (256 powers x 256 reservoir rates), n=1000, order 5"),
    "c4": dict(dim=2, n=8192, batch=1, iters=20, desc="2D 8192x8192 ring pump radius 200 var 50, order 5"),
    "c5": dict(dim=2, n=1024, batch=256, iters=20, desc="2D ensemble 256 x 1024x1024, pump radius 2..40, order 5"),
}


def build_inputs(name, iters=None, batch=None, n=None, members=None):
    """Synthetic inputs of the named shape (deterministic).  members = (lo, hi): build only that
    range of an ensemble's members (what one rank owns); w["batch"] stays the size of the whole ensemble."""
    from nls_b200.model import Problem, dimensionless_coefficients
    from nls_b200.pumping import GaussianRingPumping1D, GaussianRingPumping2D
    w = dict(WORKLOADS[name])
    if iters:
        w["iters"] = iters
    if batch:
        w["batch"] = batch
    if n:
        w["n"] = n
        w["desc"] += " [n overridden: %d]" % n
    n, B = w["n"], w["batch"]
    coeffs = dimensionless_coefficients(dict(ORIG))
    if name == "c1":
        m = Problem().model(model="1d", dx=0.1, dt=1e-3, u0=0.1, order=5, num_nodes=n, num_iters=w["iters"],
                            pumping=GaussianRingPumping1D(power=20.0, radius=10.0, variation=3.14))
        w.update(pumping=m.getPumping()[None], coeffs=coeffs[None], u0=m.getInitialSolution().astype(complex)[None])
    elif name == "c2":
        m = Problem().model(model="2d", dx=0.1, dt=1e-3, u0=0.1, order=5, num_nodes=n, num_iters=w["iters"],
                            pumping=GaussianRingPumping2D(power=20.0, radius=10.0, variation=3.14))
        w.update(pumping=m.getPumping()[None], coeffs=coeffs[None], u0=m.getInitialSolution().astype(complex)[None])
    elif name == "c3":
        m = Problem().model(model="1d", dx=0.1, dt=1e-3, u0=0.1, order=5, num_nodes=n, num_iters=w["iters"],
                            pumping=GaussianRingPumping1D(power=1.0, radius=10.0, variation=3.14))
        unit = m.getPumping()
        side = int(round(np.sqrt(B)))
        powers = np.linspace(1.0, 40.0, side)
        gammas = np.geomspace(0.05, 1.0, max(-(-B // side), 1))
        lo, hi = members if members else (0, B)
        idx = np.arange(lo, hi)                       # member i = (power index i // len(gammas), gamma index i % ...)
        C = np.array([dimensionless_coefficients(dict(ORIG, gamma_R=g)) for g in gammas])
        w.update(pumping=powers[idx // len(gammas), None] * unit[None, :], coeffs=C[idx % len(gammas)].copy(),
                 u0=np.full((hi - lo, n), 0.1 + 0j))
    elif name == "c4":
        m = Problem().model(model="2d", dx=0.1, dt=1e-3, u0=0.1, order=5, num_nodes=n, num_iters=w["iters"],
                            pumping=GaussianRingPumping2D(power=20.0, radius=200.0, variation=50.0))
        w.update(pumping=m.getPumping()[None], coeffs=coeffs[None], u0=np.full((1, n, n), 0.1 + 0j))
    elif name == "c5":
        lo, hi = members if members else (0, B)
        radii = np.linspace(2.0, 40.0, B)[lo:hi]
        P = np.empty((hi - lo, n, n))
        for b, r in enumerate(radii):
            m = Problem().model(model="2d", dx=0.1, dt=1e-3, u0=0.1, order=5, num_nodes=n, num_iters=w["iters"],
                                pumping=GaussianRingPumping2D(power=20.0, radius=float(r), variation=3.14))
            P[b] = m.getPumping()
        w.update(pumping=P, coeffs=np.broadcast_to(coeffs, (hi - lo, 23)).copy(), u0=np.full((hi - lo, n, n), 0.1 + 0j))
    w.update(dx=0.1, dt=1e-3, order=5, name=name)
    return w


def points(w):
    return w["batch"] * (w["n"] if w["dim"] == 1 else w["n"] * w["n"])


# ---- clocks -----------------------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([f.strip() for f in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.thread.join(timeout=2)

    def summary(self):
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(self.NAMES, r[3:7]):
                if val == "Active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---- reference arm: the CPU oracle ------------------------------------------------------------------
def cpu_sample(w, kind="sp", budget_points=1.0e8):
    """Times the oracle on a bounded sample of workload `w`: one member, `s` RK steps (about 10-25 s of CPU work
    for the 2D workloads at the oracle's ~5-9e6 point-steps/s; a 1D member's whole horizon is shorter than that)."""
    from oracle import oracle as O
    k = O.sp if kind == "sp" else O.dp
    per_step = w["n"] if w["dim"] == 1 else w["n"] ** 2
    s = int(max(2, min(w["iters"], budget_points // per_step)))
    P, c, u0 = w["pumping"][0], w["coeffs"][0], w["u0"][0]
    fn = k.solve_nls if w["dim"] == 1 else k.solve_nls_2d
    t0 = time.perf_counter()
    fn(w["dt"], w["dx"], w["order"], s, P, c, u0)
    dt = time.perf_counter() - t0
    sample = "1 member of %s, %d RK steps (%d point-steps), %s oracle, 1 thread" % (w["name"], s, per_step * s, kind)
    return per_step * s / dt, dt, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = build_inputs(args.workload, args.iters, 1 if WORKLOADS[args.workload]["batch"] > 1 else None)
    for _ in range(args.warmup):
        cpu_sample(w, "sp", budget_points=2.0e6)
    total_pts, total_t, sample = 0.0, 0.0, ""
    for _ in range(args.steps):
        rate, dt, sample = cpu_sample(w, "sp")
        total_pts += rate * dt
        total_t += dt
    value = total_pts / total_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "name": w["name"],
                   "note": "reference algorithm is serial; one bounded sample per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---- engine arm -------------------------------------------------------------------------------------
def run_engine(args):
    import torch
    import torch.distributed as dist
    from nls_b200 import _lib
    from nls_b200.engine import Ensemble1D, Grid2D
    from nls_b200.native import nls
    from nls_b200.engine import set_2d_path
    set_2d_path(args.path)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl engine needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, "launch with torchrun --nproc-per-node == --gpus"

    name = args.workload
    spec = WORKLOADS[name]
    # ensembles shard their members across ranks (strong scaling); single systems run as replicas (weak)
    shard = spec["batch"] > 1
    batch = None
    if args.batch:
        batch = args.batch
    total_batch = batch if batch else spec["batch"]
    members = (rank * total_batch // world, (rank + 1) * total_batch // world) if shard else None
    w = build_inputs(name, args.iters, batch, args.n, members)      # a rank builds only the members it owns
    if args.order:
        w["order"] = args.order
        w["desc"] += " [order overridden: %d]" % args.order
    if shard:
        w["batch"] = members[1] - members[0]
    iters = w["iters"]
    dev = torch.device("cuda", local)

    slabs = name == "c4" and world > 1     # one big grid: slab decomposition with halo exchange (strong scaling)
    if slabs:
        from nls_b200.multigpu import SlabGrid2D
        eng = SlabGrid2D(w["n"], w["dx"], w["dt"], w["order"], w["pumping"][0], w["coeffs"][0], w["u0"][0], device=dev)
        psi0 = eng.state_buffer().clone()
    elif w["dim"] == 1:
        eng = Ensemble1D(w["n"], w["dx"], w["dt"], order=w["order"], batch=w["batch"], pumping=w["pumping"],
                         coeffs=w["coeffs"], u0=w["u0"], device=dev)
        psi0 = eng.psi.clone()
    else:
        eng = Grid2D(w["n"], w["dx"], w["dt"], order=w["order"], batch=w["batch"], pumping=w["pumping"],
                     coeffs=w["coeffs"], u0=w["u0"], device=dev)
        psi0 = eng.psi.clone()

    def reset():
        eng.set_local_state(psi0) if slabs else eng.psi.copy_(psi0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(timed):
        reset()
        flush.zero_()
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.advance(iters)
            b.record()
            return a, b
        eng.advance(iters)
        return None

    for _ in range(args.warmup):
        one_step(False)
    barrier()
    launches0 = _lib.kernel_launches()
    with ClockSampler(local) as clocks:
        t_wall0 = time.perf_counter()
        events = [one_step(True) for _ in range(args.steps)]
        barrier()
        t_wall = time.perf_counter() - t_wall0
    launches = _lib.kernel_launches() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in events)
    final = (eng.state_buffer() if slabs else eng.psi).clone()

    # end to end: host buffers through the reference-facing entry point (H2D + solve + D2H per step)
    e2e = None
    if name in ("c1", "c2", "c4") and not slabs:
        P_h = torch.from_numpy(np.ascontiguousarray(w["pumping"][0])).pin_memory().numpy()
        u_h = torch.from_numpy(np.ascontiguousarray(w["u0"][0])).pin_memory().numpy()
        c_h = w["coeffs"][0]
        fn = nls.solve_nls if w["dim"] == 1 else nls.solve_nls_2d
        for _ in range(max(1, min(args.warmup, 3))):
            out = fn(w["dt"], w["dx"], w["order"], iters, P_h, c_h, u_h)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            flush.zero_()
            out = fn(w["dt"], w["dx"], w["order"], iters, P_h, c_h, u_h)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e_ok = bool(np.array_equal(out, final[0].cpu().numpy()))
        e2e = (e2e_s, P_h.nbytes + u_h.nbytes + c_h.nbytes, out.nbytes, e2e_ok)

    # max over ranks
    t = torch.tensor([dev_ms, e2e[0] if e2e else 0.0, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_s_max = float(t[0]), float(t[1])

    n_side = args.n if args.n else spec["n"]
    total_points = (total_batch * (n_side if spec["dim"] == 1 else n_side * n_side)) if shard else points(w) * (1 if slabs else world)
    work = float(total_points) * iters * args.steps
    value = work / (dev_ms_max * 1e-3)
    peak, peak_src = measured_hbm_peak()
    per_gpu_rate = value / world
    achieved = BYTES_PER_POINT_STEP * per_gpu_rate / 1e9
    launches_per_gpu = max(int(t[2]), 1)
    # algorithmic bytes of one launch of the dominant kernel and its average duration over the timed region
    rk_per_launch = iters if w["dim"] == 1 else 1
    bytes_per_launch = BYTES_PER_POINT_STEP * (points(w) // world if slabs else points(w)) * rk_per_launch
    dominant_launches = args.steps * (1 if w["dim"] == 1 else iters)
    traffic = NCU_TRAFFIC.get(name, (None, None)) if args.path == "auto" and not slabs else (None, None)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
        "scaling": "strong" if (shard or slabs) else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "name": name, "rk_steps_per_bench_step": iters,
                   "points_per_gpu": (points(w) // world if slabs else points(w)), "partition": ("members sharded across ranks" if shard else
                                 "row slabs, 4k-row halo exchange per RK step over NCCL" if slabs else "replicas only"),
                   "kernels_2d": args.path,
                   "l2": "256 MiB flush buffer written between timed steps; 512^2 working set is L2-resident by nature",
                   "timing": "CUDA events on the launching stream per step, summed; max over ranks"},
        "gpu_launches": int(t[2]),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic[0], "traffic_source": traffic[1], "peak_source": peak_src,
                     "fp64_pipe_pct_ncu": NCU_FP64_PIPE_PCT.get(name) if args.path == "auto" else None,
                     "kernel": DOMINANT_KERNEL[name],
                     "algorithmic_bytes_per_launch": bytes_per_launch,
                     "avg_launch_us": 1e3 * dev_ms_max / dominant_launches,
                     "algorithmic_bytes": "40 B per point-step x point-steps per launch (DESIGN.md 3)",
                     "per": "GPU", "launches_in_timed_region": launches_per_gpu},
        "wall_s": t_wall,
    }
    if e2e:
        line["e2e"] = {"value": float(points(w)) * world * iters * args.steps / e2e_s_max, "unit": UNIT,
                       "h2d_bytes_per_step": int(e2e[1]), "d2h_bytes_per_step": int(e2e[2]),
                       "api": "nls_b200.native.nls.solve_nls%s (C ABI nlsb_solve_nls%s, pinned host buffers)"
                              % (("", "") if w["dim"] == 1 else ("_2d", "_2d")),
                       "matches_device_run": e2e[3]}
    line["clocks"] = clocks.summary()

    if rank == 0:
        if world == 1 and not args.no_cpu:
            wc = build_inputs(name, args.iters, 1 if spec["batch"] > 1 else None)
            rate, dt_cpu, sample = cpu_sample(wc, "sp")
            rate_dp, _, sample_dp = cpu_sample(wc, "dp", budget_points=2.0e7)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                                    "host_cores": os.cpu_count(), "dp_value": rate_dp, "dp_sample": sample_dp}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["engine", "reference"], default="engine")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c2")
    ap.add_argument("--iters", type=int, default=None, help="RK steps per bench step (default: the workload's)")
    ap.add_argument("--batch", type=int, default=None, help="override the ensemble size")
    ap.add_argument("--grid-n", dest="n", type=int, default=None, help="override the grid size (profiling only)")
    ap.add_argument("--order", type=int, default=None, help="override the stencil order (profiling only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--path", choices=["auto", "fused", "fused32", "fused64", "tma32", "tma64", "tma32_persistent", "tma64_persistent", "stream", "resident", "staged"], default="auto", help="2D kernel family")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "engine":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
