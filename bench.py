#!/usr/bin/env python
"""bench.py -- grid-point RK-steps per second of the daskol/nls hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference] [--workload c4] [--also c2,c3,c5]

Headline workload (default, every N): BASELINE.json config 4 (C4) -- ONE 8192 x 8192 complex128 grid, ring pump,
order 5.  A bench "step" is one pass of the hot path over that grid: `rk_steps_per_bench_step` (200)
RK4 steps.  With --gpus 1 the grid lives on one GPU (strip-marching kernel, one launch per RK step); with --gpus N > 1
the SAME grid is cut into N row slabs (strong scaling) and the halos travel over NVLink on the data path: a
device-initiated exchange kernel over peer-mapped memory inside a CUDA graph (nls_b200/multigpu.py, csrc/peer.cu),
or NCCL isend/irecv when peer mapping is refused -- `config.partition` says which.

The other BASELINE configs ride in the same JSON line as sub-records under "also" (each with value, roofline
fraction, clocks): C2 (512^2, 5000 steps, replicas), C3 (65 536 x 1D n=1000, full 10 000-step horizon, members
sharded over the ranks), C5 (256 x 1024^2, members sharded).  `--workload cN` makes any of them the headline
instead (profiling / DESIGN.md tables).

Output: ONE JSON line on rank 0.  `value` is device-timed (CUDA events on the launching stream, max over ranks) with
inputs resident in HBM; `e2e` goes through the reference-facing entry point ``nls_b200.native.nls.solve_nls_2d``
(C ABI nlsb_solve_nls_2d) with pinned HOST buffers, H2D + D2H inside the timed region (slabs: every rank uploads and
downloads its rows through ``SlabGrid2D.upload / download``).  `--impl reference` times the CPU oracle (``oracle/``:
the C restatement of nls.f90 in the reference's own single precision -- the reference's Fortran cannot be compiled
in this image) on a bounded sample of the same workload.
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

DOMINANT_KERNEL = {"c1": "rk4_1d_resident", "c3": "rk4_1d_resident", "c2": "rk4_step_fused_kernel",
                   "c4": "rk4_stream_kernel", "c5": "rk4_stream_kernel"}
KERNEL_NOTE = {"c1": "whole time loop in one launch", "c3": "whole time loop in one launch",
               "c2": "TMA tile kernel, 32x64 tiles, one RK4 step per launch",
               "c4": "strip-marching kernel, one RK4 step per launch", "c5": "strip-marching kernel, one RK4 step per launch"}

ORIG = dict(R=0.0242057488654, gamma=0.0242057488654, g=0.00162178517398, tilde_g=0.0169440242057,
            gamma_R=0.242057488654)

WORKLOADS = {
    # name: dim, n, batch, iters = RK steps per bench step, description
    "c1": dict(dim=1, n=400, batch=1, iters=10000, desc="examples/solve1d.py: 1D radial n=400, ring pump, 10000 RK steps, order 5"),
    "c2": dict(dim=2, n=512, batch=1, iters=5000, desc="examples/solve2d.py at 512x512: 2D ring pump, 5000 RK steps, order 5"),
    "c3": dict(dim=1, n=1000, batch=65536, iters=10000, desc="1D ensemble 65536 members (256 powers x 256 reservoir rates), n=1000, 10000 RK steps, order 5"),
    "c4": dict(dim=2, n=8192, batch=1, iters=200, desc="2D 8192x8192 complex128 grid, ring pump radius 200 var 50, order 5, 200 RK steps per pass"),
    "c5": dict(dim=2, n=1024, batch=256, iters=100, desc="2D ensemble 256 x 1024x1024, pump radius 2..40, order 5, 100-step window of the 1e5-step horizon"),
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
    """SM clock, power and throttle reasons of one GPU sampled DURING the timed region: an in-process NVML thread
    (every 20 ms; the timed regions of the multi-GPU runs last a fraction of a second), or -- without pynvml -- the
    recipe's `nvidia-smi ... -lms` subprocess."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
    NVML_BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index, period_s=0.02):
        self.rows, self.proc, self.index, self.period = [], None, index, period_s
        self.thread, self.stop, self.source = None, threading.Event(), None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip() != ""]
            if self.index < len(ids) and ids[self.index].strip().isdigit():
                return int(ids[self.index])
        return self.index

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            smax = pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM)

            def pump():
                while not self.stop.is_set():
                    try:
                        mhz = pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
                        watts = pynvml.nvmlDeviceGetPowerUsage(handle) / 1e3
                        bits = pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle)
                        self.rows.append([str(mhz), str(smax), str(watts)] +
                                         ["Active" if bits & self.NVML_BITS[n] else "Not Active" for n in self.NAMES])
                    except Exception:
                        pass
                    self.stop.wait(self.period)
            self.thread = threading.Thread(target=pump, daemon=True)
            self.thread.start()
            self.source = "nvml thread, %d ms" % int(1e3 * self.period)
            return self
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self._physical_index()), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            self.source = "nvidia-smi -lms 50"
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([f.strip() for f in line.split(",")])

    def __exit__(self, *exc):
        self.stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)

    def summary(self):
        sm, smax, watts, reasons = [], [], [], set()
        for r in list(self.rows):
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                watts.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(self.NAMES, r[3:7]):
                if val == "Active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm), "sm_mhz_min": float(min(sm)), "power_w_max": float(max(watts)), "source": self.source}


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_record(name):
    """What the committed ncu capture of this workload's dominant kernel says (profiles/ncu_index.json, written by
    tools/ncu_index.py from the .ncu-rep files; every entry names its capture file and the commit it was taken at).
    Archived evidence, not something this run measured."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_index.json")) as fh:
            return json.load(fh).get(name)
    except Exception:
        return None


# ---- reference arm: the CPU oracle ------------------------------------------------------------------
CPU_RATE_GUESS = 5.5e6      # sp oracle, point-steps/s on one core, out-of-cache 2D grids (sizes the samples only)


def cpu_sample(w, kind="sp", budget_s=15.0):
    """Times the oracle on a bounded sample of workload `w`: ONE member; 2D grids larger than 4096^2 are sampled on
    their central 4096 x 4096 crop (same pumping values, far out of cache like the full grid); as many RK steps as
    fit `budget_s` seconds at the oracle's ~5e6 point-steps/s (at least 1, at most the workload's own horizon)."""
    from oracle import oracle as O
    k = O.sp if kind == "sp" else O.dp
    P, c, u0 = w["pumping"][0], w["coeffs"][0], w["u0"][0]
    n = w["n"]
    crop = ""
    if w["dim"] == 2 and n > 4096:
        lo = (n - 4096) // 2
        P, u0, n = np.ascontiguousarray(P[lo:lo + 4096, lo:lo + 4096]), np.ascontiguousarray(u0[lo:lo + 4096, lo:lo + 4096]), 4096
        crop = "central 4096x4096 crop of "
    per_step = n if w["dim"] == 1 else n * n
    s = int(max(1, min(w["iters"], budget_s * CPU_RATE_GUESS // per_step)))
    fn = k.solve_nls if w["dim"] == 1 else k.solve_nls_2d
    t0 = time.perf_counter()
    fn(w["dt"], w["dx"], w["order"], s, P, c, u0)
    dt = time.perf_counter() - t0
    sample = "%s1 member of %s, %d RK steps (%d point-steps), %s oracle, 1 thread" % (crop, w["name"], s, per_step * s, kind)
    return per_step * s / dt, dt, sample


def cpu_all_cores(w, budget_s=6.0):
    """What the box's host cores deliver when EVERY core runs its own independent grid with the serial sp oracle (the
    way schedulr farms parameter points out; one grid cannot use more than one thread in the reference): one 1024^2
    crop (or the whole member if smaller / 1D) per core, the same number of RK steps each.  Reported beside the serial
    figure, never instead of it."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    P, c, u0, n = w["pumping"][0], w["coeffs"][0], w["u0"][0], w["n"]
    if w["dim"] == 2 and n > 1024:
        lo = (n - 1024) // 2
        P, u0, n = np.ascontiguousarray(P[lo:lo + 1024, lo:lo + 1024]), np.ascontiguousarray(u0[lo:lo + 1024, lo:lo + 1024]), 1024
    per_step = n if w["dim"] == 1 else n * n
    s = int(max(1, min(w["iters"], budget_s * CPU_RATE_GUESS // per_step)))
    fn = O.sp.solve_nls if w["dim"] == 1 else O.sp.solve_nls_2d
    jobs = [(np.array(P), np.array(u0)) for _ in range(cores)]          # private copies: no shared pages
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as pool:                            # ctypes releases the GIL during the call
        list(pool.map(lambda j: fn(w["dt"], w["dx"], w["order"], s, j[0], c, j[1]), jobs))
    dt = time.perf_counter() - t0
    return {"value": cores * per_step * s / dt, "unit": UNIT, "cores": cores,
            "sample": "%d independent %s grids (one per core), %d RK steps each, sp oracle" %
                      (cores, "%dx%d" % (n, n) if w["dim"] == 2 else "n=%d" % n, s)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = build_inputs(args.workload, args.iters, 1 if WORKLOADS[args.workload]["batch"] > 1 else None)
    # the whole run (K timed + W warm-up samples) stays within a few minutes: at most ~150 s of timed samples
    budget = max(2.0, min(20.0, 150.0 / max(args.steps, 1)))
    for _ in range(args.warmup):
        cpu_sample(w, "sp", budget_s=0.3)
    total_pts, total_t, sample = 0.0, 0.0, ""
    for _ in range(args.steps):
        rate, dt, sample = cpu_sample(w, "sp", budget_s=budget)
        total_pts += rate * dt
        total_t += dt
    value = total_pts / total_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 and args.workload in ("c3", "c4", "c5") else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "name": w["name"],
                   "note": "the reference algorithm is serial (no threads in nls.f90): one bounded sample per step, "
                           "normalised per point-step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count(), "all_cores_independent_grids": cpu_all_cores(w)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---- engine arm -------------------------------------------------------------------------------------
class Env(object):
    """Process placement and the torch.distributed plumbing of one bench run."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl engine needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        assert self.world == args.gpus, "launch with torchrun --nproc-per-node == --gpus"
        self.flush = None

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def flush_l2(self):
        if self.flush is None:
            self.flush = self.torch.empty(256 * 1024 * 1024, dtype=self.torch.uint8, device=self.dev)   # > 126 MB L2
        self.flush.zero_()

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def gather_objects(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out


def merge_clocks(env, mine):
    """Rank 0's record plus the worst case over the ranks (lowest median clock, union of the reasons)."""
    everyone = env.gather_objects(mine)
    out = dict(mine)
    if len(everyone) > 1:
        meds = [c["sm_mhz"] for c in everyone if c.get("sm_mhz")]
        out["sm_mhz_min_over_ranks"] = min(meds) if meds else None
        out["reasons"] = sorted(set(r for c in everyone for r in c.get("reasons", [])))
        out["samples_per_rank"] = [c.get("samples", 0) for c in everyone]
    return out


def measure(env, args, name, steps, warmup, iters=None, warmup_iters=None, with_e2e=False):
    """Device-timed throughput of workload `name` on env.world GPUs; returns the record (dict) on every rank."""
    import torch
    from nls_b200 import _lib
    from nls_b200.engine import Ensemble1D, Grid2D
    spec = WORKLOADS[name]
    world, rank, dev = env.world, env.rank, env.dev
    shard = spec["batch"] > 1                               # ensembles: members sharded over the ranks
    total_batch = args.batch if (args.batch and name == args.workload) else spec["batch"]
    members = (rank * total_batch // world, (rank + 1) * total_batch // world) if shard else None
    w = build_inputs(name, iters, total_batch if shard else None, args.n if name == args.workload else None, members)
    if args.order and name == args.workload:
        w["order"] = args.order
        w["desc"] += " [order overridden: %d]" % args.order
    if shard:
        w["batch"] = members[1] - members[0]
    iters = w["iters"]
    slabs = name == "c4" and world > 1     # one big grid: slab decomposition with halo exchange (strong scaling)
    if slabs:
        from nls_b200.multigpu import SlabGrid2D
        eng = SlabGrid2D(w["n"], w["dx"], w["dt"], w["order"], w["pumping"][0], w["coeffs"][0], w["u0"][0], device=dev,
                         exchange=args.exchange, halo_steps=args.halo_steps, fused_exchange=args.fused_exchange != "off")
        state = lambda: eng.state_buffer()
    elif w["dim"] == 1:
        eng = Ensemble1D(w["n"], w["dx"], w["dt"], order=w["order"], batch=w["batch"], pumping=w["pumping"],
                         coeffs=w["coeffs"], u0=w["u0"], device=dev)
        state = lambda: eng.psi
    else:
        eng = Grid2D(w["n"], w["dx"], w["dt"], order=w["order"], batch=w["batch"], pumping=w["pumping"],
                     coeffs=w["coeffs"], u0=w["u0"], device=dev)
        state = lambda: eng.psi
    # Grids far larger than the 126 MB L2 stream from HBM whatever ran before: the time loop simply continues from
    # step to step.  Small working sets are reset to the initial field and the L2 is flushed between timed steps.
    local_bytes = BYTES_PER_POINT_STEP * (points(w) // world if slabs else points(w))   # psi in + psi out + P
    resident = local_bytes > 2 * 126e6
    psi0 = None if resident else state().clone()

    def one_step(n_iters, timed):
        if not resident:
            state().copy_(psi0)
            env.flush_l2()
        a = b = None
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        eng.advance(n_iters)
        if timed:
            b.record()
        return a, b

    for _ in range(warmup):
        one_step(warmup_iters or iters, False)
    env.barrier()
    launches0 = _lib.kernel_launches()
    with ClockSampler(env.local) as clocks:
        t_wall0 = time.perf_counter()
        events = [one_step(iters, True) for _ in range(steps)]
        env.barrier()
        t_wall = time.perf_counter() - t_wall0
    launches = _lib.kernel_launches() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in events)
    dev_ms_max, launches_max = env.max_over_ranks([dev_ms, float(launches)])

    n_side = w["n"]
    per_member = n_side if spec["dim"] == 1 else n_side * n_side
    total_points = total_batch * per_member if shard else per_member * (1 if slabs else world)
    points_per_gpu = per_member // world if slabs else points(w)
    value = float(total_points) * iters * steps / (dev_ms_max * 1e-3)
    peak, peak_src = measured_hbm_peak()
    achieved = BYTES_PER_POINT_STEP * (value / world) / 1e9
    rk_per_launch = iters if w["dim"] == 1 else 1
    dominant_launches = steps * (1 if w["dim"] == 1 else iters)
    if slabs:
        partition = ("row slabs (strong scaling), %d-row halos, one exchange per %d RK steps, %s; %.1f %% of a rank's rows "
                     "recomputed redundantly between exchanges"
                     % (eng.plan.halo, eng.plan.halo_steps,
                        {"peer": ("device-initiated exchange over peer-mapped memory (NVLink) carried by the last step launch of "
                                  "every cycle (boundary rows stored into the neighbours' halo rows by the step kernel's store "
                                  "stage, READY / DATA flags in device memory), m-step cycle in a CUDA graph"
                                  if getattr(eng, "fused_exchange", False) else
                                  "device-initiated exchange kernel over peer-mapped memory (NVLink), m-step cycle in a CUDA graph"),
                         "nccl": "host-issued NCCL isend/irecv"}[eng.exchange] + ((" [" + eng.exchange_note + "]") if eng.exchange_note else ""),
                        100.0 * eng.plan.step_halo * (eng.plan.halo_steps - 1) / max(eng.plan.rows_local, 1)))
    else:
        partition = "members sharded across ranks, no data-path collective" if shard else "replicas only"
    rec = {
        "value": value, "unit": UNIT, "ms_per_step": dev_ms_max / steps, "steps": steps, "warmup": warmup,
        "scaling": "strong" if (shard or slabs) else "weak",
        "config": {"workload": w["desc"], "name": name, "rk_steps_per_bench_step": iters, "points_per_gpu": points_per_gpu,
                   "partition": partition, "kernels_2d": args.path,
                   "l2": ("per-GPU working set of %.0f MB (psi in, psi out, pumping) > 2 x the 126 MB L2: every step streams "
                          "from HBM, the time loop continues from step to step" % (local_bytes / 1e6)) if resident else
                         "field reset and a 256 MiB buffer written (L2 flush) between timed steps; the working set is L2 / "
                         "SM-resident by nature",
                   "timing": "CUDA events on the launching stream per step, summed; max over ranks"},
        "gpu_launches": int(launches_max),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "kernel": DOMINANT_KERNEL[name], "kernel_note": KERNEL_NOTE[name],
                     "algorithmic_bytes_per_launch": BYTES_PER_POINT_STEP * points_per_gpu * rk_per_launch,
                     "avg_launch_us": 1e3 * dev_ms_max / dominant_launches,
                     "algorithmic_bytes": "40 B per point-step x point-steps per launch (DESIGN.md 3)",
                     "per": "GPU", "launches_in_timed_region": int(launches_max)},
        "wall_s": t_wall,
    }
    if slabs:
        rec["roofline"]["avg_launch_us_note"] = "step time of the slowest rank / RK steps: includes its share of the exchange kernel"
        epoch, timeouts = eng.peer.status() if eng.peer is not None else (None, 0)
        rec["config"]["halo_exchanges_total"] = epoch
        rec["config"]["exchange_in_step_launch"] = bool(getattr(eng, "fused_exchange", False))
        rec["config"]["halo_wait_timeouts"] = timeouts
    archived = ncu_record(name) if (args.path == "auto" and not slabs) else None
    rec["roofline"]["traffic"] = archived.get("dram_bytes_per_launch") if archived else None
    if archived:
        rec["roofline"]["ncu_archived"] = archived
    rec["clocks"] = merge_clocks(env, clocks.summary())

    # ---- end to end: host buffers through the reference-facing entry point (H2D + solve + D2H per step) ----
    if with_e2e:
        e2e = None
        if slabs:
            p = eng.plan
            lo, hi = max(p.global_row0, 0), min(p.global_row0 + p.rows_alloc, p.n)
            P_h = torch.zeros((p.rows_alloc, p.n), dtype=torch.float64).pin_memory()
            u_h = torch.zeros((p.rows_alloc, p.n), dtype=torch.complex128).pin_memory()
            P_h[lo - p.global_row0:hi - p.global_row0] = torch.from_numpy(w["pumping"][0][lo:hi])
            u_h[lo - p.global_row0:hi - p.global_row0] = torch.from_numpy(w["u0"][0][lo:hi])
            out_h = torch.empty((p.rows_local, p.n), dtype=torch.complex128).pin_memory()
            for _ in range(2):
                eng.upload(P_h, u_h).advance(iters).download(out_h)
            env.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                eng.upload(P_h, u_h).advance(iters).download(out_h)
            env.barrier()
            e2e_s = time.perf_counter() - t0
            e2e = (e2e_s, P_h.nbytes + u_h.nbytes, out_h.nbytes, None,
                   "nls_b200.multigpu.SlabGrid2D.upload / advance / download (pinned host slabs per rank)")
        elif name in ("c1", "c2", "c4"):
            from nls_b200.native import nls
            P_h = torch.from_numpy(np.ascontiguousarray(w["pumping"][0])).pin_memory().numpy()
            u_h = torch.from_numpy(np.ascontiguousarray(w["u0"][0])).pin_memory().numpy()
            c_h = w["coeffs"][0]
            fn = nls.solve_nls if w["dim"] == 1 else nls.solve_nls_2d
            for _ in range(2):
                out = fn(w["dt"], w["dx"], w["order"], iters, P_h, c_h, u_h)
            env.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                if not resident:
                    env.flush_l2()
                out = fn(w["dt"], w["dx"], w["order"], iters, P_h, c_h, u_h)
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - t0
            check = bool(np.array_equal(out, state_after_one(env, args, w, iters)))
            e2e = (e2e_s, P_h.nbytes + u_h.nbytes + c_h.nbytes, out.nbytes, check,
                   "nls_b200.native.nls.solve_nls%s (C ABI nlsb_solve_nls%s, pinned host buffers)"
                   % (("", "") if w["dim"] == 1 else ("_2d", "_2d")))
        if e2e:
            e2e_s_max = env.max_over_ranks([e2e[0]])[0]
            rec["e2e"] = {"value": float(total_points) * iters * steps / e2e_s_max, "unit": UNIT,
                          "h2d_bytes_per_step": int(e2e[1]), "d2h_bytes_per_step": int(e2e[2]), "api": e2e[4],
                          "bytes_are": "per rank" if slabs else "per call", "ms_per_step": 1e3 * e2e_s_max / steps}
            if e2e[3] is not None:
                rec["e2e"]["matches_device_run"] = e2e[3]
    if slabs:
        eng.close()
    return rec


def state_after_one(env, args, w, iters):
    """The device-resident engine's result for ONE pass from the initial field (what a host-buffer solve must equal)."""
    from nls_b200.engine import Ensemble1D, Grid2D
    if w["dim"] == 1:
        eng = Ensemble1D(w["n"], w["dx"], w["dt"], order=w["order"], batch=1, pumping=w["pumping"], coeffs=w["coeffs"],
                         u0=w["u0"], device=env.dev)
    else:
        eng = Grid2D(w["n"], w["dx"], w["dt"], order=w["order"], batch=1, pumping=w["pumping"], coeffs=w["coeffs"],
                     u0=w["u0"], device=env.dev)
    return eng.advance(iters).psi[0].cpu().numpy()


def run_engine(args):
    from nls_b200.engine import set_2d_path
    set_2d_path(args.path)
    env = Env(args)
    name = args.workload
    rec = measure(env, args, name, args.steps, args.warmup, iters=args.iters, warmup_iters=args.warmup_iters, with_e2e=True)
    line = {"metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": env.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": rec["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": rec["config"],
            "gpu_launches": rec["gpu_launches"], "roofline": rec["roofline"], "wall_s": rec["wall_s"]}
    if "e2e" in rec:
        line["e2e"] = rec["e2e"]
    line["clocks"] = rec["clocks"]

    also = [] if args.also == "none" else [a for a in args.also.split(",") if a and a != name]
    if args.also == "default":
        also = [a for a in (["c2", "c3", "c5"] if env.world == 1 else ["c3", "c5"]) if a != name]
    line["also"] = {}
    for other in also:
        # shorter warm-up passes for the whole-horizon 1D launch; one timed pass of C3 is 6 s on one GPU
        # sharded sub-records get more timed passes as the ranks' shares shrink: >= 1 s per timed region, so that the
        # 20 ms clock sampler sees it
        sub_steps = {"c2": 3, "c3": 1 if env.world < 4 else 2, "c5": 3 * env.world, "c1": 3, "c4": 2}[other]
        warm_iters = {"c3": 500}.get(other)
        sub = measure(env, args, other, sub_steps, 3, warmup_iters=warm_iters)
        line["also"][other] = {"value": sub["value"], "unit": UNIT, "frac": sub["roofline"]["frac"],
                               "ms_per_step": sub["ms_per_step"], "steps": sub["steps"], "warmup": sub["warmup"],
                               "scaling": sub["scaling"], "config": sub["config"], "roofline": sub["roofline"],
                               "gpu_launches": sub["gpu_launches"], "clocks": sub["clocks"]}

    if env.rank == 0:
        if env.world == 1 and not args.no_cpu:
            wc = build_inputs(name, args.iters, 1 if WORKLOADS[name]["batch"] > 1 else None)
            rate, dt_cpu, sample = cpu_sample(wc, "sp", budget_s=15.0)
            rate_dp, _, sample_dp = cpu_sample(wc, "dp", budget_s=4.0)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                                    "host_cores": os.cpu_count(), "dp_value": rate_dp, "dp_sample": sample_dp,
                                    "all_cores_independent_grids": cpu_all_cores(wc)}
        print(json.dumps(line))
    if env.world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["engine", "reference"], default="engine")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c4")
    ap.add_argument("--also", default="default", help="comma list of other workloads measured as sub-records, 'none', or "
                    "'default' (c2,c3,c5 on one GPU; c3,c5 sharded on several)")
    ap.add_argument("--iters", type=int, default=None, help="RK steps per bench step (default: the workload's)")
    ap.add_argument("--warmup-iters", type=int, default=None, help="RK steps per WARM-UP step (default: as the timed steps; "
                    "set it low for full-horizon runs such as --workload c5 --iters 100000 --steps 1)")
    ap.add_argument("--batch", type=int, default=None, help="override the ensemble size")
    ap.add_argument("--grid-n", dest="n", type=int, default=None, help="override the grid size (profiling only)")
    ap.add_argument("--order", type=int, default=None, help="override the stencil order (profiling only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--exchange", choices=["peer", "nccl"], default=None, help="halo transport of the slab run (default: peer)")
    ap.add_argument("--fused-exchange", choices=["on", "off"], default="on", help="slab run: the last step launch of a cycle carries "
                    "the halo exchange (on, default) or a separate exchange kernel follows it (off)")
    ap.add_argument("--halo-steps", type=int, default=None, help="slab run: RK steps between two halo exchanges (default: the plan's)")
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
