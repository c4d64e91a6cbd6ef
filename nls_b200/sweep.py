"""Local multi-GPU parameter sweeps: the ensemble job front-end.

The reference distributes parameter scans through ``schedulr`` -- an HTTP job queue whose workers each run
``Problem().model(..., pumping=GaussianRingPumping2D(power=desc['power'])).solve()`` for one point at a time and
store the result (``schedulr/workr.py:39-107``, ``schedulr/schedulr.py:78-147``).  On one box of B200s the same
scan is ONE batched launch per GPU: the points are cut into contiguous ranges (``multigpu.shard_range``), every
rank builds its members' pumping profiles on its device (``engine.device_pumping``), advances them as an ensemble
(``Ensemble1D`` / ``Grid2D``: no data-path collective) and reduces the diagnostics on the device; only a few
scalars per point travel back.  Rank 0 receives the table of all points.

Launch under ``torchrun`` (NCCL or gloo process group) or in a single process.  ``runner`` exists for the CPU
tests of the sharding / gathering logic: the engine itself has no CPU path.
"""

from __future__ import annotations

import numpy as np
import torch.distributed as dist

from .model import dimensionless_coefficients
from .multigpu import shard_range

__all__ = ["SweepPoint", "run_sweep"]

ORIGINALS = dict(R=0.0242057488654, gamma=0.0242057488654, g=0.00162178517398, tilde_g=0.0169440242057,
                 gamma_R=0.242057488654)

FIELDS = ("chemical_potential", "damping_integral", "particles", "max_density", "max_reservoir")


class SweepPoint(dict):
    """One point of a scan: pumping parameters (``power``, ``radius``, ``variation``, ``x0``, ``y0``) and any of the
    original (dimensional) model parameters of ``nls/model.py:138-160`` (``R``, ``gamma``, ``g``, ``tilde_g``,
    ``gamma_R``) that differs from the defaults."""

    PUMP = dict(power=20.0, radius=10.0, variation=3.14, x0=0.0, y0=0.0)

    def pump(self, key):
        return float(self.get(key, self.PUMP[key]))

    def coefficients(self):
        return dimensionless_coefficients(dict(ORIGINALS, **{k: self[k] for k in ORIGINALS if k in self}))


def _cuda_runner(model, kind, n, dx, dt, order, iters, u0, points, keep_fields):
    """Advance `points` as one ensemble on the current CUDA device; returns {field: array over points}."""
    from .engine import Ensemble1D, Grid2D, device_pumping
    dim = 1 if model == "1d" else 2
    pumping = device_pumping(dim, kind, n, dx, [p.pump("power") for p in points], [p.pump("variation") for p in points],
                             radius=[p.pump("radius") for p in points], x0=[p.pump("x0") for p in points],
                             y0=[p.pump("y0") for p in points])
    coeffs = np.array([p.coefficients() for p in points])
    cls = Ensemble1D if dim == 1 else Grid2D
    ens = cls(n, dx, dt, order=order, batch=len(points), pumping=pumping, coeffs=coeffs, u0=u0)
    ens.advance(iters)
    out = ens.diagnostics()
    if keep_fields:
        out["solution"] = ens.solution()
    return out


def run_sweep(points, model="2d", kind="ring", num_nodes=200, dx=0.1, dt=1e-3, order=5, num_iters=2000, u0=0.1,
              keep_fields=False, group=None, runner=None):
    """Run every point of `points` (dicts / ``SweepPoint``s) for `num_iters` RK4 steps and return, on rank 0, a dict
    of arrays over ALL points in their original order (``None`` on the other ranks): the diagnostics of
    ``Grid2D.diagnostics`` / ``Ensemble1D.diagnostics`` plus, with ``keep_fields``, the final fields.

    The defaults are the fixed model of the reference's worker (200 x 200 nodes, 2000 steps, order 5,
    ``schedulr/workr.py:39-66``).  Each rank advances the contiguous range ``shard_range(len(points), rank, world)``.
    """
    points = [p if isinstance(p, SweepPoint) else SweepPoint(p) for p in points]
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    lo, hi = shard_range(len(points), rank, world)
    mine = points[lo:hi]
    local = None
    if mine:
        fn = runner if runner is not None else _cuda_runner
        local = fn(model, kind, int(num_nodes), float(dx), float(dt), int(order), int(num_iters), u0, mine, keep_fields)
        for key, value in local.items():
            if len(value) != len(mine):
                raise ValueError("runner returned %d values of %r for %d points" % (len(value), key, len(mine)))
    if not distributed:
        return {k: np.asarray(v) for k, v in local.items()} if local else {}
    parts = [None] * world if rank == 0 else None
    dist.gather_object((lo, hi, local), parts, dst=0, group=group)
    if rank != 0:
        return None
    table = {}
    for plo, phi, part in sorted(p for p in parts if p[2] is not None):
        for key, value in part.items():
            table.setdefault(key, []).append(np.asarray(value))
    return {k: np.concatenate(v, axis=0) for k, v in table.items()}
