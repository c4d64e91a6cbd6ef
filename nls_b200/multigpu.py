"""Multi-GPU execution on one box: ensemble sharding and 2D slab decomposition.

The reference has no parallelism of any kind; these two partitionings are what
``BASELINE.json`` asks the engine to add:

* **Ensembles** (independent parameter points) shard trivially: rank r advances members
  ``shard_range(batch, r, world)`` with the single-GPU engine -- no communication on the path.
* **One large 2D grid** is cut into slabs along its slow (row) axis.  A fused RK4 step consumes a halo
  of 4k rows (k = (order-1)/2) per side, so neighbours exchange 4k rows ONCE per step (not once per RK
  stage): point-to-point ``isend/irecv`` to rank +-1 through ``torch.distributed`` (NCCL over NVLink on
  GPUs, gloo in the CPU tests).  The rows a neighbour needs are computed first on a side stream and
  sent while the interior rows are still being computed on the main stream.

The arithmetic of a node does not depend on the partition (the kernel treats rows outside the square
as zero and every node follows the same sequence of roundings), so 1/2/4/8-rank results are bitwise
identical -- ``tests/test_multigpu_cpu.py`` and ``tests/test_gpu_slabs.py`` check exactly that.

On GPUs the halo exchange is DEVICE-INITIATED by default (``exchange="peer"``): the psi buffers are peer-mapped
between neighbouring ranks (CUDA IPC over NVLink, ``csrc/peer.cu``), an exchange is one kernel that stores the
boundary rows straight into the neighbours' halo rows and synchronises with them through flags in device memory,
and the cycle "m RK4 steps + exchange" is captured in a CUDA graph -- no host work and no NCCL call on the data
path.  ``exchange="nccl"`` keeps the host-issued grouped isend/irecv (the fallback when peer mapping is refused).

One process per GPU; launch with ``torchrun`` (RANK / LOCAL_RANK / WORLD_SIZE from the environment).
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_range", "halo_rows", "SlabPlan", "SlabGrid2D", "advance_emulated", "advance_emulated_peer", "link_local_peers"]

STRIP_ROWS = 32   # boundary strips are one tile row of the fused kernel


def shard_range(total, rank, world):
    """Contiguous, balanced range [lo, hi) of `total` items owned by `rank` out of `world`."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return rank * total // world, (rank + 1) * total // world


def halo_rows(order):
    """Rows of psi a fused RK4 step reads beyond the rows it writes: 4 stages x stencil half-width."""
    if order not in (3, 5, 7):
        raise ValueError("order must be 3, 5 or 7 (got %r)" % (order,))
    return 4 * ((order - 1) // 2)


class SlabPlan(object):
    """Row partition of an n-row grid over `world` ranks and the local buffer geometry of one rank."""

    def __init__(self, n, order, rank, world, halo_steps=1):
        self.n, self.order, self.rank, self.world = int(n), int(order), int(rank), int(world)
        # halo_steps = m: the halo is deep enough for m RK4 steps between two exchanges (the rows a neighbour
        # owns are recomputed redundantly on a shrinking region: step s of m trusts (m - s) * 4k halo rows)
        self.halo_steps = max(1, int(halo_steps)) if world > 1 else 1
        self.step_halo = halo_rows(order)
        self.halo = self.step_halo * self.halo_steps
        self.row_lo, self.row_hi = shard_range(self.n, rank, world)
        self.rows_local = self.row_hi - self.row_lo
        smallest = min(shard_range(self.n, r, world)[1] - shard_range(self.n, r, world)[0] for r in range(world))
        if world > 1 and smallest < self.halo:
            raise ValueError("slabs of %d rows are thinner than the %d-row halo: use fewer ranks" % (smallest, self.halo))
        self.rows_alloc = self.rows_local + 2 * self.halo
        self.global_row0 = self.row_lo - self.halo            # global row of local row 0
        self.up = rank - 1 if rank > 0 else None              # neighbour holding the rows above mine
        self.down = rank + 1 if rank < world - 1 else None

    # local row ranges ---------------------------------------------------------------------------
    @property
    def owned(self):
        return self.halo, self.halo + self.rows_local

    def step_rows(self, s):
        """Local rows to produce in the s-th step (0-based) after an exchange: the rows whose 4k-row neighbourhood
        is still valid, clipped to the rows inside the global square."""
        dom_lo, dom_hi = max(0, -self.global_row0), min(self.rows_alloc, self.n - self.global_row0)
        margin = (s + 1) * self.step_halo
        return max(margin, dom_lo), min(self.rows_alloc - margin, dom_hi)

    def strips(self):
        """(top strip, bottom strip, interior) as local row ranges; strips hold the rows neighbours need."""
        lo, hi = self.owned
        if self.world == 1:
            return None, None, (lo, hi)
        top = (lo, min(lo + STRIP_ROWS, hi))
        bottom = (max(top[1], hi - STRIP_ROWS), hi)
        interior = (top[1], bottom[0])
        return top, (bottom if bottom[1] > bottom[0] else None), (interior if interior[1] > interior[0] else None)

    def send_up(self):       # my first `halo` owned rows -> the lower halo of the rank above
        lo, _ = self.owned
        return lo, lo + self.halo

    def send_down(self):     # my last `halo` owned rows -> the upper halo of the rank below
        _, hi = self.owned
        return hi - self.halo, hi

    def recv_from_up(self):  # upper halo
        return 0, self.halo

    def recv_from_down(self):
        _, hi = self.owned
        return hi, hi + self.halo


def _cuda_stepper(plan, cols, dx, dt, order, coeffs):
    """The product stepper: one fused RK4 step of a row range through the C ABI, on the planar layout
    (re plane, im plane, c12*P) that the kernel fetches with TMA."""
    from . import _lib
    from .engine import cross_weights
    wx, wy = cross_weights(order, dx)
    coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)

    def step(planes_in, planes_out, cp, row0, row1):
        _lib.call("nlsb_dev_rk4_step_2d_slab_planar", plan.rows_alloc, cols, order, float(dt),
                  wx.ctypes.data_as(C.c_void_p), wy.ctypes.data_as(C.c_void_p), plan.global_row0, plan.n,
                  int(row0), int(row1), C.c_void_p(cp.data_ptr()), coeffs.ctypes.data_as(C.c_void_p),
                  C.c_void_p(planes_in.data_ptr()), C.c_void_p(planes_out.data_ptr()),
                  C.c_void_p(torch.cuda.current_stream().cuda_stream))

    step.keepalive = (wx, wy, coeffs)
    step.planar = True
    return step


def _cuda_stepper_interleaved(plan, cols, dx, dt, order, coeffs):
    """One fused RK4 step of a row range on interleaved complex128 slabs (C ABI ``nlsb_dev_rk4_step_2d_slab``).
    The library sends large row ranges to the strip-marching kernel and thin boundary strips to the tile kernel;
    both perform a node's arithmetic identically, so the mix does not change a bit of the result."""
    from . import _lib
    from .engine import cross_weights
    wx, wy = cross_weights(order, dx)
    coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)

    def step(psi_in, psi_out, pumping, row0, row1):
        _lib.call("nlsb_dev_rk4_step_2d_slab", plan.rows_alloc, cols, order, float(dt),
                  wx.ctypes.data_as(C.c_void_p), wy.ctypes.data_as(C.c_void_p), plan.global_row0, plan.n,
                  int(row0), int(row1), C.c_void_p(pumping.data_ptr()), coeffs.ctypes.data_as(C.c_void_p),
                  C.c_void_p(psi_in.data_ptr()), C.c_void_p(psi_out.data_ptr()),
                  C.c_void_p(torch.cuda.current_stream().cuda_stream))

    step.keepalive = (wx, wy, coeffs)
    step.planar = False
    return step


def default_halo_steps(n, order, world):
    """Steps between two halo exchanges for the CUDA steppers.  One exchange per step costs ~150 us of host time
    (NCCL group + three launches): measured on 8 B200s it bounds the 8192^2 grid at 250 us per step where the
    kernels need 120.  With a 4x deeper halo one launch per step and one exchange per 4 steps remain; the price
    is 4k m (m - 1) redundantly computed rows per m steps (2 % of a 1024-row slab)."""
    if world <= 1:
        return 1
    rows = n // world
    for m in (4, 2):
        if halo_rows(order) * m * 8 <= rows:
            return m
    return 1


def _default_cuda_stepper(plan, cols, dx, dt, order, coeffs):
    # slabs large enough for the strip-marching kernel (api.cu: stream_preferred) use the interleaved layout
    lo, hi = plan.owned
    if cols % 2 == 0 and (hi - lo) * cols >= (1 << 20):
        return _cuda_stepper_interleaved(plan, cols, dx, dt, order, coeffs)
    return _cuda_stepper(plan, cols, dx, dt, order, coeffs)


class _DeviceArray(object):
    """A device allocation the library owns, presented to torch through ``__cuda_array_interface__``."""

    def __init__(self, ptr, shape, typestr):
        self.ptr = int(ptr)
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (self.ptr, False), "version": 2}


class _PeerSlab(object):
    """The peer-mapped memory of one slab: two psi buffers and a flag block from ``nlsb_peer_alloc`` (exportable
    allocations), a private state block, and -- once ``connect`` ran -- the mapped pointers of both neighbours."""

    FLAG_BYTES, STATE_BYTES = 256, 64

    def __init__(self, plan, cols, device):
        from . import _lib
        self.plan, self.cols, self.device = plan, int(cols), device
        self.row_bytes = 16 * self.cols
        self.psi_bytes = self.row_bytes * plan.rows_alloc
        self._own, self._mapped = [], []
        with torch.cuda.device(device):
            self.psi_ptr = [self._alloc(self.psi_bytes), self._alloc(self.psi_bytes)]
            self.flags_ptr = self._alloc(self.FLAG_BYTES)
            self.state_ptr = self._alloc(self.STATE_BYTES)
        self.psi = [torch.as_tensor(_DeviceArray(p, (plan.rows_alloc, self.cols), "<c16"), device=device) for p in self.psi_ptr]
        self.up_psi = self.down_psi = (None, None)
        self.up_flags = self.down_flags = None
        self._lib = _lib

    def _alloc(self, nbytes):
        from . import _lib
        out = C.c_void_p()
        _lib.call("nlsb_peer_alloc", C.c_size_t(nbytes), C.byref(out))
        self._own.append(out.value)
        return out.value

    def handles(self):
        """IPC handles of (psi buffer 0, psi buffer 1, flag block) as bytes."""
        out = []
        for ptr in (self.psi_ptr[0], self.psi_ptr[1], self.flags_ptr):
            h = (C.c_ubyte * 64)()
            self._lib.call("nlsb_peer_export", C.c_void_p(ptr), h)
            out.append(bytes(h))
        return out

    def _open(self, handle):
        h = (C.c_ubyte * 64).from_buffer_copy(handle)
        out = C.c_void_p()
        self._lib.call("nlsb_peer_open", h, C.byref(out))
        self._mapped.append(out.value)
        return out.value

    def connect(self, up_handles, down_handles):
        with torch.cuda.device(self.device):
            if up_handles is not None:
                self.up_psi = (self._open(up_handles[0]), self._open(up_handles[1]))
                self.up_flags = self._open(up_handles[2])
            if down_handles is not None:
                self.down_psi = (self._open(down_handles[0]), self._open(down_handles[1]))
                self.down_flags = self._open(down_handles[2])

    def connect_local(self, up, down):
        """Neighbours living in the same process (single-GPU emulation of several ranks): plain pointers."""
        if up is not None:
            self.up_psi, self.up_flags = tuple(up.psi_ptr), up.flags_ptr
        if down is not None:
            self.down_psi, self.down_flags = tuple(down.psi_ptr), down.flags_ptr

    def exchange(self, buf, timeout_seconds=5.0):
        """Enqueue the halo exchange of psi buffer `buf` (0 / 1) on the current stream."""
        p, rb = self.plan, self.row_bytes
        a_up, _ = p.send_up()
        a_dn, _ = p.send_down()
        up_plan = SlabPlan(p.n, p.order, p.rank - 1, p.world, p.halo_steps) if p.up is not None else None
        dn_plan = SlabPlan(p.n, p.order, p.rank + 1, p.world, p.halo_steps) if p.down is not None else None
        null = C.c_void_p(None)
        self._lib.call(
            "nlsb_dev_halo_exchange",
            C.c_void_p(self.psi_ptr[buf] + a_up * rb) if up_plan else null,
            C.c_void_p(self.up_psi[buf] + up_plan.recv_from_down()[0] * rb) if up_plan else null,
            C.c_void_p(self.psi_ptr[buf] + a_dn * rb) if dn_plan else null,
            C.c_void_p(self.down_psi[buf] + dn_plan.recv_from_up()[0] * rb) if dn_plan else null,
            C.c_size_t(p.halo * self.cols), C.c_void_p(self.state_ptr), C.c_void_p(self.flags_ptr),
            C.c_void_p(self.up_flags) if up_plan else null, C.c_void_p(self.down_flags) if dn_plan else null,
            float(timeout_seconds), C.c_void_p(torch.cuda.current_stream().cuda_stream))

    def step_and_exchange(self, src, dst, pumping, row0, row1, dx, dt, order, coeffs, wx, wy, timeout_seconds=5.0):
        """The last step of a cycle with the halo exchange INSIDE the launch: psi buffer `src` -> `dst` (indices) on
        local rows [row0, row1); the boundary rows go to the neighbours' halo rows from the kernel's store stage."""
        p, rb = self.plan, self.row_bytes
        a_up, _ = p.send_up()
        a_dn, _ = p.send_down()
        up_plan = SlabPlan(p.n, p.order, p.rank - 1, p.world, p.halo_steps) if p.up is not None else None
        dn_plan = SlabPlan(p.n, p.order, p.rank + 1, p.world, p.halo_steps) if p.down is not None else None
        null = C.c_void_p(None)
        self._lib.call(
            "nlsb_dev_rk4_step_2d_slab_exchange", p.rows_alloc, self.cols, int(order), float(dt),
            wx.ctypes.data_as(C.c_void_p), wy.ctypes.data_as(C.c_void_p), p.global_row0, p.n, int(row0), int(row1),
            C.c_void_p(pumping.data_ptr()), coeffs.ctypes.data_as(C.c_void_p),
            C.c_void_p(self.psi_ptr[src]), C.c_void_p(self.psi_ptr[dst]), p.halo,
            a_up, C.c_void_p(self.up_psi[dst] + up_plan.recv_from_down()[0] * rb) if up_plan else null,
            a_dn, C.c_void_p(self.down_psi[dst] + dn_plan.recv_from_up()[0] * rb) if dn_plan else null,
            C.c_void_p(self.state_ptr), C.c_void_p(self.flags_ptr),
            C.c_void_p(self.up_flags) if up_plan else null, C.c_void_p(self.down_flags) if dn_plan else null,
            float(timeout_seconds), C.c_void_p(torch.cuda.current_stream().cuda_stream))

    def status(self):
        """(exchanges completed, waits that timed out) -- synchronises the device."""
        epoch, timeouts = C.c_ulonglong(), C.c_ulonglong()
        with torch.cuda.device(self.device):
            self._lib.call("nlsb_dev_halo_status", C.c_void_p(self.state_ptr), C.byref(epoch), C.byref(timeouts))
        return int(epoch.value), int(timeouts.value)

    def close(self):
        lib = self._lib.load()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for ptr in self._mapped:
                lib.nlsb_peer_close(C.c_void_p(ptr))
            self._mapped = []
            self.psi = []
            for ptr in self._own:
                lib.nlsb_peer_free(C.c_void_p(ptr))
            self._own = []


class SlabGrid2D(object):
    """One n x n grid advanced by `world` ranks, each owning a slab of rows (BASELINE config 4).

    pumping, u0: the FULL (n, n) arrays (every rank slices its rows) or a scalar u0.  `stepper` is for
    the CPU tests only: a callable with the signature of the product stepper above.
    """

    def __init__(self, n, dx, dt, order=5, pumping=None, coeffs=None, u0=0.1, group=None, device=None, stepper=None,
                 rank=None, world=None, halo_steps=None, exchange=None, fused_exchange=True):
        """fused_exchange: let the last step of every cycle carry the halo exchange in its own launch (strip-marching
        kernel; peer / local exchange only).  exchange: "peer" (device-initiated exchange over peer-mapped memory, CUDA graphs; default for the CUDA
        steppers on interleaved slabs), "nccl" (host-issued isend/irecv through torch.distributed), "local" (peer
        kernels between slabs of ONE process, linked afterwards by ``link_local_peers`` -- tests) or "none"
        (no exchange of its own: ``advance_emulated`` copies the halos)."""
        self.group = group
        if rank is not None or world is not None:
            # explicit placement: several slabs emulated inside one process (see advance_emulated)
            self.rank, self.world = int(rank), int(world)
        else:
            self.world = dist.get_world_size(group) if dist.is_initialized() else 1
            self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if halo_steps is None:
            halo_steps = default_halo_steps(int(n), int(order), self.world) if stepper is None else 1
        self.plan = SlabPlan(n, order, self.rank, self.world, halo_steps)
        self.since_exchange = 0          # steps taken since the halos were last refreshed (deep-halo mode)
        self.n, self.dx, self.dt, self.order = int(n), float(dx), float(dt), int(order)
        self.coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
        if self.coeffs.shape != (23,):
            raise ValueError("coeffs must be a vector of 23 reals")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if stepper is None else torch.device("cpu")
        self.device = torch.device(device)
        self.on_gpu = self.device.type == "cuda"
        p = self.plan
        if stepper is None:
            stepper = _default_cuda_stepper
        self.stepper = stepper(p, n, dx, dt, order, self.coeffs)
        self.planar = bool(getattr(self.stepper, "planar", False))
        pumping_buf = torch.zeros((p.rows_alloc, n), dtype=torch.float64, device=self.device)
        psi0 = torch.zeros((p.rows_alloc, n), dtype=torch.complex128, device=self.device)
        self._load(pumping_buf, pumping, with_halo=True)
        self._load(psi0, u0, with_halo=True)
        if self.planar:
            # re plane, im plane and c12*P, rows padded to an even pitch (16-byte aligned rows for TMA);
            # padding columns and rows outside the square are zero and are never written
            from . import _lib
            self.pitch = int(_lib.load().nlsb_planar_pitch(n))
            self.psi = [torch.zeros((2, p.rows_alloc, self.pitch), dtype=torch.float64, device=self.device)
                        for _ in range(2)]
            self.psi[0][0, :, :n] = psi0.real
            self.psi[0][1, :, :n] = psi0.imag
            self.pumping = torch.zeros((p.rows_alloc, self.pitch), dtype=torch.float64, device=self.device)
            self.pumping[:, :n] = float(self.coeffs[11]) * pumping_buf      # c12 * P (nls.f90:580 association)
        else:
            self.psi = [psi0, torch.zeros_like(psi0)]
            self.pumping = pumping_buf
        self.cur = 0
        self.steps_done = 0
        self.side = torch.cuda.Stream(device=self.device) if self.on_gpu else None
        # ---- how halos travel --------------------------------------------------------------------------
        explicit = rank is not None or world is not None
        if exchange is None:
            exchange = "none" if explicit else ("peer" if (self.on_gpu and not self.planar and self.world > 1) else "nccl")
        if exchange in ("peer", "local") and (self.planar or not self.on_gpu):
            raise ValueError("exchange=%r needs the interleaved CUDA stepper" % exchange)
        self.exchange, self.peer, self.graphs, self.exchange_note = exchange, None, {}, ""
        self.use_graphs = True
        self.fused_exchange = False
        self.allow_fused_exchange = fused_exchange
        if exchange in ("peer", "local") and self.world > 1:
            try:
                self._setup_peer(psi0, local=(exchange == "local"))
            except Exception as exc:            # peer mapping refused (no NVLink / IPC disabled): host-issued exchange
                if exchange == "local":
                    raise
                self.exchange, self.peer = "nccl", None
                self.exchange_note = "peer mapping failed (%s); using NCCL isend/irecv" % (str(exc)[:160],)
            if not explicit and dist.is_initialized():
                # every rank must have taken the same decision
                ok = torch.tensor([1 if self.exchange == "peer" else 0], device=self.device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
                if int(ok.item()) == 0 and self.exchange == "peer":
                    self.peer.close()
                    self.peer, self.exchange = None, "nccl"
                    self.psi = [psi0, torch.zeros_like(psi0)]
                    self.exchange_note = "a neighbour could not map peer memory; using NCCL isend/irecv"

    def _setup_peer(self, psi0, local):
        """Move the psi buffers into exportable allocations, exchange IPC handles with the neighbours and map
        theirs; warm the step kernel up (function attributes, tensor maps) so that the cycle can be captured."""
        p = self.plan
        peer = _PeerSlab(p, self.n, self.device)
        peer.psi[0].copy_(psi0)
        self.psi = peer.psi
        self.peer = peer
        if not local:
            mine = peer.handles()
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=self.group)
            peer.connect(everyone[p.up] if p.up is not None else None, everyone[p.down] if p.down is not None else None)
            dist.barrier(group=self.group)
        # one throw-away step 0 -> 1 over the rows the first real step also writes (same values)
        self.stepper(self.psi[0], self.psi[1], self.pumping, *p.step_rows(0))
        torch.cuda.synchronize(self.device)
        # does the last step of a cycle take the strip-marching kernel?  Then it carries the exchange itself.
        from . import _lib
        lo, hi = p.step_rows(p.halo_steps - 1)
        out = [C.c_int() for _ in range(4)]
        _lib.call("nlsb_dev_rk4_2d_plan", 1, hi - lo, self.n, self.order, *[C.byref(v) for v in out])
        self.fused_exchange = bool(self.allow_fused_exchange and out[0].value == 2 and self.n % 2 == 0)

    def _load(self, buf, value, with_halo):
        """Fill a local buffer (owned rows and, from a full array, the halo rows inside the domain)."""
        p = self.plan
        g0 = max(p.global_row0, 0) if with_halo else p.row_lo
        g1 = min(p.global_row0 + p.rows_alloc, p.n) if with_halo else p.row_hi
        l0 = g0 - p.global_row0
        if isinstance(value, (int, float, complex)):
            buf[l0:l0 + (g1 - g0)] = value
            return
        arr = value if isinstance(value, torch.Tensor) else torch.from_numpy(np.asarray(value))
        if tuple(arr.shape) != (p.n, p.n):
            raise ValueError("expected the full (%d, %d) array, got %r" % (p.n, p.n, tuple(arr.shape)))
        buf[l0:l0 + (g1 - g0)] = arr[g0:g1].to(device=buf.device, dtype=buf.dtype)

    # ---- halo exchange -------------------------------------------------------------------------------
    def _exchange(self, buf):
        p = self.plan
        ops = []
        parts = [buf[0], buf[1]] if self.planar else [buf]     # planar: the rows of both planes travel
        for part in parts:
            if p.up is not None:
                a, b = p.send_up()
                ops.append(dist.P2POp(dist.isend, part[a:b], p.up, self.group))
                a, b = p.recv_from_up()
                ops.append(dist.P2POp(dist.irecv, part[a:b], p.up, self.group))
            if p.down is not None:
                a, b = p.send_down()
                ops.append(dist.P2POp(dist.isend, part[a:b], p.down, self.group))
                a, b = p.recv_from_down()
                ops.append(dist.P2POp(dist.irecv, part[a:b], p.down, self.group))
        if ops:
            for work in dist.batch_isend_irecv(ops):
                work.wait()

    # ---- time stepping -----------------------------------------------------------------------------
    # ---- device-initiated exchange (peer-mapped memory): m steps + one exchange kernel, replayed from a graph --------
    def _peer_step(self):
        """One RK4 launch on the rows that are still valid; the m-th step of a cycle also exchanges the halos -- inside
        the same launch when the strip-marching kernel takes it (``fused_exchange``), else by the exchange kernel."""
        p = self.plan
        last = self.since_exchange + 1 == p.halo_steps
        rows = p.step_rows(self.since_exchange)
        if last and self.fused_exchange:
            wx, wy, coeffs = self.stepper.keepalive
            self.peer.step_and_exchange(self.cur, 1 - self.cur, self.pumping, rows[0], rows[1], self.dx, self.dt, self.order,
                                        coeffs, wx, wy)
        else:
            self.stepper(self.psi[self.cur], self.psi[1 - self.cur], self.pumping, *rows)
            if last:
                self.peer.exchange(1 - self.cur)
        self.cur = 1 - self.cur
        self.since_exchange = 0 if last else self.since_exchange + 1
        self.steps_done += 1

    def _cycle_steps(self):
        m = self.plan.halo_steps
        return m if m % 2 == 0 else 2 * m        # a replayed cycle must start and end in the same psi buffer

    def _cycle_graph(self):
        """The CUDA graph of one cycle starting from the current buffer (captured on first use)."""
        from . import _lib
        key = self.cur
        if key not in self.graphs:
            lib = _lib.load()
            before = lib.nlsb_kernel_launches()
            graph = torch.cuda.CUDAGraph()
            state = (self.cur, self.since_exchange, self.steps_done)
            with torch.cuda.graph(graph, capture_error_mode="relaxed"):
                for _ in range(self._cycle_steps()):
                    self._peer_step()
            kernels = lib.nlsb_kernel_launches() - before      # recorded, not run
            lib.nlsb_add_kernel_launches(C.c_ulonglong((1 << 64) - kernels))
            self.cur, self.since_exchange, self.steps_done = state
            self.graphs[key] = (graph, int(kernels))
        return self.graphs[key]

    def prepare(self, iters):
        """Capture the graph(s) an ``advance(iters)`` from the current state would replay, without running anything
        (graph capture synchronises the device: slabs emulated inside one process capture before any of them starts)."""
        if self.peer is not None and self.use_graphs and self.since_exchange == 0 and int(iters) >= self._cycle_steps():
            with torch.cuda.device(self.device):
                self._cycle_graph()
        return self

    def _advance_peer(self, iters):
        from . import _lib
        iters = int(iters)
        cycle = self._cycle_steps()
        with torch.cuda.device(self.device):
            while iters > 0:
                if self.use_graphs and self.since_exchange == 0 and iters >= cycle:
                    graph, kernels = self._cycle_graph()
                    replays = iters // cycle
                    for _ in range(replays):
                        graph.replay()
                    _lib.load().nlsb_add_kernel_launches(C.c_ulonglong(kernels * replays))
                    self.steps_done += replays * cycle
                    iters -= replays * cycle
                else:
                    self._peer_step()
                    iters -= 1
        return self

    def advance(self, iters):
        if self.peer is not None:
            return self._advance_peer(iters)
        p = self.plan
        top, bottom, interior = p.strips()
        for _ in range(int(iters)):
            src, dst = self.psi[self.cur], self.psi[1 - self.cur]
            if p.halo_steps > 1:
                # deep halo: one launch per step on the rows that are still valid, one exchange per halo_steps
                last = self.since_exchange + 1 == p.halo_steps
                lo, hi = p.step_rows(self.since_exchange)
                if last and self.on_gpu and hi - lo > 4 * p.halo:
                    # the step before an exchange: the rows the neighbours need first (side stream), their
                    # transfer over NVLink overlaps the interior rows (main stream)
                    main = torch.cuda.current_stream(self.device)
                    self.side.wait_stream(main)
                    with torch.cuda.stream(self.side):
                        self.stepper(src, dst, self.pumping, lo, lo + p.halo)
                        self.stepper(src, dst, self.pumping, hi - p.halo, hi)
                        self._exchange(dst)
                    self.stepper(src, dst, self.pumping, lo + p.halo, hi - p.halo)
                    main.wait_stream(self.side)
                else:
                    self.stepper(src, dst, self.pumping, lo, hi)
                    if last:
                        self._exchange(dst)
                self.since_exchange = 0 if last else self.since_exchange + 1
            elif self.world == 1:
                self.stepper(src, dst, self.pumping, *interior)
            elif self.on_gpu:
                main = torch.cuda.current_stream(self.device)
                self.side.wait_stream(main)                       # previous step (and its halos) complete
                with torch.cuda.stream(self.side):
                    self.stepper(src, dst, self.pumping, *top)
                    if bottom:
                        self.stepper(src, dst, self.pumping, *bottom)
                    self._exchange(dst)                           # NVLink transfer overlaps the interior below
                if interior:
                    self.stepper(src, dst, self.pumping, *interior)
                main.wait_stream(self.side)
            else:
                self.stepper(src, dst, self.pumping, *top)
                if bottom:
                    self.stepper(src, dst, self.pumping, *bottom)
                if interior:
                    self.stepper(src, dst, self.pumping, *interior)
                self._exchange(dst)
            self.cur = 1 - self.cur
            self.steps_done += 1
        return self

    def compute_step(self):
        """One step of this slab WITHOUT the halo exchange (for in-process emulation of several ranks)."""
        src, dst = self.psi[self.cur], self.psi[1 - self.cur]
        if self.plan.halo_steps > 1:
            self.stepper(src, dst, self.pumping, *self.plan.step_rows(self.since_exchange))
            return dst
        for rows in self.plan.strips():
            if rows:
                self.stepper(src, dst, self.pumping, *rows)
        return dst

    # ---- results -----------------------------------------------------------------------------------
    def local_solution(self):
        lo, hi = self.plan.owned
        buf = self.psi[self.cur]
        if self.planar:
            return torch.complex(buf[0, lo:hi, :self.n], buf[1, lo:hi, :self.n])
        return buf[lo:hi]

    def set_local_state(self, other_buffer):
        """Overwrite the current state buffer (same layout, halo rows included) -- used by bench.py to reset between
        runs.  Only valid right after an exchange (or at the start)."""
        if self.since_exchange != 0:
            raise RuntimeError("set_local_state between two halo exchanges")
        self.psi[self.cur].copy_(other_buffer)

    def upload(self, pumping_local, psi_local):
        """Load this rank's slab from HOST arrays of the local layout (rows_alloc x n, halo rows included; pinned
        memory makes the copies asynchronous): the end-to-end path of bench.py.  Only valid right after an exchange."""
        if self.since_exchange != 0:
            raise RuntimeError("upload between two halo exchanges")
        if self.planar:
            raise NotImplementedError("upload is implemented for interleaved slabs")
        self.pumping.copy_(pumping_local, non_blocking=True)
        self.psi[self.cur].copy_(psi_local, non_blocking=True)
        return self

    def download(self, out_host):
        """Copy the owned rows of the current state to a HOST array (rows_local x n) and wait for it."""
        out_host.copy_(self.local_solution(), non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return out_host

    def close(self):
        """Release peer mappings and graphs (call on every rank before the process group is destroyed)."""
        self.graphs = {}
        if self.peer is not None:
            self.peer.close()
            self.peer = None
            self.psi = []

    def state_buffer(self):
        return self.psi[self.cur]

    def gather(self):
        """The full (n, n) solution as a numpy array on every rank."""
        local = self.local_solution().contiguous()
        if self.world == 1:
            return local.cpu().numpy()
        most = max(shard_range(self.n, r, self.world)[1] - shard_range(self.n, r, self.world)[0]
                   for r in range(self.world))
        padded = torch.zeros((most, self.n), dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
        parts = [torch.empty_like(padded) for _ in range(self.world)]
        dist.all_gather(parts, padded, group=self.group)
        rows = []
        for r, part in enumerate(parts):
            lo, hi = shard_range(self.n, r, self.world)
            rows.append(part[:hi - lo].cpu())
        return torch.cat(rows, dim=0).numpy()


def advance_emulated(slabs, iters):
    """Advance the slabs of ALL ranks inside one process (one GPU or the CPU): every slab computes its
    step, then halos are copied directly.  Same arithmetic and data movement as the distributed run,
    used to check partition invariance where only one device is available."""
    slabs = sorted(slabs, key=lambda g: g.rank)
    for _ in range(int(iters)):
        new = [g.compute_step() for g in slabs]
        for g, buf in zip(slabs, new):
            p = g.plan
            g.since_exchange += 1
            if g.since_exchange < p.halo_steps:
                continue                      # deep halo: the rows still trusted shrink, no exchange yet
            g.since_exchange = 0
            if p.up is not None:
                a, b = p.recv_from_up()
                c, d = slabs[p.up].plan.send_down()
                buf[..., a:b, :] = new[p.up][..., c:d, :]
            if p.down is not None:
                a, b = p.recv_from_down()
                c, d = slabs[p.down].plan.send_up()
                buf[..., a:b, :] = new[p.down][..., c:d, :]
        for g in slabs:
            g.cur = 1 - g.cur
            g.steps_done += 1
    return slabs


def link_local_peers(slabs):
    """Connect slabs built with ``exchange="local"`` inside ONE process (plain pointers instead of IPC mappings)."""
    slabs = sorted(slabs, key=lambda g: g.rank)
    for g in slabs:
        g.peer.connect_local(slabs[g.rank - 1].peer if g.plan.up is not None else None,
                             slabs[g.rank + 1].peer if g.plan.down is not None else None)
    return slabs


def advance_emulated_peer(slabs, iters, streams=None):
    """Advance slabs linked by ``link_local_peers`` with the PRODUCT exchange kernel: every slab runs its own
    ``advance`` on its own stream (the exchange kernels of neighbouring slabs wait for each other on the device, so
    they must be able to run concurrently).  One GPU stands in for several."""
    slabs = sorted(slabs, key=lambda g: g.rank)
    dev = slabs[0].device
    if streams is None:
        streams = [torch.cuda.Stream(device=dev) for _ in slabs]
    main = torch.cuda.current_stream(dev)
    for g in slabs:
        g.prepare(iters)
    for g, st in zip(slabs, streams):
        st.wait_stream(main)
        with torch.cuda.stream(st):
            g.advance(iters)
    for st in streams:
        main.wait_stream(st)
    return slabs
