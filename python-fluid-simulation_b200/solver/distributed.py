"""Slab-partitioned multi-GPU viscosity CG (new work — the reference is single-GPU; SURVEY.md §8 e).

One process per GPU (``torch.distributed``, NCCL).  The global ``nx x ny x nz`` grid is cut into contiguous
x-slabs of cells (x is the storage-slowest axis of the reference's C-order arrays, so a halo is one contiguous
plane).  Each rank holds its slab EXTENDED by one cell towards every existing neighbour and runs the ordinary
single-GPU solver object on that extended grid with a communicator attached
(``fs_visc3d_set_slab``): rows of the overlap cells are owned by the neighbour and are never computed locally,
their ``d`` planes arrive by halo exchange before each operator apply, and the two CG scalars are all-reduced,
so every rank executes the same iteration sequence and stops on the same iteration.

The per-rank arrays passed to ``solve()`` are the extended slabs, cut from global arrays with
``SlabPartition.slab(arr, kind)`` or generated directly per rank (``scenes.viscous_column(..., x0=, gx_total=)``).
"""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _arrays as A
from . import _native as N
from .ViscosityCGSolver3D import _DT, _TORCH, _Engine, _fine_shape, _mac_args


def balanced_starts(cost, world, min_cells=2):
    """Slab boundaries that equalise the summed per-plane cost (prefix-sum splitting, every slab >= min_cells)."""
    cost = np.asarray(cost, dtype=np.float64)
    nx = cost.size
    pre = np.concatenate([[0.0], np.cumsum(cost)])
    starts = [0]
    for r in range(1, world):
        target = pre[-1] * r / world
        b = int(np.searchsorted(pre, target, side="left"))
        if b > 0 and abs(pre[b - 1] - target) < abs(pre[b] - target):
            b -= 1
        b = max(b, starts[-1] + min_cells)
        b = min(b, nx - min_cells * (world - r))
        starts.append(b)
    starts.append(nx)
    return starts


def plane_cost_from_sphi(sphi, gres, fluid_weight=1.0):
    """Relative cost of each x-plane of cells for the viscosity iteration: K2/K3 stream every row alike, K1 skips solid
    rows, so a plane costs 1 + fluid_weight * (fraction of its cell centres inside the fluid region sphi >= 0).
    fluid_weight was fitted on B200 at 256^3 fp64: a fully fluid plane costs ~2x a solid one (K1 ~1 us per 40 %-fluid plane on
    top of ~2.3 us of K2+K3 per plane)."""
    g = tuple(int(n) for n in gres)
    centres = sphi[1::2, 1::2, 1::2][: g[0], : g[1], : g[2]]
    frac = (centres >= 0).to(torch.float64).mean(dim=(1, 2)).cpu().numpy()
    return 1.0 + fluid_weight * frac


def plane_cost_active(sphi, lvol, gres, active_set="nonzero", cg_weight=20.0):
    """Relative cost of each x-plane of cells with the active-set CG kernels: the once-per-solve passes (pack, load,
    extrapolation, RHS) stream every plane alike (cost 1), the CG iterations only touch active rows (cg_weight x the
    plane's active fraction; ~20 for a few hundred iterations on B200).  The active fraction is estimated from the inputs:
    "fluid": cell centres with sphi >= 0; "nonzero": fine-grid nodes of the plane's three fine layers with lvol != 0."""
    g = tuple(int(n) for n in gres)
    if active_set == "fluid":
        centres = sphi[1::2, 1::2, 1::2][: g[0], : g[1], : g[2]]
        frac = (centres >= 0).to(torch.float64).mean(dim=(1, 2))
    else:
        nz = (lvol != 0).to(torch.float64).mean(dim=(1, 2))          # per fine x-layer
        frac = (nz[0:-1:2] + nz[1::2] + nz[2::2]) / 3.0
        frac = torch.clamp(frac * 2.0, max=1.0)                        # one-node dilation of the non-zero set
    return 1.0 + cg_weight * frac.cpu().numpy()


class SlabPartition:
    """x-slabs.  ``[c0, c1)`` = owned cells, ``[e0, e1)`` = extended cells held locally.  With ``plane_cost`` (one
    relative cost per x-plane of cells) the cuts equalise the summed cost instead of the cell count."""

    def __init__(self, gres, world, rank, plane_cost=None, ext=1):
        """ext: cells of overlap held towards each neighbour (1 for the slab solver's halo; 4 = extrapolation sweeps + 1 for
        the gathered solver, whose ranks never exchange halos during set-up)."""
        self.ext = int(ext)
        self.gres = tuple(int(n) for n in gres)
        self.world, self.rank = int(world), int(rank)
        nx = self.gres[0]
        if nx < 2 * self.world:
            raise ValueError(f"grid too thin for {world} slabs: nx={nx}")
        if plane_cost is not None:
            if len(plane_cost) != nx:
                raise ValueError("plane_cost needs one entry per x-plane of cells")
            starts = balanced_starts(plane_cost, self.world)
        else:
            base, rem = divmod(nx, self.world)
            starts = [r * base + min(r, rem) for r in range(self.world + 1)]
        self.starts = starts
        self.c0, self.c1 = starts[self.rank], starts[self.rank + 1]
        self.has_lo = self.rank > 0
        self.has_hi = self.rank < self.world - 1
        self.e0 = max(0, self.c0 - self.ext) if self.has_lo else self.c0
        self.e1 = min(nx, self.c1 + self.ext) if self.has_hi else self.c1
        self.local_gres = (self.e1 - self.e0,) + self.gres[1:]

    def slab(self, arr, kind):
        """Cut this rank's extended slab out of a GLOBAL array.  kind: 'u' (x-faces, nx+1 planes), 'v'/'w'/'cell'
        (nx planes) or 'fine' ((2nx+1) planes, also for arrays with trailing component axes)."""
        if kind == "u":
            return arr[self.e0: self.e1 + 1]
        if kind in ("v", "w", "cell"):
            return arr[self.e0: self.e1]
        if kind == "fine":
            return arr[2 * self.e0: 2 * self.e1 + 1]
        raise ValueError(kind)

    def owned_planes(self, kind):
        """(lo, hi) local plane range of the entries this rank owns inside its extended slab."""
        off = self.c0 - self.e0
        n = self.c1 - self.c0
        if kind == "u":
            return off, off + n + (0 if self.has_hi else 1)
        return off, off + n


_comm_cache = {}


def get_comm(group=None):
    """Create (once per process group) the native NCCL communicator; the unique id travels over torch.distributed."""
    key = ("world",) if group is None else tuple(dist.get_process_group_ranks(group))     # (id(group) can be reused after GC)
    if key in _comm_cache:
        return _comm_cache[key]
    lib = N.load()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        N.check(lib.fs_comm_unique_id(buf), "fs_comm_unique_id")
    box = [bytes(buf.raw)]
    dist.broadcast_object_list(box, src=0, group=group)
    idbuf = ctypes.create_string_buffer(box[0], 128)
    h = ctypes.c_void_p()
    N.check(lib.fs_comm_create(ctypes.byref(h), rank, world, idbuf), "fs_comm_create")
    _comm_cache[key] = h
    return h


class SlabViscosityCGSolver3D:
    """Multi-GPU counterpart of ViscosityCGSolver3D: same ``solve()`` argument list, per-rank extended slabs."""

    def __init__(self, gres, bound_size, dtype=torch.float64, group=None, transport=None, partition=None, active_set="nonzero", cg_mode="auto"):
        """partition: a SlabPartition (e.g. cost-balanced); default = equal cell counts.
        transport: "p2p" (default; collectives fused into the kernels over CUDA-IPC peer memory, one NVSwitch box)
        or "nccl" (one NCCL halo exchange + two NCCL all-reduces per iteration; also the fallback if IPC is unavailable)."""
        transport = transport or os.environ.get("FLUIDSOLVER_B200_TRANSPORT", "p2p")
        if transport not in ("p2p", "nccl"):
            raise ValueError("transport must be 'p2p' or 'nccl'")
        if not dist.is_initialized():
            raise RuntimeError("SlabViscosityCGSolver3D needs an initialised torch.distributed process group")
        self.gres = gres
        self._g = A.to_host_ints(gres)
        self.part = partition if partition is not None else SlabPartition(self._g, dist.get_world_size(group), dist.get_rank(group))
        if self.part.gres != tuple(self._g) or self.part.world != dist.get_world_size(group) or self.part.rank != dist.get_rank(group):
            raise ValueError("partition does not match gres / process group")
        self.cell_size = A.to_host_f64(bound_size, 3) / np.asarray(self._g, dtype=np.float64)
        self.cell_vol = float(np.prod(self.cell_size))
        self._code = _DT[dtype]
        self.transport = transport
        self._e = _Engine(self.part.local_gres, self._code, shared=(transport == "p2p"))
        self._e.set_active_mode(active_set)
        self._e.set_cg_mode(cg_mode)
        self._comm = get_comm(group)
        N.check(self._e.lib.fs_visc3d_set_slab(self._e.h, self._comm, int(self.part.has_lo), int(self.part.has_hi)), "fs_visc3d_set_slab")
        self._mapped = []
        if transport == "p2p" and self.part.world > 1:
            self._connect_peers(group)
        for vec, nm in ((N.VEC_D, "d"), (N.VEC_R, "r"), (N.VEC_Q, "q"), (N.VEC_X, "x"), (N.VEC_B, "b")):
            for c, ax in enumerate("xyz"):
                setattr(self, f"{nm}_{ax}", self._e.view(vec, c))
        self.alpha = self.beta = self.delta = 0.0
        self.iterations = 0
        self.max_iter = int(np.prod(np.asarray(self._g, dtype=np.int64)))
        self._closed = False

    def _connect_peers(self, group):
        """Exchange CUDA-IPC handles of the slab workspaces and scalar mailboxes and map the peers' buffers."""
        lib, e, part = self._e.lib, self._e, self.part
        rank, world = part.rank, part.world
        self._mbox = lib.fs_shared_alloc(1024)
        if not self._mbox:
            raise N.NativeError("fs_shared_alloc(mailbox) failed")
        hw, hm = ctypes.create_string_buffer(64), ctypes.create_string_buffer(64)
        N.check(lib.fs_shared_get_handle(e.shared_ptr, hw), "fs_shared_get_handle")
        N.check(lib.fs_shared_get_handle(self._mbox, hm), "fs_shared_get_handle")
        infos = [None] * world
        dist.all_gather_object(infos, (bytes(hw.raw), bytes(hm.raw), int(part.local_gres[0])), group=group)

        def _open(raw):
            p = lib.fs_shared_open(ctypes.create_string_buffer(raw, 64))
            if not p:
                raise N.NativeError("fs_shared_open failed: " + (lib.fs_last_error() or b"?").decode())
            self._mapped.append(p)
            return p

        lo_ws = _open(infos[rank - 1][0]) if part.has_lo else None
        hi_ws = _open(infos[rank + 1][0]) if part.has_hi else None
        lo_nx = infos[rank - 1][2] if part.has_lo else 0
        hi_nx = infos[rank + 1][2] if part.has_hi else 0
        boxes = (ctypes.c_void_p * world)()
        for r in range(world):
            boxes[r] = self._mbox if r == rank else _open(infos[r][1])
        N.check(lib.fs_visc3d_set_peers(e.h, lo_ws, lo_nx, hi_ws, hi_nx, boxes), "fs_visc3d_set_peers")
        dist.barrier(group=group)

    def close(self, group=None):
        """Collective teardown (MANDATORY for the p2p transport): unmap the peers' buffers before any rank frees its own, then
        release the workspace.  The ``x_x … b_z`` views are dropped: they alias the workspace."""
        if getattr(self, "_closed", True):
            return
        lib = self._e.lib
        torch.cuda.synchronize()
        if dist.is_initialized() and self.part.world > 1:
            dist.barrier(group=group)
        for p in self._mapped:
            lib.fs_shared_close(p)
        self._mapped = []
        if dist.is_initialized() and self.part.world > 1:
            dist.barrier(group=group)
        if getattr(self, "_mbox", None):
            lib.fs_shared_free(self._mbox)
            self._mbox = None
        for nm in "drqxb":
            for ax in "xyz":
                setattr(self, f"{nm}_{ax}", None)
        self._e.close()
        self._closed = True

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        # No collective is possible here (garbage collection order differs between ranks), so memory that peers may still
        # have mapped — the workspace and the mailbox — is deliberately LEAKED rather than freed under a neighbour that
        # could still be storing halo rows or mailbox words into it; only the local mappings of the peers are dropped.
        try:
            if not getattr(self, "_closed", True):
                import warnings
                warnings.warn("SlabViscosityCGSolver3D was not close()d: its IPC-exported workspace is leaked to stay safe", ResourceWarning)
                lib = self._e.lib
                for p in getattr(self, "_mapped", []):
                    lib.fs_shared_close(p)
                self._mapped = []
                self._e.leak()
        except Exception:
            pass

    def solve(self, dt, mu, rho, vx, vy, vz, sphi, sv, lphi, lvol, tol=1e-3):
        g, e = self.part.local_gres, self._e
        v = _mac_args(g, (vx, vy, vz), ("vx", "vy", "vz"))
        s = A.as_arg(sphi, "sphi", shape=_fine_shape(g), want=torch.float64)
        vl = A.as_arg(lvol, "lvol", shape=_fine_shape(g), want=torch.float64)
        st = N.CgStats()
        status = N.check(
            e.lib.fs_visc3d_solve(e.h, float(dt), float(mu), float(rho), self.cell_vol, v[0].ptr, v[1].ptr, v[2].ptr, v[0].code,
                                  s.ptr, vl.ptr, float(tol), int(self.max_iter), ctypes.byref(st), A.stream_ptr()),
            "fs_visc3d_solve")
        self.delta, self.alpha, self.beta, self.iterations = st.delta, st.alpha, st.beta, int(st.iterations)
        if self.transport == "p2p" and e.lib.fs_visc3d_peer_error(e.h):
            raise RuntimeError("a peer GPU did not answer inside a fused all-reduce")
        if status == N.FS_NOT_CONVERGED:
            raise ValueError("Failed to converge!")
        for a in v:
            a.sync_back()


class GatheredViscosityCGSolver3D:
    """Multi-GPU solve for active sets that fit one GPU's L2 (the benchmark scene's 0.45 % of the rows): the once-per-solve
    passes are sharded over the ranks, the CG itself is NOT split.

    Every rank holds a lattice for the GLOBAL grid but packs / loads / extrapolates only the x-window its input arrays
    cover (owned cells extended by ``EXT`` = 3 sweeps + 1 cells, so no halo exchange is needed during set-up), publishes
    the lattice segments the CG will touch on its own planes, all-gathers those records over NCCL and runs the complete CG
    locally with the single-GPU kernels: no NVLink round trip inside an iteration, identical iterates on every rank, and
    each rank writes the rows of its own planes back into its (windowed) velocity arrays.

    ``solve()`` takes the reference's argument list with per-rank window arrays cut by ``SlabPartition(..., ext=4).slab``."""

    EXT = 4

    def __init__(self, gres, bound_size, dtype=torch.float64, group=None, partition=None, active_set="nonzero", cg_mode="auto"):
        """Without an initialised process group a ``partition`` must be given: the object then only offers the three steps
        (``export_step`` / ``import_step`` / ``finish_step``) so that several ranks can be emulated in one process (tests)."""
        self.gres = gres
        self._g = A.to_host_ints(gres)
        self.group = group
        self._dist = dist.is_initialized()
        if not self._dist and partition is None:
            raise RuntimeError("GatheredViscosityCGSolver3D needs an initialised torch.distributed process group (or an explicit partition)")
        if partition is None:
            partition = SlabPartition(self._g, dist.get_world_size(group), dist.get_rank(group), ext=self.EXT)
        self.part = partition
        if self.part.gres != tuple(self._g) or self.part.ext < self.EXT:
            raise ValueError("partition does not match gres, or extends fewer than 4 cells")
        if self._dist and (self.part.world != dist.get_world_size(group) or self.part.rank != dist.get_rank(group)):
            raise ValueError("partition does not match the process group")
        world = self.part.world
        self.cell_size = A.to_host_f64(bound_size, 3) / np.asarray(self._g, dtype=np.float64)
        self.cell_vol = float(np.prod(self.cell_size))
        self._code = _DT[dtype]
        self._e = _Engine(self._g, self._code)                   # lattice of the WHOLE grid on every rank
        self._e.set_active_mode(active_set)
        self._e.set_cg_mode(cg_mode)
        lib = self._e.lib
        N.check(lib.fs_visc3d_set_window(self._e.h, self.part.e0, self.part.e1), "fs_visc3d_set_window")
        self._rec = int(lib.fs_visc3d_gather_record_bytes(self._e.h))
        self._send = torch.empty(0, dtype=torch.uint8, device=A.device())
        self._recv = torch.empty(0, dtype=torch.uint8, device=A.device())
        self._cap = 0
        self._count = torch.zeros(1, dtype=torch.int64, device=A.device())
        self._counts = torch.zeros(world, dtype=torch.int64, device=A.device())
        # planes this rank writes back: its cells' low-side faces, plus the closing plane on the last rank
        self.own_lo = self.part.c0
        self.own_hi = self.part.c1 + (0 if self.part.has_hi else 1)
        self.alpha = self.beta = self.delta = 0.0
        self.iterations = 0
        self.published = 0                                       # segments in the last solve's exchange (all ranks)
        self.max_iter = int(np.prod(np.asarray(self._g, dtype=np.int64)))
        for vec, nm in ((N.VEC_D, "d"), (N.VEC_R, "r"), (N.VEC_Q, "q"), (N.VEC_X, "x"), (N.VEC_B, "b")):
            for c, ax in enumerate("xyz"):                       # GLOBAL-grid views; valid on the active segments (x: also on this rank's window)
                setattr(self, f"{nm}_{ax}", self._e.view(vec, c))

    def _grow(self, n):
        if n > self._cap:
            old = self._send
            self._cap = max(int(n * 1.25) + 64, 4096)
            self._send = torch.empty(self._cap * self._rec, dtype=torch.uint8, device=A.device())
            self._send[: old.numel()].copy_(old)

    def active_info(self):
        return self._e.active_info()

    def close(self, group=None):
        self._e = None

    # ---- the three steps of a solve ---------------------------------------------------------------------------
    def export_step(self, vx, vy, vz, sphi, lvol):
        """pack + load + extrapolate this rank's window and write the records of the segments it publishes.
        Returns (record count, uint8 tensor holding them)."""
        e, part, lib = self._e, self.part, self._e.lib
        g_win = (part.e1 - part.e0,) + tuple(self._g[1:])
        v = _mac_args(g_win, (vx, vy, vz), ("vx", "vy", "vz"))
        if len({a.code for a in v}) != 1:
            raise TypeError("vx, vy, vz must share one dtype")
        s = A.as_arg(sphi, "sphi", shape=_fine_shape(g_win), want=torch.float64)
        vl = A.as_arg(lvol, "lvol", shape=_fine_shape(g_win), want=torch.float64)
        stream = A.stream_ptr()
        self._grow(4096)
        cnt = ctypes.c_int64()
        N.check(lib.fs_visc3d_gather_export(e.h, v[0].ptr, v[1].ptr, v[2].ptr, v[0].code, s.ptr, vl.ptr, self.cell_vol * 0.125,
                                            self.own_lo, self.own_hi, self._send.data_ptr(), self._cap, ctypes.byref(cnt), stream),
                "fs_visc3d_gather_export")
        if cnt.value > self._cap:                                # first solve, or the active set grew: enlarge and write again
            self._grow(cnt.value)
            N.check(lib.fs_visc3d_gather_reexport(e.h, self._send.data_ptr(), self._cap, stream), "fs_visc3d_gather_reexport")
        self._v = v
        return cnt.value, self._send

    def import_step(self, records, counts_dev, stride):
        """scatter every rank's records (rank r at r*stride records; counts_dev: int64 device tensor) into the lattice"""
        e = self._e
        N.check(e.lib.fs_visc3d_gather_import(e.h, records.data_ptr(), counts_dev.data_ptr(), self.part.world, int(stride), self.part.rank,
                                              A.stream_ptr()), "fs_visc3d_gather_import")

    def finish_step(self, dt, mu, rho, tol):
        """RHS + CG on the complete active set, write-back of this rank's rows into the arrays given to export_step"""
        e, v = self._e, self._v
        st = N.CgStats()
        status = N.check(
            e.lib.fs_visc3d_solve_packed(e.h, float(dt), float(mu), float(rho), self.cell_vol, v[0].ptr, v[1].ptr, v[2].ptr, v[0].code,
                                         self.own_lo, self.own_hi, float(tol), int(self.max_iter), ctypes.byref(st), A.stream_ptr()),
            "fs_visc3d_solve_packed")
        self.delta, self.alpha, self.beta, self.iterations = st.delta, st.alpha, st.beta, int(st.iterations)
        if status == N.FS_NOT_CONVERGED:
            raise ValueError("Failed to converge!")
        for a in v:
            a.sync_back()
        self._v = None

    def solve(self, dt, mu, rho, vx, vy, vz, sphi, sv, lphi, lvol, tol=1e-3):
        if not self._dist:
            raise RuntimeError("solve() needs torch.distributed; use export_step / import_step / finish_step to emulate ranks")
        world = self.part.world
        timing = os.environ.get("FLUIDSOLVER_B200_TIMING", "0") != "0"
        if timing:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            ev[0].record()
        count, _ = self.export_step(vx, vy, vz, sphi, lvol)
        if timing:
            ev[1].record()
        self._count.fill_(count)
        if world > 1:
            dist.all_gather_into_tensor(self._counts, self._count, group=self.group)
            counts = self._counts.tolist()
        else:
            self._counts.copy_(self._count)
            counts = [count]
        stride = max(counts)
        self.published = int(sum(counts))
        self._grow(stride)                                       # the all-gather reads `stride` records from every rank's send buffer
        need = world * stride * self._rec
        if self._recv.numel() < max(need, 1):
            self._recv = torch.empty(int(need * 1.25) + 64, dtype=torch.uint8, device=A.device())
        if world > 1 and stride > 0:
            if os.environ.get("FLUIDSOLVER_B200_GATHER_EXACT", "0") != "0":
                # option: every rank contributes exactly its own records (a list all-gather with uneven sizes = one coalesced group
                # of NCCL broadcasts) instead of padding every contribution to the largest one (93 MB instead of 31 MB on the benchmark
                # scene at 8 ranks, 0.39 ms).  Measured at 2 ranks the group of broadcasts is SLOWER than the single padded all-gather
                # (step 2.77 vs 2.53 ms), so the padded form stays the default.
                rec = self._rec
                outs = [self._recv[r * stride * rec: r * stride * rec + max(c, 1) * rec] for r, c in enumerate(counts)]
                dist.all_gather(outs, self._send[: max(count, 1) * rec], group=self.group)
            else:
                dist.all_gather_into_tensor(self._recv[:need], self._send[: stride * self._rec], group=self.group)
            if timing:
                ev[2].record()
            self.import_step(self._recv, self._counts, stride)
        else:
            if timing:
                ev[2].record()
            self.import_step(self._send, self._counts, stride)  # one rank: its own block is skipped, only the list is rebuilt
        if timing:
            ev[3].record()
        try:
            self.finish_step(dt, mu, rho, tol)
        finally:
            if timing:
                ev[4].record()
                torch.cuda.synchronize()
                self.timing = {"export_ms": ev[0].elapsed_time(ev[1]), "exchange_ms": ev[1].elapsed_time(ev[2]), "import_ms": ev[2].elapsed_time(ev[3]),
                               "cg_and_store_ms": ev[3].elapsed_time(ev[4]), "records_sent": int(count), "stride": int(stride),
                               "exchange_MB": need / 1e6}


def emulate_gathered_solve(solvers, scenes_per_rank, dt, mu, rho, tol=1e-3):
    """Run the gathered solve of ``len(solvers)`` ranks inside ONE process on one GPU (no process group): the exchange
    becomes a concatenation.  Test helper for boxes with fewer GPUs than ranks; the steps are the ones solve() runs."""
    world = len(solvers)
    dev = A.device()
    counts, bufs = [], []
    for s, sc in zip(solvers, scenes_per_rank):
        n, buf = s.export_step(sc["vx"], sc["vy"], sc["vz"], sc["sphi"], sc["lvol"])
        counts.append(n)
        bufs.append(buf)
    stride = max(counts)
    rec = solvers[0]._rec
    allrec = torch.zeros(max(world * stride * rec, 1), dtype=torch.uint8, device=dev)
    for r, (n, buf) in enumerate(zip(counts, bufs)):
        allrec[r * stride * rec: r * stride * rec + n * rec].copy_(buf[: n * rec])
    cdev = torch.tensor(counts, dtype=torch.int64, device=dev)
    for s in solvers:
        s.published = int(sum(counts))
        s.import_step(allrec, cdev, stride)
    err = None
    for s in solvers:
        try:
            s.finish_step(dt, mu, rho, tol)
        except ValueError as e:          # every emulated rank takes the same decision; finish them all before re-raising
            err = e
    if err is not None:
        raise err


def scatter_scene(sc, part):
    """Per-rank extended slabs of a GLOBAL scene dict (scenes.buckling etc.)."""
    out = dict(sc)
    for k, kind in (("vx", "u"), ("vy", "v"), ("vz", "w"), ("sphi", "fine"), ("lvol", "fine"), ("lphi", "cell"), ("sv", "fine")):
        if k in sc and sc[k] is not None:
            out[k] = part.slab(sc[k], kind).contiguous()
    return out


# ------------------------------------------------------------------------------------------------------------
# bench.py --gpus N>1
# ------------------------------------------------------------------------------------------------------------

def bench_distributed(args, metric, unit, config, peak, peak_src):
    """bench.py --gpus N>1 (one rank per GPU).  Default workload (active_set "nonzero": an L2-sized CG problem): the gathered
    solve — set-up and host transfers sharded over the ranks, the CG replicated, no NVLink traffic inside the iteration.
    The HBM-bound variant of the same scene (active_set "fluid": every row the reference computes) runs on x-slabs with the
    collectives fused into the kernels and is reported beside it (`hbm_variant`), like in the N=1 line.  Both are checked
    against the single-GPU solver on the same scene inside the run (`parity`)."""
    import faulthandler
    import scenes
    from bench import ClockSampler, counts, kernel_study
    from .ViscosityCGSolver3D import ViscosityCGSolver3D

    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # watchdog: a multi-GPU run that stalls (a peer died, a collective mismatched) dumps every thread's stack and exits
    # instead of hanging until somebody's time limit
    faulthandler.dump_traceback_later(float(os.environ.get("FLUIDSOLVER_B200_WATCHDOG_S", "420")), exit=True, file=sys.stderr)
    t_start = time.time()
    trace_on = os.environ.get("FLUIDSOLVER_B200_TRACE", "0") != "0"

    def trace(msg):
        if trace_on:
            print(f"[bench r{rank} +{time.time() - t_start:6.1f}s] {msg}", file=sys.stderr, flush=True)

    tdtype = torch.float64 if args.dtype == "f64" else torch.float32
    esz = 8 if args.dtype == "f64" else 4
    n = args.size
    g = (n, n, n)
    full = scenes.buckling(n, device="cuda", mu=args.mu)
    bound = full["bound_size"]
    aset = getattr(args, "active_set", "nonzero")
    mode = os.environ.get("FLUIDSOLVER_B200_MULTI", "auto")
    if mode == "auto":
        mode = "gathered" if aset == "nonzero" else "slab"
    balance = os.environ.get("FLUIDSOLVER_B200_BALANCE", "1") != "0"

    def make(which, active_set):
        if which == "gathered":
            part = SlabPartition(g, world, rank, ext=GatheredViscosityCGSolver3D.EXT)
            sol = GatheredViscosityCGSolver3D(g, bound, dtype=tdtype, partition=part, active_set=active_set, cg_mode=getattr(args, "cg_mode", "auto"))
        else:
            # cost of an x-plane = its share of the once-per-solve passes (1) + the CG work of its active rows: a fully active plane
            # costs ~4.2 us per iteration against ~3.5 us of set-up per plane (measured on B200), i.e. weight ~ 1.2 x iterations
            part = SlabPartition(g, world, rank, plane_cost=plane_cost_active(full["sphi"], full["lvol"], g, active_set, cg_weight=1.2 * args.iters)
                                 if balance else None)
            sol = SlabViscosityCGSolver3D(g, bound, dtype=tdtype, partition=part, active_set=active_set, cg_mode=getattr(args, "cg_mode", "auto"))
        return sol, part, scatter_scene(full, part)

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)          # device time, max over ranks
        return float(t.item())

    def window_step(sol, sc, arrays):
        def step():
            try:
                sol.solve(sc["dt"], args.mu, sc["rho"], arrays["vx"], arrays["vy"], arrays["vz"], arrays["sphi"], None, None, arrays["lvol"], tol=0.0)
            except ValueError:
                pass
            assert sol.iterations == args.iters, sol.iterations
        return step

    def parity(sol, part, sc, active_set):
        """a converging solve (tol 1e-3) on the multi-GPU path and on the single-GPU solver (whole grid, this rank's GPU);
        every rank compares the planes it owns, the worst rank is reported"""
        keep = sol.max_iter
        cap_it = 20000                            # (both sides bounded: a solve that stalls must fail the check, not hang the run)
        sol.max_iter = cap_it
        v = [sc[k].clone() for k in ("vx", "vy", "vz")]
        failed = False
        try:
            sol.solve(sc["dt"], args.mu, sc["rho"], *v, sc["sphi"], None, None, sc["lvol"], tol=1e-3)
        except ValueError:
            failed = True
        ref = ViscosityCGSolver3D(g, bound, dtype=tdtype, active_set=active_set)
        ref.max_iter = cap_it
        rv = [full[k].clone() for k in ("vx", "vy", "vz")]
        try:
            ref.solve(full["dt"], args.mu, full["rho"], *rv, full["sphi"], None, None, full["lvol"], tol=1e-3)
        except ValueError:
            failed = True
        worst = 0.0
        for a, b, kind in zip(v, rv, ("u", "v", "w")):
            lo, hi = part.owned_planes(kind)
            mine, theirs = a[lo:hi].double(), b[part.e0 + lo: part.e0 + hi].double()
            worst = max(worst, float((mine - theirs).norm() / theirs.norm().clamp_min(1e-300)))
        t = torch.tensor([worst, float(sol.iterations), -float(sol.iterations)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out = {"tol": 1e-3, "iters_multi_gpu": int(sol.iterations), "iters_single_gpu": int(ref.iterations),
               "iters_equal": int(t[1].item()) == int(-t[2].item()) == int(ref.iterations),
               "iters_within_2pct": abs(sol.iterations - ref.iterations) <= max(1, round(0.02 * ref.iterations)),
               "delta_rel_diff": abs(sol.delta - ref.delta) / max(ref.delta, 1e-300), "owned_rel_l2": float(t[0].item())}
        out["converged"] = not failed
        out["ok"] = bool(not failed and out["iters_within_2pct"] and int(t[1].item()) == int(-t[2].item()) and out["owned_rel_l2"] < 1e-4)
        sol.max_iter = keep
        del ref
        torch.cuda.empty_cache()
        return out

    trace(f"scene ready, mode={mode}")
    hbm_variant = None
    want_variant = bool(getattr(args, "hbm_leg", 1)) and aset == "nonzero"
    slab_first = os.environ.get("FLUIDSOLVER_B200_SLAB_FIRST", "0") != "0"
    # ---- HBM-bound variant of the same scene on slabs (every fluid row) -----------------------------------------------
    def run_slab_variant():
        no_clocks = os.environ.get("FLUIDSOLVER_B200_NOCLOCKS", "0") != "0"
        s2, p2, sc2 = make("slab", "fluid")
        trace(f"slab solver built, starts={p2.starts}")
        s2.max_iter = args.iters
        st2 = window_step(s2, sc2, sc2)
        for _ in range(2):
            st2()
        trace("slab warm-up done")
        nst = max(2, min(args.steps, 5))
        if no_clocks:
            ms2 = timed(st2, nst) / nst
            clk = None
        else:
            with ClockSampler(local) as cl2:
                ms2 = timed(st2, nst) / nst
            clk = cl2.summary()
        trace(f"slab timed: {ms2:.3f} ms/step")
        par2 = parity(s2, p2, sc2, "fluid")
        trace(f"slab parity: {par2}")
        segs2 = s2._e.active_info()
        out = {"what": "same scene and window, active_set='fluid' (every row the reference's kernels compute), x-slabs with fused collectives",
               "value": args.iters / (ms2 * 1e-3), "unit": unit, "ms_per_step": ms2, "transport": s2.transport, "slab_starts": p2.starts,
               "segments_rank0": segs2[0], "parity": par2, "clocks": clk}
        s2.close()
        del s2, sc2
        torch.cuda.empty_cache()
        return out

    if want_variant and slab_first:
        hbm_variant = run_slab_variant()

    # ---- the default workload ----------------------------------------------------------------------------------
    solver, part, sc = make(mode, aset)
    trace("solver built")
    config = dict(config, multi_gpu=(f"gathered: set-up sharded over {world} ranks (x-windows, 4-cell overlap), CG replicated on every rank "
                                     "(no inter-GPU traffic inside the iteration); records all-gathered over NCCL" if mode == "gathered" else
                                     f"x-slabs over {world} ranks, transport {solver.transport}: halo rows and the CG reduction fused into the kernels over peer memory"),
                  slab_starts=part.starts)
    solver.max_iter = args.iters
    step = window_step(solver, sc, sc)
    for _ in range(max(args.warmup, 3)):
        step()
    trace("warm-up done")
    l0 = N.launch_count()
    with ClockSampler(local) as clocks:
        ms = timed(step, args.steps)
    trace(f"timed: {ms / args.steps:.3f} ms/step; timing of the last solve: {getattr(solver, 'timing', None)}")
    launches = N.launch_count() - l0
    value = args.iters * args.steps / (ms * 1e-3)

    host = {k: sc[k].cpu().pin_memory() for k in ("vx", "vy", "vz", "sphi", "lvol")}
    xs = (solver.x_x, solver.x_y, solver.x_z)
    if mode == "gathered":                      # global-grid views: this rank's share of the result = its owned planes
        xs = [x[part.e0 + part.owned_planes(k)[0]: part.e0 + part.owned_planes(k)[1]] for x, k in zip(xs, ("u", "v", "w"))]
    out_host = [torch.empty(tuple(a.shape), dtype=host["vx"].dtype).pin_memory() for a in xs]
    h2d = sum(host[k].numel() * host[k].element_size() for k in host)
    d2h = sum(a.numel() * a.element_size() for a in out_host)
    step_h = window_step(solver, sc, host)

    def step_e2e():
        step_h()
        for o, x in zip(out_host, xs):
            o.copy_(x, non_blocking=True)         # the step's result in the caller's precision (fp32 velocities)
        torch.cuda.synchronize()

    e2e_steps = max(2, min(args.steps, 5))
    step_e2e()
    with ClockSampler(local) as clocks_e2e:
        e2e_ms = timed(step_e2e, e2e_steps)
    e2e_value = args.iters * e2e_steps / (e2e_ms * 1e-3)
    trace(f"e2e: {e2e_ms / e2e_steps:.3f} ms/step")
    del host
    tot = torch.tensor([float(h2d), float(d2h), float(launches)], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot)
    par = parity(solver, part, sc, aset)
    trace(f"parity: {par}")

    # per-iteration / per-kernel figures of rank 0's engine (gathered: the complete CG, identical on every rank)
    scale = sc["dt"] / solver.cell_vol / sc["rho"]
    kin = None
    if mode == "gathered":
        solver.max_iter = 0
        try:
            solver.solve(sc["dt"], args.mu, sc["rho"], sc["vx"], sc["vy"], sc["vz"], sc["sphi"], None, None, sc["lvol"], tol=0.0)
        except ValueError:
            pass
        solver.max_iter = args.iters
        if rank == 0:
            kin = kernel_study(solver, solver._e.lib, N, torch, scale, args, esz)
    trace("kernel study done")
    dist.barrier()
    published = getattr(solver, "published", None)
    solver.close()
    del solver, sc
    torch.cuda.empty_cache()
    trace("default leg closed")

    if want_variant and not slab_first:
        hbm_variant = run_slab_variant()

    if rank == 0:
        F, V7 = counts(n)
        words_iter = 11 * F + V7
        iter_gbs = words_iter * esz * value / 1e9
        roofline = {"bound": "l2-latency" if (kin and kin["working_set"] < 100e6) else "hbm", "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
                    "peak_source": peak_src,
                    "note": "the default scene's CG is an L2-resident problem (see the N=1 line); no HBM fraction is claimed for it. "
                            "hbm_variant is the HBM-bound form of the same scene"}
        if kin:
            roofline.update({"kernel": kin["persistent_name"] if kin["persistent"] else max(kin["kernel_ms"], key=kin["kernel_ms"].get),
                             "achieved": kin["iter_bytes"] / (kin["iter_ms"] * 1e-3) / 1e9, "cg_iteration_us": kin["iter_ms"] * 1e3,
                             "cg_mode": kin["mode"], "active_set": kin["active"], "per_kernel_ms_standalone": kin["kernel_ms"],
                             "setup_ms_per_step": ms / args.steps - kin["iter_ms"] * args.iters})
        roofline["dense_equivalent"] = {"algorithmic_GB_per_iter": words_iter * esz / 1e9, "equivalent_GBps_all_gpus": iter_gbs}
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(tot[0].item()), "d2h_bytes_per_step": int(tot[1].item()),
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps, "clocks": clocks_e2e.summary()},
            "gpu_launches": int(tot[2].item()), "parity": par, "published_segments": published,
            "roofline": roofline, "hbm_variant": hbm_variant, "clocks": clocks.summary(),
        }
        print(json.dumps(line), file=getattr(args, "_json_out", None) or sys.stdout, flush=True)
    faulthandler.cancel_dump_traceback_later()
    dist.barrier()
    ok = par["ok"] and (hbm_variant is None or hbm_variant["parity"]["ok"])
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1
