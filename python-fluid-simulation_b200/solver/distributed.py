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

    def __init__(self, gres, world, rank, plane_cost=None):
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
        self.e0 = self.c0 - (1 if self.has_lo else 0)
        self.e1 = self.c1 + (1 if self.has_hi else 0)
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
    key = id(group)
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
        """Collective teardown: unmap the peers' buffers before any rank frees its own."""
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

    def __del__(self):
        try:                       # best effort (no collective here: GC order differs between ranks)
            lib = self._e.lib
            for p in getattr(self, "_mapped", []):
                lib.fs_shared_close(p)
            self._mapped = []
            if getattr(self, "_mbox", None):
                lib.fs_shared_free(self._mbox)
                self._mbox = None
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
    import scenes
    from bench import ClockSampler, counts

    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    tdtype = torch.float64 if args.dtype == "f64" else torch.float32
    esz = 8 if args.dtype == "f64" else 4
    n = args.size
    g = (n, n, n)
    full = scenes.buckling(n, device="cuda", mu=args.mu)
    balance = os.environ.get("FLUIDSOLVER_B200_BALANCE", "1") != "0"
    aset = getattr(args, "active_set", "nonzero")
    part = SlabPartition(g, world, rank, plane_cost=plane_cost_active(full["sphi"], full["lvol"], g, aset) if balance else None)
    sc = scatter_scene(full, part)
    bound = full["bound_size"]
    del full
    torch.cuda.empty_cache()
    solver = SlabViscosityCGSolver3D(g, bound, dtype=tdtype, partition=part, active_set=getattr(args, "active_set", "nonzero"),
                                     cg_mode=getattr(args, "cg_mode", "auto"))
    config = dict(config, transport=solver.transport, slab_starts=part.starts, balanced=balance)
    solver.max_iter = args.iters
    dev_in = [sc[k] for k in ("vx", "vy", "vz")]

    def step_device():
        try:
            solver.solve(sc["dt"], args.mu, sc["rho"], *dev_in, sc["sphi"], None, None, sc["lvol"], tol=0.0)
        except ValueError:
            pass
        assert solver.iterations == args.iters, solver.iterations

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

    for _ in range(max(args.warmup, 3)):
        step_device()
    l0 = N.launch_count()
    with ClockSampler(local) as clocks:
        ms = timed(step_device, args.steps)
    launches = N.launch_count() - l0
    value = args.iters * args.steps / (ms * 1e-3)

    host = {k: sc[k].cpu().pin_memory() for k in ("vx", "vy", "vz", "sphi", "lvol")}
    out_host = [torch.empty(tuple(a.shape), dtype=tdtype).pin_memory() for a in (solver.x_x, solver.x_y, solver.x_z)]
    h2d = sum(host[k].numel() * host[k].element_size() for k in host)
    d2h = sum(a.numel() * a.element_size() for a in out_host)

    def step_e2e():
        try:
            solver.solve(sc["dt"], args.mu, sc["rho"], host["vx"], host["vy"], host["vz"], host["sphi"], None, None, host["lvol"], tol=0.0)
        except ValueError:
            pass
        for o, x in zip(out_host, (solver.x_x, solver.x_y, solver.x_z)):
            o.copy_(x, non_blocking=True)
        torch.cuda.synchronize()

    e2e_steps = max(2, min(args.steps, 5))
    step_e2e()
    e2e_ms = timed(step_e2e, e2e_steps)
    e2e_value = args.iters * e2e_steps / (e2e_ms * 1e-3)
    tot = torch.tensor([float(h2d), float(d2h), float(launches)], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot)

    # per-kernel timing on this rank's slab (no communication inside these launches)
    F, V7 = counts(n)
    lib = solver._e.lib
    scale = sc["dt"] / solver.cell_vol / sc["rho"]
    stream = torch.cuda.current_stream().cuda_stream
    segs, segs_total, rows = solver._e.active_info()
    pts = segs * 32
    kbytes = {"K1 visc3d_apply_dot": pts * (13 * esz + 1), "K2 cg_update_xr": pts * 18 * esz, "K3 cg_update_d": pts * 9 * esz}
    kern = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for which, name in ((1, "K1 visc3d_apply_dot"), (2, "K2 cg_update_xr"), (3, "K3 cg_update_d")):
        N.check(lib.fs_visc3d_kernel_enqueue(solver._e.h, which, scale, args.mu, 3, stream), "warm")
        torch.cuda.synchronize()
        ev0.record()
        N.check(lib.fs_visc3d_kernel_enqueue(solver._e.h, which, scale, args.mu, 30, stream), "time")
        ev1.record()
        torch.cuda.synchronize()
        kern[name] = ev0.elapsed_time(ev1) / 30
    dist.barrier()
    if rank == 0:
        dom = max(kern, key=kern.get)
        achieved = kbytes[dom] / (kern[dom] * 1e-3) / 1e9
        words_iter = 11 * F + V7
        iter_gbs = words_iter * esz * value / 1e9
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(tot[0].item()), "d2h_bytes_per_step": int(tot[1].item()),
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps},
            "gpu_launches": int(tot[2].item()),
            "roofline": {"bound": "hbm", "kernel": dom + " (rank 0 slab)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "per_kernel_ms": kern,
                         "active_set": {"mode": getattr(args, "active_set", "nonzero"), "segments_rank0": segs, "segments_total_rank0": segs_total,
                                        "computed_rows_rank0": rows},
                         "dense_equivalent": {"algorithmic_GB_per_iter": words_iter * esz / 1e9, "achieved_GBps_all_gpus": iter_gbs,
                                       "frac_of_aggregate_peak": iter_gbs / (peak * world)}},
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), file=getattr(args, "_json_out", None) or sys.stdout, flush=True)
    solver.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0
