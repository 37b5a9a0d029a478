"""CPU: host-side logic of the slab partition (world_size-2 gloo processes + pure partition arithmetic)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_partition_covers_grid_exactly():
    from solver.distributed import SlabPartition
    for nx in (8, 37, 256):
        for world in (1, 2, 3, 4, 8):
            if nx < 2 * world:
                continue
            parts = [SlabPartition((nx, 5, 6), world, r) for r in range(world)]
            assert parts[0].c0 == 0 and parts[-1].c1 == nx
            for a, b in zip(parts, parts[1:]):
                assert a.c1 == b.c0 and a.has_hi and b.has_lo
                assert a.e1 == a.c1 + 1 and b.e0 == b.c0 - 1
            assert not parts[0].has_lo and not parts[-1].has_hi
            sizes = [p.c1 - p.c0 for p in parts]
            assert max(sizes) - min(sizes) <= 1
            # owned u planes tile [0, nx] exactly once, v planes [0, nx)
            for kind, total in (("u", nx + 1), ("v", nx)):
                covered = []
                for p in parts:
                    lo, hi = p.owned_planes(kind)
                    covered += list(range(p.e0 + lo, p.e0 + hi))
                assert covered == list(range(total))


def test_cost_balanced_partition():
    from solver.distributed import SlabPartition, balanced_starts
    cost = np.ones(64)
    cost[16:48] = 1.5                      # fluid planes cost more
    starts = balanced_starts(cost, 8)
    assert starts[0] == 0 and starts[-1] == 64 and all(b - a >= 2 for a, b in zip(starts, starts[1:]))
    sums = [cost[a:b].sum() for a, b in zip(starts, starts[1:])]
    assert max(sums) <= 1.15 * (cost.sum() / 8)
    eq = [cost[a:b].sum() for a, b in zip(range(0, 64, 8), range(8, 72, 8))]
    assert max(sums) < max(eq)
    parts = [SlabPartition((64, 4, 4), 8, r, plane_cost=cost) for r in range(8)]
    assert [p.c0 for p in parts] + [64] == starts
    # degenerate: all the cost in one plane still leaves every slab >= 2 cells
    spike = np.full(16, 1e-6)
    spike[7] = 1.0
    st = balanced_starts(spike, 4)
    assert all(b - a >= 2 for a, b in zip(st, st[1:])) and st[-1] == 16


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import scenes
    from solver.distributed import SlabPartition, scatter_scene
    g = (10, 6, 7)
    full = scenes.viscous_column(g, seed=11)
    part = SlabPartition(g, world, rank)
    sc = scatter_scene(full, part)
    # slab generated directly per rank must equal the slab cut from the global scene (partition-independent noise)
    direct = scenes.viscous_column(part.local_gres, seed=11, x0=part.e0, gx_total=g[0])
    ok = all(torch.equal(sc[k], direct[k]) for k in ("vx", "vy", "vz", "sphi", "lvol", "lphi"))
    # halo consistency: my high halo planes equal my neighbour's first owned planes
    if part.has_hi:
        dist.send(sc["vx"][-2].contiguous(), rank + 1)      # u(c1): owned by rank+1, my halo copy
    if part.has_lo:
        buf = torch.empty_like(sc["vx"][1])
        dist.recv(buf, rank - 1)
        ok = ok and torch.equal(buf, sc["vx"][1])          # my first owned u plane
    # reassemble owned planes on every rank
    lo, hi = part.owned_planes("u")
    own = sc["vx"][lo:hi].contiguous()
    gathered = [None] * world
    dist.all_gather_object(gathered, own.numpy())
    ok = ok and np.array_equal(np.concatenate(gathered, axis=0), full["vx"].numpy())
    res = [None] * world
    dist.all_gather_object(res, bool(ok))
    if rank == 0:
        out.put(all(res))
    dist.destroy_process_group()


def test_scatter_and_halos_world2_gloo():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_active_set_plane_cost():
    """plane_cost_active: planes without liquid (non-zero mode) / without fluid (fluid mode) cost 1, the others more, and the
    balanced cuts give the liquid planes to more ranks than an equal split would."""
    import scenes
    from solver.distributed import SlabPartition, balanced_starts, plane_cost_active
    sc = scenes.buckling(32)
    g = sc["gres"]
    for mode in ("nonzero", "fluid"):
        cost = plane_cost_active(sc["sphi"], sc["lvol"], g, mode)
        assert cost.shape == (g[0],) and cost.min() >= 1.0 and cost.max() > 1.0
        centres = sc["sphi"][1::2, 1::2, 1::2]
        outside = (centres >= 0).sum(dim=(1, 2)).numpy() == 0                      # planes entirely inside the solid
        if mode == "fluid":
            assert np.all(cost[outside] == 1.0)
        nz_planes = (sc["lvol"] != 0).sum(dim=(1, 2)).numpy()
        no_liquid = (nz_planes[0:-1:2] + nz_planes[1::2] + nz_planes[2::2]) == 0
        if mode == "nonzero":
            assert np.all(cost[no_liquid] == 1.0) and np.all(cost[~no_liquid] > 1.0)
        starts = balanced_starts(cost, 4)
        parts = [SlabPartition(g, 4, r, plane_cost=cost) for r in range(4)]
        assert [p.c0 for p in parts] + [g[0]] == starts
        sums = [cost[a:b].sum() for a, b in zip(starts, starts[1:])]
        eq = [cost[a:b].sum() for a, b in zip(range(0, 32, 8), range(8, 40, 8))]
        assert max(sums) <= max(eq) + 1e-9


def test_density_module_surface_without_gpu():
    """The drop-in module exposes the reference's names (DensityCGSolver3D.py:250-291, :283) and refuses to run without CUDA."""
    import inspect
    from solver import DensityCGSolver3D as Dn
    for name in ("initialize_density", "fix_volume", "initialize_solver", "matvecmul", "compute_displacement", "apply_displacement",
                 "DensityCGSolver3D", "compute_solid_frac", "edge_in_fraction"):
        assert hasattr(Dn, name), name
    sig = inspect.signature(Dn.DensityCGSolver3D.solve)
    assert list(sig.parameters)[1:] == ["rho0", "dt", "px", "pm", "pvol", "vx", "vy", "vz", "sphi", "sv", "lphi", "lvol", "wx", "wy", "wz", "tol"]
    assert sig.parameters["tol"].default == 1e-3 and sig.parameters["wx"].default is None
    assert list(inspect.signature(Dn.DensityCGSolver3D.__init__).parameters)[1:] == ["buf", "gres", "bound_min", "bound_size"]
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            Dn.DensityCGSolver3D(None, (8, 8, 8), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0))


def test_gathered_partition_windows():
    """ext=4 windows of the gathered solve: every window covers its owned cells plus 4 (clipped at the grid), owned planes
    tile the grid once, and the window arrays cut by slab() have the shapes the C ABI expects."""
    from solver.distributed import GatheredViscosityCGSolver3D, SlabPartition
    assert GatheredViscosityCGSolver3D.EXT == 4
    for nx, world in ((48, 2), (48, 3), (64, 8), (37, 4), (256, 8)):
        parts = [SlabPartition((nx, 5, 6), world, r, ext=4) for r in range(world)]
        for p in parts:
            assert p.e0 == max(0, p.c0 - 4) if p.has_lo else p.e0 == p.c0 == 0
            assert p.e1 == min(nx, p.c1 + 4) if p.has_hi else p.e1 == p.c1 == nx
            u = np.zeros((nx + 1, 5, 6))
            fine = np.zeros((2 * nx + 1, 11, 13))
            assert p.slab(u, "u").shape[0] == p.e1 - p.e0 + 1
            assert p.slab(u[:-1], "v").shape[0] == p.e1 - p.e0
            assert p.slab(fine, "fine").shape[0] == 2 * (p.e1 - p.e0) + 1
        covered = []
        for p in parts:
            lo, hi = p.owned_planes("u")
            covered += list(range(p.e0 + lo, p.e0 + hi))
        assert covered == list(range(nx + 1))


def test_new_module_surfaces_without_gpu():
    """notebook_kernels / sdf3D / unet_surrogate expose the reference's names and argument lists (ipynb cells 2-7, 12; sdf3D.py)"""
    import inspect
    import notebook_kernels as K
    import unet_surrogate as US
    from solver import sdf3D as sdf
    assert list(inspect.signature(K.p2g).parameters) == ["p", "g"]
    assert list(inspect.signature(K.g2p).parameters) == ["p", "g"]
    assert list(inspect.signature(K.compute_fluid_levelset).parameters) == ["p", "ls", "gdx"]
    assert list(inspect.signature(K.compute_fluid_volume).parameters) == ["p", "fv", "pvol"]
    assert list(inspect.signature(K.extrapolate).parameters) == ["gres", "num_iter", "vx", "vy", "vz", "mx", "my", "mz"]
    assert list(inspect.signature(K.apply_boundary_condition).parameters) == ["g", "solid", "dx"]
    assert list(inspect.signature(sdf.evaluate).parameters) == ["rb_d", "sd", "vel", "position"]
    assert list(inspect.signature(sdf.project).parameters) == ["rb_d", "position"]
    assert list(inspect.signature(sdf.generate_rb).parameters) == ["rb_d", "rb_map", "name", "rbparam", "flip", "center", "axis", "angle"]
    assert list(inspect.signature(US.unet_solve).parameters)[:5] == ["vx", "vy", "vz", "sphi", "lvol"] and len(inspect.signature(US.unet_solve).parameters) == 19
    assert US.default_data_size((48, 80, 48)) == (112, 176, 112)          # the notebook's hard-coded volume (ipynb cell 12)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            K.extrapolate((4, 4, 4), 1, *[np.zeros((5, 4, 4), dtype=np.float32)] * 6)
