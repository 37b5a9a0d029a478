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
