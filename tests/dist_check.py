"""Multi-GPU parity check, launched as  python -m torch.distributed.run --nproc-per-node N tests/dist_check.py
(one rank per GPU, NCCL).  Every rank solves its slab of a buckling scene with SlabViscosityCGSolver3D; rank 0 also
solves the whole grid with the single-GPU ViscosityCGSolver3D and checks iteration count (+-2 %) and the owned
entries of every slab (1e-4 relative L2, gathered through torch.distributed).  Exits non-zero on mismatch."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "python-fluid-simulation_b200"))
sys.path.insert(0, REPO)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    import scenes
    from solver.distributed import GatheredViscosityCGSolver3D, SlabPartition, SlabViscosityCGSolver3D, scatter_scene
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D

    ok = True
    # (N, mu, dtype, gres, transport, cg_mode, active_set)
    cases = [(24, 10.0, torch.float64, None, "nccl", "auto", "nonzero"), (24, 10.0, torch.float64, None, "p2p", "persistent", "nonzero"),
             (24, 10.0, torch.float64, None, "p2p", "kernels", "fluid"),
             (32, 100.0, torch.float64, (37, 24, 28), "p2p", "persistent", "fluid"), (32, 100.0, torch.float64, (37, 24, 28), "p2p", "kernels", "nonzero"),
             (32, 100.0, torch.float32, None, "p2p", "auto", "nonzero"), (32, 100.0, torch.float64, (37, 24, 28), "p2p", "persistent_sr", "nonzero"),
             (32, 100.0, torch.float64, (37, 24, 28), "p2p", "kernels_sr", "fluid"),
             (32, 100.0, torch.float64, (37, 24, 28), "nccl", "auto", "fluid"),
             # gathered solve: set-up sharded, records all-gathered over NCCL, CG replicated (transport column = "gathered")
             (48, 100.0, torch.float64, None, "gathered", "auto", "nonzero"), (32, 10.0, torch.float64, (67, 24, 28), "gathered", "auto", "fluid"),
             (48, 100.0, torch.float32, None, "gathered", "auto", "nonzero")]
    repeat = 1
    if os.environ.get("DIST_CHECK_BIG"):      # diagnostic sizes (HBM-sized slabs, several segments per warp), two solves per object
        big = int(os.environ["DIST_CHECK_BIG"])
        cases = [(big, 100.0, torch.float64, None, "p2p", m, "fluid") for m in ("persistent_sr", "persistent", "kernels_sr", "kernels")]
        cases.append((big, 100.0, torch.float64, None, "gathered", "auto", "nonzero"))
        repeat = 2
    for N, mu, dtype, g, transport, cg_mode, aset in cases:
        full = scenes.buckling(N, device="cuda", mu=mu, gres=g)
        gres = full["gres"]
        if transport == "gathered":
            if gres[0] < 2 * world:
                continue
            part = SlabPartition(gres, world, rank, ext=4)
            sc = scatter_scene(full, part)
            s = GatheredViscosityCGSolver3D(gres, full["bound_size"], dtype=dtype, partition=part, cg_mode=cg_mode, active_set=aset)
        else:
            cost = None
            if os.environ.get("DIST_CHECK_BALANCED"):     # the cost-balanced (uneven) cuts bench.py uses
                from solver.distributed import plane_cost_active
                cost = plane_cost_active(full["sphi"], full["lvol"], gres, aset)
            part = SlabPartition(gres, world, rank, plane_cost=cost)
            if rank == 0 and cost is not None:
                print(f"[dist_check] slab starts {part.starts}", flush=True)
            sc = scatter_scene(full, part)
            s = SlabViscosityCGSolver3D(gres, full["bound_size"], dtype=dtype, transport=transport, cg_mode=cg_mode, active_set=aset, partition=part)
        if os.environ.get("DIST_CHECK_WINDOWS"):          # bench.py's pattern: fixed 200-iteration windows (tol = 0) first
            s.max_iter = 200
            for _ in range(int(os.environ["DIST_CHECK_WINDOWS"])):
                try:
                    s.solve(full["dt"], mu, full["rho"], sc["vx"], sc["vy"], sc["vz"], sc["sphi"], None, None, sc["lvol"], tol=0.0)
                except ValueError:
                    pass
        s.max_iter = 5000
        for _ in range(repeat):
            v = [sc[k].clone() for k in ("vx", "vy", "vz")]
            try:
                s.solve(full["dt"], mu, full["rho"], *v, sc["sphi"], None, None, sc["lvol"])
            except ValueError:
                pass                              # not converged within the cap: reported as a mismatch below
        its = torch.tensor([s.iterations], device="cuda")
        allits = [torch.zeros_like(its) for _ in range(world)]
        dist.all_gather(allits, its)
        # gather owned planes on rank 0
        pieces = []
        for a, kind in zip(v, ("u", "v", "w")):
            lo, hi = part.owned_planes(kind)
            own = a[lo:hi].contiguous()
            shapes = [None] * world
            dist.all_gather_object(shapes, tuple(own.shape))
            bufs = [torch.empty(sh, dtype=own.dtype, device="cuda") for sh in shapes]
            dist.all_gather(bufs, own) if len({tuple(b.shape) for b in bufs}) == 1 else _gather_uneven(bufs, own, rank, world)
            pieces.append(torch.cat(bufs, dim=0))
        if rank == 0:
            ref = ViscosityCGSolver3D(gres, full["bound_size"], dtype=dtype)
            rv = [full[k].clone() for k in ("vx", "vy", "vz")]
            ref.solve(full["dt"], mu, full["rho"], *rv, full["sphi"], None, None, full["lvol"])
            same_it = all(int(t.item()) == s.iterations for t in allits)
            it_ok = abs(s.iterations - ref.iterations) <= max(1, round(0.02 * ref.iterations))
            errs = [float((a.double() - b.double()).norm() / b.double().norm()) for a, b in zip(pieces, rv)]
            shp_ok = all(a.shape == b.shape for a, b in zip(pieces, rv))
            good = same_it and it_ok and shp_ok and max(errs) < 1e-4
            ok = ok and good
            print(f"[dist_check] {transport}/{cg_mode}/{aset} gres={gres} mu={mu} {dtype}: world={world} iters={s.iterations} (single-GPU {ref.iterations}) "
                  f"lockstep={same_it} rel_l2={['%.2e' % e for e in errs]} -> {'OK' if good else 'MISMATCH'}", flush=True)
        dist.barrier()
        s.close()
        del s
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


def _gather_uneven(bufs, own, rank, world):
    for r in range(world):
        if r == rank:
            bufs[r].copy_(own)
        dist.broadcast(bufs[r], r)


if __name__ == "__main__":
    main()
