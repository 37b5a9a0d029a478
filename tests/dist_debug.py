"""debug helper (2 GPUs): per-iteration CG scalars and halo consistency for both transports"""
import ctypes, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "python-fluid-simulation_b200")); sys.path.insert(0, REPO)
import torch, torch.distributed as dist

def main():
    local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    import scenes
    from solver import _native as N
    from solver.distributed import SlabPartition, SlabViscosityCGSolver3D, scatter_scene
    full = scenes.buckling(24, device="cuda", mu=10.0)
    gres = full["gres"]; part = SlabPartition(gres, world, rank); sc = scatter_scene(full, part)
    for transport in ("nccl", "p2p"):
        s = SlabViscosityCGSolver3D(gres, full["bound_size"], transport=transport)
        for k in range(0, 5):
            s.max_iter = k
            v = [sc[n].clone() for n in ("vx", "vy", "vz")]
            try:
                s.solve(full["dt"], 10.0, full["rho"], *v, sc["sphi"], None, None, sc["lvol"], tol=0.0)
            except ValueError:
                pass
            torch.cuda.synchronize()
            # halo consistency of r, d, q: my high halo plane (X-2) vs neighbour's plane 1
            msgs = []
            for nm in ("r", "d", "q", "x"):
                for c, ax in enumerate("xyz"):
                    a = getattr(s, f"{nm}_{ax}")
                    X = a.shape[0] + (0 if c == 0 else 1)     # lattice planes = nx_ext + 1
                    if part.has_hi:
                        mine = a[X - 2].contiguous() if X - 2 < a.shape[0] else None
                        other = torch.empty_like(mine); dist.recv(other, rank + 1)
                        msgs.append(f"{nm}{ax}:hi {float((mine - other).abs().max()):.1e}")
                    if part.has_lo:
                        dist.send(a[1].contiguous(), rank - 1)
            print(f"[{transport} r{rank}] k={k} iters={s.iterations} delta={s.delta:.6e} alpha={s.alpha:.6e} beta={s.beta:.6e} {' '.join(msgs)}", flush=True)
            dist.barrier()
        s.close(); del s
    dist.destroy_process_group()
main()
