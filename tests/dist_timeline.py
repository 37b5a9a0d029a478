"""Phase timeline of the persistent CG kernel on x-slabs:
python -m torch.distributed.run --nproc-per-node N tests/dist_timeline.py [size] [active_set].  Measurement aid, not a test."""
import ctypes
import os
import sys

os.environ["FLUIDSOLVER_B200_PROFILE"] = "1"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "python-fluid-simulation_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
import scenes  # noqa: E402
from solver import _native as N  # noqa: E402
from solver.distributed import SlabPartition, SlabViscosityCGSolver3D, plane_cost_active, scatter_scene  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
aset = sys.argv[2] if len(sys.argv) > 2 else "nonzero"
lib = N.load()
full = scenes.buckling(n, device="cuda", mu=100.0)
part = SlabPartition(full["gres"], world, rank, plane_cost=plane_cost_active(full["sphi"], full["lvol"], full["gres"], aset))
sc = scatter_scene(full, part)
s = SlabViscosityCGSolver3D(full["gres"], full["bound_size"], partition=part, cg_mode="persistent", active_set=aset)
s.max_iter = 0
v = [sc[k].clone() for k in ("vx", "vy", "vz")]
try:
    s.solve(full["dt"], 100.0, full["rho"], *v, sc["sphi"], None, None, sc["lvol"], tol=0.0)
except ValueError:
    pass
scale = full["dt"] / s.cell_vol / full["rho"]
for _ in range(3):
    dist.barrier()
    N.check(lib.fs_visc3d_cg_enqueue(s._e.h, scale, 100.0, 64, 0), "warm")
    torch.cuda.synchronize()
buf = np.zeros(7 * 64 + 8, dtype=np.uint64)
N.check(lib.fs_visc3d_debug_read(s._e.h, 0, buf.ctypes.data_as(ctypes.c_void_p), buf.nbytes), "read")
t = buf[: 7 * 64].astype(np.int64).reshape(64, 7)
d = np.diff(t, axis=1)[8:]
names = ["K1 body", "allreduce d.q", "K2 body", "allreduce r.r", "K3 body", "grid sync"]
per_it = np.diff(t[:, 0])[8:]
for r in range(world):
    dist.barrier()
    if r == rank:
        print(f"[rank {rank}/{world}] N={n} active={aset} segments={s._e.active_info()[0]} slab={part.starts}  iteration {per_it.mean()/1e3:.2f} us", flush=True)
        for k, nm in enumerate(names):
            print(f"    {nm:14s} {d[:, k].mean()/1e3:7.2f} us  (min {d[:, k].min()/1e3:.2f}, max {d[:, k].max()/1e3:.2f})", flush=True)
dist.barrier()
s.close()
dist.destroy_process_group()
