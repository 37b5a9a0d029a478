"""Phase timeline of the persistent CG kernel (block 0's global-timer stamps): python tests/phase_timeline.py [N] [grid]
Needs FLUIDSOLVER_B200_PROFILE=1 in the environment (set below).  Not a test; a measurement aid."""
import ctypes
import os
import sys

os.environ["FLUIDSOLVER_B200_PROFILE"] = "1"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "python-fluid-simulation_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import scenes  # noqa: E402
from solver import _native as N  # noqa: E402
from solver.ViscosityCGSolver3D import ViscosityCGSolver3D  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
if len(sys.argv) > 2:
    os.environ["FLUIDSOLVER_B200_PERSIST_GRID"] = sys.argv[2]
aset = sys.argv[3] if len(sys.argv) > 3 else "nonzero"
cgm = sys.argv[4] if len(sys.argv) > 4 else "persistent_sr"
lib = N.load()
sc = scenes.buckling(n, device="cuda", mu=100.0)
s = ViscosityCGSolver3D(sc["gres"], sc["bound_size"], cg_mode=cgm, active_set=aset)
s.max_iter = 0
v = [sc[k].clone() for k in ("vx", "vy", "vz")]
try:
    s.solve(sc["dt"], 100.0, sc["rho"], *v, sc["sphi"], None, None, sc["lvol"], tol=0.0)
except ValueError:
    pass
scale = sc["dt"] / s.cell_vol / sc["rho"]
for _ in range(3):
    N.check(lib.fs_visc3d_cg_enqueue(s._e.h, scale, 100.0, 64, 0), "warm")
torch.cuda.synchronize()
sr = cgm.endswith("_sr")
ns = 5 if sr else 7                          # stamps per iteration
buf = np.zeros(ns * 64 + 8, dtype=np.uint64)
N.check(lib.fs_visc3d_debug_read(s._e.h, 0, buf.ctypes.data_as(ctypes.c_void_p), buf.nbytes), "read")
t = buf[: ns * 64].astype(np.int64).reshape(64, ns)
d = np.diff(t, axis=1)[8:]                   # skip the first iterations
names = ["A apply+dots", "allreduce(2)", "B fused update", "barrier"] if sr else ["K1 body", "barrier1+dq", "K2 body", "barrier2+rr", "K3 body", "barrier3"]
per_it = np.diff(t[:, 0])[8:]
print(f"mode={cgm} N={n} grid={os.environ.get('FLUIDSOLVER_B200_PERSIST_GRID', 'auto')} active={aset} segments={s.active_info()[0]}  iteration {per_it.mean()/1e3:.2f} us")
for k, nm in enumerate(names):
    print(f"  {nm:14s} {d[:, k].mean()/1e3:7.2f} us  (min {d[:, k].min()/1e3:.2f}, max {d[:, k].max()/1e3:.2f})")
