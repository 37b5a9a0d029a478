"""A/B of the result-equivalent kernel forms behind fs_set_option, in ONE process on one GPU (a measurement aid, not a test):

    python tests/ab_forms.py [N]

  * default scene (buckling N^3, mu=100): per-iteration time of the persistent CG window, whole-step rate, phase timeline and
    the iterate after a fixed 200-iteration solve for resident_form in {0, 2};
  * dense scene (viscous column N^3, fluid rows) and the benchmark scene with every fluid row visited: K1s alone and the
    iteration for the forms of the stand-alone apply (k1_block, k1_tile).
One JSON line per configuration on stdout."""
import ctypes
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "python-fluid-simulation_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import scenes  # noqa: E402
from solver import _native as N  # noqa: E402
from solver.ViscosityCGSolver3D import ViscosityCGSolver3D  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
MU = 100.0
lib = N.load()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
stream = lambda: torch.cuda.current_stream().cuda_stream


def emit(**kw):
    print(json.dumps(kw), flush=True)


def fixed_solve(s, sc, iters):
    s.max_iter = iters
    v = [sc[k].clone() for k in ("vx", "vy", "vz")]
    try:
        s.solve(sc["dt"], MU, sc["rho"], *v, sc["sphi"], None, None, sc["lvol"], tol=0.0)
    except ValueError:
        pass
    return v


def timed(fn, reps):
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps


def timeline(s, scale):
    os.environ["FLUIDSOLVER_B200_PROFILE"] = "1"
    try:
        N.check(lib.fs_visc3d_cg_enqueue(s._e.h, scale, MU, 64, stream()), "profiled window")
        torch.cuda.synchronize()
        buf = np.zeros(5 * 64 + 8, dtype=np.uint64)
        N.check(lib.fs_visc3d_debug_read(s._e.h, 0, buf.ctypes.data_as(ctypes.c_void_p), buf.nbytes), "read")
        t = buf[: 5 * 64].astype(np.int64).reshape(64, 5)
        d = np.diff(t, axis=1)[8:]
        return {"A": d[:, 0].mean() / 1e3, "reduce": d[:, 1].mean() / 1e3, "B": d[:, 2].mean() / 1e3, "barrier": d[:, 3].mean() / 1e3,
                "iteration": float(np.diff(t[:, 0])[8:].mean() / 1e3)}
    finally:
        del os.environ["FLUIDSOLVER_B200_PROFILE"]


# ---- default scene: the persistent CG forms --------------------------------------------------------------------
sc = scenes.buckling(n, device="cuda", mu=MU)
s = ViscosityCGSolver3D(sc["gres"], sc["bound_size"], cg_mode="persistent_sr")
scale = sc["dt"] / s.cell_vol / sc["rho"]
ref = None
for form in (2, 0, 2):
    try:
        N.set_option("resident_form", form)
        fixed_solve(s, sc, 200)
        it, delta = s.iterations, float(s.delta)
        x = [a.double().cpu().numpy().copy() for a in (s.x_x, s.x_y, s.x_z)]
        if ref is None:
            ref = (delta, x)
        err = max(float(np.linalg.norm(a - b) / np.linalg.norm(b)) for a, b in zip(x, ref[1]))
        N.check(lib.fs_visc3d_cg_enqueue(s._e.h, scale, MU, 64, stream()), "warm")
        ms_win = timed(lambda: N.check(lib.fs_visc3d_cg_enqueue(s._e.h, scale, MU, 512, stream()), "window"), 2)
        ms_step = timed(lambda: fixed_solve(s, sc, 200), 5)
        tl = timeline(s, scale)
        emit(scene=f"buckling-{n}", resident_form=form, iterations=it, delta=delta, delta_rel_vs_first=abs(delta - ref[0]) / abs(ref[0]),
             x_rel_l2_vs_first=err, cg_iteration_us=1e3 * ms_win / 512, ms_per_step=ms_step, iters_per_s=200 / (ms_step * 1e-3), phases_us=tl,
             segments=s.active_info()[0])
    except Exception as e:                                      # keep going: the other forms are still worth their numbers
        emit(scene=f"buckling-{n}", resident_form=form, error=repr(e))
N.set_option("resident_form", -1)
for sp in (0, 1, 0, 1):                                  # dense / sparse set-up of solve() on the default forms
    try:
        N.set_option("sparse_setup", sp)
        v = fixed_solve(s, sc, 200)
        d200 = float(s.delta)
        r = [a.clone() for a in (s.r_x, s.r_y, s.r_z)]
        if sp == 0:
            ref_r, ref_d = r, d200
        ms_step = timed(lambda: fixed_solve(s, sc, 200), 10)
        emit(scene=f"buckling-{n}", sparse_setup=sp, delta=d200, delta_equal_dense=(d200 == ref_d),
             r_equal_dense=all(bool(torch.equal(a, b)) for a, b in zip(r, ref_r)), ms_per_step=ms_step, iters_per_s=200 / (ms_step * 1e-3))
    except Exception as e:
        emit(scene=f"buckling-{n}", sparse_setup=sp, error=repr(e))
N.set_option("sparse_setup", -1)
del s, sc
torch.cuda.empty_cache()

# ---- dense scene: K1s with / without the L2 prefetch -----------------------------------------------------------------
col = scenes.viscous_column((n, n, n), device="cuda", mu=MU)
s = ViscosityCGSolver3D(col["gres"], col["bound_size"], cg_mode="kernels_sr", active_set="fluid")
scale = col["dt"] / s.cell_vol / col["rho"]
F = 3 * n * n * (n + 1)
V7 = F + n ** 3 + 3 * (n + 1) * (n + 1) * n
ref = None
T = lambda rows, planes: 1000 * planes + 10 * rows + 1
for blk, tile in ((0, 0), (4, 0), (0, 1), (0, T(1, 16)), (0, T(2, 8)), (0, T(2, 32)), (0, T(2, 64)), (0, 1), (4, 0)):
    try:
        N.set_option("k1_block", blk)
        N.set_option("k1_tile", tile)
        fixed_solve(s, col, 20)
        delta = float(s.delta)
        x = [a.double().cpu().numpy().copy() for a in (s.x_x, s.x_y, s.x_z)]
        if ref is None:
            ref = (delta, x)
        err = max(float(np.linalg.norm(a - b) / np.linalg.norm(b)) for a, b in zip(x, ref[1]))
        k = {}
        for which, name in ((1, "K1s"), (2, "K2s")):
            N.check(lib.fs_visc3d_kernel_enqueue(s._e.h, which, scale, MU, 3, stream()), "warm")
            k[name] = timed(lambda: N.check(lib.fs_visc3d_kernel_enqueue(s._e.h, which, scale, MU, 30, stream()), "time"), 1) / 30
        N.check(lib.fs_visc3d_cg_enqueue(s._e.h, scale, MU, 8, stream()), "warm")
        it_ms = timed(lambda: N.check(lib.fs_visc3d_cg_enqueue(s._e.h, scale, MU, 30, stream()), "window"), 1) / 30
        emit(scene=f"column-{n}", k1_block=blk, k1_tile=tile, delta_rel_vs_first=abs(delta - ref[0]) / abs(ref[0]), x_rel_l2_vs_first=err,
             K1s_ms=k["K1s"], K2s_ms=k["K2s"], iteration_ms=it_ms, K1s_frac_of_peak=(2 * F + V7) * 8 / (k["K1s"] * 1e-3) / 1e9 / 6548.2,
             iteration_frac_of_peak=(11 * F + V7) * 8 / (it_ms * 1e-3) / 1e9 / 6548.2, segments=s.active_info()[0])
    except Exception as e:
        emit(scene=f"column-{n}", k1_block=blk, k1_tile=tile, error=repr(e))
N.set_option("k1_block", -1)
N.set_option("k1_tile", -1)
del s, col
torch.cuda.empty_cache()

# ---- buckling scene, every fluid row (20 % of the segments, runs of 1-1.5 KB): the same mappings --------------------------
sc = scenes.buckling(n, device="cuda", mu=MU)
s = ViscosityCGSolver3D(sc["gres"], sc["bound_size"], cg_mode="kernels_sr", active_set="fluid")
scale = sc["dt"] / s.cell_vol / sc["rho"]
for blk in (0, 2, 4, 6, 0):
    try:
        N.set_option("k1_block", blk)
        fixed_solve(s, sc, 20)
        segs = s.active_info()[0]
        N.check(lib.fs_visc3d_kernel_enqueue(s._e.h, 1, scale, MU, 3, stream()), "warm")
        k1 = timed(lambda: N.check(lib.fs_visc3d_kernel_enqueue(s._e.h, 1, scale, MU, 30, stream()), "time"), 1) / 30
        N.check(lib.fs_visc3d_cg_enqueue(s._e.h, scale, MU, 8, stream()), "warm")
        it_ms = timed(lambda: N.check(lib.fs_visc3d_cg_enqueue(s._e.h, scale, MU, 60, stream()), "window"), 1) / 60
        emit(scene=f"buckling-{n}-fluid", k1_block=blk, delta=float(s.delta), K1s_ms=k1, iteration_ms=it_ms, segments=segs,
             K1s_frac_of_peak_active_bytes=segs * 32 * (13 * 8 + 1) / (k1 * 1e-3) / 1e9 / 6548.2)
    except Exception as e:
        emit(scene=f"buckling-{n}-fluid", k1_block=blk, error=repr(e))
N.set_option("k1_block", -1)
