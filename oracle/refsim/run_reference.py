#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — generate golden fixtures by executing the UNMODIFIED reference.

Runs ``/root/reference/solver/*.py`` (Numba ``@cuda.jit`` kernels + CuPy vector algebra) on the
CPU under Numba's CUDA simulator with the NumPy-backed ``cupy`` shim next to this file, and
writes inputs + outputs to ``tests/golden/*.npz``.  The reference ships no tests or golden vectors
of its own, so these files ARE the parity pins.  ``/root/reference`` exists only in
the build container: this script is never run on the GPU box; the committed ``.npz`` files travel.

Usage:  python oracle/refsim/run_reference.py [case ...]      (no args = all cases)
"""
import os
import sys
import time

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("FLUID_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, HERE)      # the cupy shim
sys.path.insert(0, REF)       # the reference's `solver` package

import numpy as np  # noqa: E402
import cupy as cp  # noqa: E402  (the shim)

GOLD = os.path.join(REPO, "tests", "golden")


def _save(name, **arrs):
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrs.items()})
    print(f"  wrote {path} ({os.path.getsize(path)/1024:.1f} KiB)")


def _c(a):
    return cp.asarray(np.array(a, copy=True))


def _mac_shapes(g):
    d = len(g)
    return [tuple(n + (1 if i == a else 0) for i, n in enumerate(g)) for a in range(d)]


def _fine(g):
    return tuple(2 * n + 1 for n in g)


# ---------------------------------------------------------------------------------------------
# small deterministic scene used for the full-solve fixtures (container + liquid column + pool)
# ---------------------------------------------------------------------------------------------

def small_scene(g, dx, seed, wall=1.5, col_r=2.0, pool=2.0, top_gap=2.0):
    """Box container (solid outside), liquid column on the axis + pool on the floor.

    Returns fp64 sphi, lvol on the (2g+1) fine grid, lphi on cells, fp32 MAC velocities.
    Pure NumPy, no reference code involved (inputs only).
    """
    g = tuple(int(n) for n in g)
    d = len(g)
    L = [n * dx for n in g]
    ax = [np.arange(2 * n + 1) * (dx / 2) for n in g]
    P = np.meshgrid(*ax, indexing="ij")
    # container: inside of an axis-aligned box inset `wall` cells from the domain boundary
    lo = [wall * dx] * d
    hi = [L[i] - wall * dx for i in range(d)]
    inside = np.minimum.reduce([np.minimum(P[i] - lo[i], hi[i] - P[i]) for i in range(d)])
    sphi = inside.astype(np.float64)          # >0 inside the container (fluid region), <0 in the wall
    # liquid: column along axis 1 (y) + pool
    cen = [L[i] / 2 for i in range(d)]
    if d == 3:
        r = np.sqrt((P[0] - cen[0]) ** 2 + (P[2] - cen[2]) ** 2)
    else:
        r = np.abs(P[0] - cen[0])
    col = np.maximum(r - col_r * dx, P[1] - (L[1] - (wall + top_gap) * dx))
    poolphi = P[1] - (wall + pool) * dx
    lphi_f = np.minimum(col, poolphi)
    lvol = np.clip(0.5 - lphi_f / (dx / 2), 0.0, 1.0) * (dx / 2) ** d * (sphi > 0)
    cs = tuple(slice(1, None, 2) for _ in range(d))
    lphi = np.ascontiguousarray(lphi_f[cs])
    rng = np.random.default_rng(seed)
    vel = []
    for a, sh in enumerate(_mac_shapes(g)):
        base = -2.0 if a == 1 else 0.0
        vel.append((base + 0.1 * rng.standard_normal(sh)).astype(np.float32))
    return sphi, lvol.astype(np.float64), lphi.astype(np.float64), vel


# ---------------------------------------------------------------------------------------------
# cases
# ---------------------------------------------------------------------------------------------

def case_visc3d_kernels():
    """Single apply / RHS / extrapolation / write-back on random inputs, non-cubic grid."""
    from solver import ViscosityCGSolver3D as V
    g = (6, 7, 8)
    rng = np.random.default_rng(101)
    sphi = rng.standard_normal(_fine(g))
    sphi[rng.random(_fine(g)) < 0.03] = 0.0         # exact zeros: exercises >=0 vs <0
    vol = rng.random(_fine(g))
    vel = [rng.standard_normal(s) for s in _mac_shapes(g)]
    scale, mu = 0.7, 1.3
    gres = cp.array(g, dtype=cp.int64)
    # NaN-poisoned outputs show exactly which entries the reference writes
    q = [cp.asarray(np.full(s, np.nan)) for s in _mac_shapes(g)]
    V.matvecmul(gres, scale, mu, *[_c(v) for v in vel], *q, _c(sphi), _c(vol))
    b = [cp.asarray(np.full(s, np.nan)) for s in _mac_shapes(g)]
    V.initialize_solver(gres, scale, mu, *[_c(v) for v in vel], _c(sphi), None, _c(vol), *b)
    ex = [_c(v) for v in vel]
    V.extrapolate(gres, 3, *ex, _c(sphi))
    wb = [cp.asarray(np.full(s, np.nan, dtype=np.float32)) for s in _mac_shapes(g)]
    V.apply_viscosity(gres, *wb, *[_c(v) for v in vel], _c(sphi), None)
    _save("visc3d_kernels_6x7x8", gres=np.array(g), scale=scale, mu=mu, sphi=sphi, vol=vol,
          vx=vel[0], vy=vel[1], vz=vel[2], qx=q[0], qy=q[1], qz=q[2], bx=b[0], by=b[1], bz=b[2],
          ex=ex[0], ey=ex[1], ez=ex[2], wbx=wb[0], wby=wb[1], wbz=wb[2])


def _count_calls(mod, name):
    orig = getattr(mod, name)
    box = {"n": 0}

    def wrapped(*a, **k):
        box["n"] += 1
        return orig(*a, **k)

    setattr(mod, name, wrapped)
    return box, lambda: setattr(mod, name, orig)


def case_visc3d_solve(tag="8x10x8", g=(8, 10, 8), mu=1.0, seed=42, tol=1e-3):
    """Full ViscosityCGSolver3D.solve on the small container scene (fp32 caller velocities)."""
    from solver import ViscosityCGSolver3D as V
    dx = 0.0125
    sphi, lvol, lphi, vel = small_scene(g, dx, seed)
    gres = cp.array(g, dtype=cp.int64)
    bound = cp.array([n * dx for n in g], dtype=cp.float32)
    s = V.ViscosityCGSolver3D(gres, bound)
    out = [_c(v) for v in vel]
    sv = cp.zeros(_fine(g) + (3,))
    box, restore = _count_calls(V, "matvecmul")
    dt, rho = 1.0 / 300, 1000.0
    t = time.time()
    s.solve(dt, mu, rho, *out, _c(sphi), sv, _c(lphi), _c(lvol), tol=tol)
    restore()
    iters = box["n"] - 1
    print(f"  visc3d solve {tag}: {iters} iterations, delta={s.delta:.3e}, {time.time()-t:.0f}s")
    _save(f"visc3d_solve_{tag}", gres=np.array(g), bound_size=np.asarray(bound), dt=dt, mu=mu, rho=rho, tol=tol,
          sphi=sphi, lvol=lvol, lphi=lphi, vx=vel[0], vy=vel[1], vz=vel[2],
          vx_new=out[0], vy_new=out[1], vz_new=out[2], x_x=s.x_x, x_y=s.x_y, x_z=s.x_z,
          b_x=s.b_x, b_y=s.b_y, b_z=s.b_z, iterations=iters, delta=s.delta, alpha=s.alpha, beta=s.beta,
          cell_vol=s.cell_vol)


def case_visc3d_solve_stiff():
    case_visc3d_solve(tag="stiff_6x8x6", g=(6, 8, 6), mu=50.0, seed=43)


def case_solidfrac3d():
    from solver.SolidFraction3D import compute_solid_frac
    g = (6, 7, 8)
    rng = np.random.default_rng(202)
    sphi = rng.standard_normal(_fine(g))
    sphi[rng.random(_fine(g)) < 0.05] = 0.0
    # a smooth field too (plane + sphere), where partially-inside faces occur in runs
    ax = [np.linspace(-1, 1, 2 * n + 1) for n in g]
    X, Y, Z = np.meshgrid(*ax, indexing="ij")
    smooth = np.minimum(np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - 0.55, 0.4 * X + 0.3 * Y - 0.2 * Z + 0.1)
    outs = {}
    for nm, f in (("rand", sphi), ("smooth", smooth)):
        w = [cp.asarray(np.full(s, np.nan)) for s in _mac_shapes(g)]
        compute_solid_frac(cp.array(g, dtype=cp.int64), _c(f), *w)
        outs.update({f"sphi_{nm}": f, f"wx_{nm}": w[0], f"wy_{nm}": w[1], f"wz_{nm}": w[2]})
    _save("solidfrac3d_6x7x8", gres=np.array(g), **outs)


def case_solidfrac2d():
    from solver.SolidFraction2D import compute_solid_frac
    g = (9, 7)
    rng = np.random.default_rng(203)
    sphi = rng.standard_normal(_fine(g))
    sphi[rng.random(_fine(g)) < 0.05] = 0.0
    w = [cp.asarray(np.full(s, np.nan)) for s in _mac_shapes(g)]
    compute_solid_frac(cp.array(g, dtype=cp.int64), _c(sphi), *w)
    _save("solidfrac2d_9x7", gres=np.array(g), sphi=sphi, wx=w[0], wy=w[1])


def _press_random_inputs(g, seed):
    rng = np.random.default_rng(seed)
    d = len(g)
    lphi = rng.standard_normal(g) - 0.3
    lphi[rng.random(g) < 0.03] = 0.0
    ws = [rng.choice([0.0, 0.5, 0.75, 1.0], size=s, p=[0.1, 0.1, 0.1, 0.7]) for s in _mac_shapes(g)]
    vel = [rng.standard_normal(s).astype(np.float32) for s in _mac_shapes(g)]
    sv = rng.standard_normal(_fine(g) + (d,))
    pv = rng.standard_normal(g)
    return lphi, ws, vel, sv, pv


def case_press3d_kernels():
    from solver import PressureCGSolver3D as P
    g = (7, 6, 8)
    lphi, ws, vel, sv, pv = _press_random_inputs(g, 301)
    gres = cp.array(g, dtype=cp.int64)
    cell = cp.array([0.011, 0.013, 0.017], dtype=cp.float64)
    q = cp.asarray(np.full(g, np.nan))
    P.matvecmul(gres, _c(pv), q, *[_c(w) for w in ws], _c(lphi))
    b = cp.asarray(np.full(g, np.nan))
    P.initialize_solver(cell, gres, *[_c(v) for v in vel], None, _c(sv), _c(lphi), b, *[_c(w) for w in ws])
    upd = [_c(v) for v in vel]
    P.apply_pressure(gres, cell, *upd, _c(pv), *[_c(w) for w in ws], _c(sv), _c(lphi))
    _save("press3d_kernels_7x6x8", gres=np.array(g), cell_size=np.asarray(cell), lphi=lphi, wx=ws[0], wy=ws[1], wz=ws[2],
          vx=vel[0], vy=vel[1], vz=vel[2], sv=sv, pv=pv, q=q, b=b, ux=upd[0], uy=upd[1], uz=upd[2])


def case_press3d_solve():
    from solver import PressureCGSolver3D as P
    from solver.CGSolverBuffer import CGSolverBuffer
    g = (8, 10, 8)
    dx = 0.0125
    sphi, lvol, lphi, vel = small_scene(g, dx, 44)
    rng = np.random.default_rng(45)
    sv = np.zeros(_fine(g) + (3,))
    sv[..., 1] = 0.05 * rng.standard_normal(_fine(g))     # exercise the solid-velocity terms
    gres = cp.array(g, dtype=cp.int64)
    buf = CGSolverBuffer(gres)
    s = P.PressureCGSolver3D(buf, gres, dx)               # scalar bound_size = GDX, as the notebook does
    out = [_c(v) for v in vel]
    box, restore = _count_calls(P, "matvecmul")
    t = time.time()
    s.solve(*out, _c(sphi), _c(sv), _c(lphi), tol=1e-3)
    restore()
    iters = box["n"] - 1
    print(f"  press3d solve: {iters} iterations, delta={s.delta:.3e}, {time.time()-t:.0f}s")
    _save("press3d_solve_8x10x8", gres=np.array(g), bound_size=dx, tol=1e-3, sphi=sphi, sv=sv, lphi=lphi,
          vx=vel[0], vy=vel[1], vz=vel[2], vx_new=out[0], vy_new=out[1], vz_new=out[2],
          x=s.x, wx=s.wx, wy=s.wy, wz=s.wz, b=buf.b, iterations=iters, delta=s.delta, alpha=s.alpha, beta=s.beta)


def case_visc2d_kernels():
    from solver import ViscosityCGSolver2D as V
    g = (9, 7)
    rng = np.random.default_rng(401)
    sphi = rng.standard_normal(_fine(g))
    sphi[rng.random(_fine(g)) < 0.05] = 0.0          # exact zeros: solid in 2-D (<=0)
    vol = rng.random(_fine(g))
    vel = [rng.standard_normal(s) for s in _mac_shapes(g)]
    scale, mu = 0.9, 0.6
    gres = cp.array(g, dtype=cp.int64)
    q = [cp.asarray(np.full(s, np.nan)) for s in _mac_shapes(g)]
    V.matvecmul(gres, scale, mu, *[_c(v) for v in vel], *q, _c(sphi), _c(vol))
    b = [cp.asarray(np.full(s, np.nan)) for s in _mac_shapes(g)]
    V.initialize_solver(gres, scale, mu, *[_c(v) for v in vel], _c(sphi), None, _c(vol), *b)
    wb = [cp.asarray(np.full(s, np.nan, dtype=np.float32)) for s in _mac_shapes(g)]
    V.apply_viscosity(gres, *wb, *[_c(v) for v in vel], _c(sphi), None)
    _save("visc2d_kernels_9x7", gres=np.array(g), scale=scale, mu=mu, sphi=sphi, vol=vol, vx=vel[0], vy=vel[1],
          qx=q[0], qy=q[1], bx=b[0], by=b[1], wbx=wb[0], wby=wb[1])


def case_visc2d_solve():
    from solver import ViscosityCGSolver2D as V
    g = (14, 12)
    dx = 1.0 / 16
    sphi, lvol, lphi, vel = small_scene(g, dx, 46)
    gres = cp.array(g, dtype=cp.int64)
    bound = cp.array([n * dx for n in g], dtype=cp.float32)
    s = V.ViscosityCGSolver2D(gres, bound)
    out = [_c(v) for v in vel]
    sv = cp.zeros(_fine(g) + (2,))
    box, restore = _count_calls(V, "matvecmul")
    dt, mu, rho = 1.0 / 300, 1.0, 1000.0
    s.solve(dt, mu, rho, *out, _c(sphi), sv, _c(lphi), _c(lvol))
    restore()
    iters = box["n"] - 1
    print(f"  visc2d solve: {iters} iterations, delta={s.delta:.3e}")
    _save("visc2d_solve_14x12", gres=np.array(g), bound_size=np.asarray(bound), dt=dt, mu=mu, rho=rho, tol=1e-4,
          sphi=sphi, lvol=lvol, lphi=lphi, vx=vel[0], vy=vel[1], vx_new=out[0], vy_new=out[1],
          x_x=s.x_x, x_y=s.x_y, iterations=iters, delta=s.delta, cell_vol=s.cell_vol)


def case_press2d_kernels():
    from solver import PressureCGSolver2D as P
    g = (9, 7)
    lphi, ws, vel, sv, pv = _press_random_inputs(g, 501)
    ws = [np.random.default_rng(502 + i).random(w.shape) for i, w in enumerate(ws)]   # continuous weights in 2-D
    gres = cp.array(g, dtype=cp.int64)
    cell = cp.array([0.011, 0.013], dtype=cp.float64)
    q = cp.asarray(np.full(g, np.nan))
    P.matvecmul(gres, _c(pv), q, *[_c(w) for w in ws], _c(lphi))
    b = cp.asarray(np.full(g, np.nan))
    P.initialize_solver(cell, gres, *[_c(v) for v in vel], None, _c(sv), _c(lphi), b, *[_c(w) for w in ws])
    upd = [_c(v) for v in vel]
    P.apply_pressure(gres, cell, *upd, _c(pv), *[_c(w) for w in ws], _c(sv), _c(lphi))
    _save("press2d_kernels_9x7", gres=np.array(g), cell_size=np.asarray(cell), lphi=lphi, wx=ws[0], wy=ws[1],
          vx=vel[0], vy=vel[1], sv=sv, pv=pv, q=q, b=b, ux=upd[0], uy=upd[1])


def case_press2d_solve():
    from solver import PressureCGSolver2D as P
    from solver.CGSolverBuffer import CGSolverBuffer
    g = (14, 12)
    dx = 1.0 / 16
    sphi, lvol, lphi, vel = small_scene(g, dx, 47)
    sv = np.zeros(_fine(g) + (2,))
    gres = cp.array(g, dtype=cp.int64)
    buf = CGSolverBuffer(gres)
    s = P.PressureCGSolver2D(buf, gres, cp.array([n * dx for n in g], dtype=cp.float32))
    out = [_c(v) for v in vel]
    box, restore = _count_calls(P, "matvecmul")
    s.solve(*out, _c(sphi), _c(sv), _c(lphi), tol=1e-3)
    restore()
    iters = box["n"] - 1
    print(f"  press2d solve: {iters} iterations, delta={s.delta:.3e}")
    _save("press2d_solve_14x12", gres=np.array(g), bound_size=np.asarray([n * dx for n in g], dtype=np.float32), tol=1e-3,
          sphi=sphi, sv=sv, lphi=lphi, vx=vel[0], vy=vel[1], vx_new=out[0], vy_new=out[1], x=s.x, wx=s.wx, wy=s.wy,
          b=buf.b, iterations=iters, delta=s.delta)


def _density_inputs(g, dx, seed, n_per_cell=2):
    """particles jittered inside the liquid of small_scene + the grid fields the density solve needs"""
    sphi, lvol, lphi, vel = small_scene(g, dx, seed)
    rng = np.random.default_rng(seed + 1000)
    cells = np.argwhere(lphi < 0.5 * dx)
    reps = np.repeat(cells, n_per_cell, axis=0)
    px = (reps + rng.random(reps.shape)) * dx
    px = np.clip(px, 0.02 * dx, (np.array(g) - 0.02) * dx)
    pm = np.full(px.shape[0], 1000.0 * (dx / 2) ** 3) * (1 + 0.1 * rng.standard_normal(px.shape[0]))
    pvol = (dx / 2) ** 3
    return sphi, lvol, lphi, vel, px, pm, pvol


def case_density3d_kernels():
    from solver import DensityCGSolver3D as Dn
    from solver.SolidFraction3D import compute_solid_frac
    g = (7, 8, 6)
    dx = 0.05
    sphi, lvol, lphi, vel, px, pm, pvol = _density_inputs(g, dx, 61)
    gres = cp.array(g, dtype=cp.int64)
    bound_min = cp.array([0.0, 0.0, 0.0])
    cell = cp.array([dx, dx, dx])
    ws = [cp.zeros(s) for s in _mac_shapes(g)]
    compute_solid_frac(gres, _c(sphi), *ws)
    gm, gvol = cp.zeros(g), cp.zeros(g)
    Dn.initialize_density(bound_min, cell, gres, _c(px), _c(pm), pvol, gm, gvol, _c(sphi), _c(lphi))
    gvol_fixed = _c(gvol)
    Dn.fix_volume(cell, gres, _c(lvol), gvol_fixed, _c(sphi), _c(lphi), *ws)
    b = cp.asarray(np.full(g, np.nan))
    Dn.initialize_solver(1000.0, 1.0 / 300, gres, cell, gm, gvol_fixed, _c(lphi), *ws, b)
    rng = np.random.default_rng(62)
    pv = rng.standard_normal(g)
    q = cp.asarray(np.full(g, np.nan))
    Dn.matvecmul(gres, _c(pv), q, *ws, _c(lphi))
    disp = [cp.asarray(np.full(s, np.nan)) for s in _mac_shapes(g)]
    Dn.compute_displacement(gres, 1.0 / 300, cell, *disp, _c(pv), _c(lphi))
    dfield = [rng.standard_normal(s) * 1e-3 for s in _mac_shapes(g)]
    pmoved = _c(px)
    bias = ([0, 0.5, 0.5], [0.5, 0, 0.5], [0.5, 0.5, 0])
    for a in range(3):
        Dn.apply_displacement(pmoved, _c(dfield[a]), bound_min, cell, cp.array(bias[a], dtype=cp.float64), a)
    _save("density3d_kernels_7x8x6", gres=np.array(g), dx=dx, sphi=sphi, lvol=lvol, lphi=lphi, px=px, pm=pm, pvol=pvol,
          wx=ws[0], wy=ws[1], wz=ws[2], gm=gm, gvol=gvol, gvol_fixed=gvol_fixed, b=b, pv=pv, q=q,
          dispx=disp[0], dispy=disp[1], dispz=disp[2], dfx=dfield[0], dfy=dfield[1], dfz=dfield[2], pmoved=pmoved)


def case_density3d_solve():
    from solver import DensityCGSolver3D as Dn
    from solver.CGSolverBuffer import CGSolverBuffer
    g = (8, 10, 8)
    dx = 0.0125
    sphi, lvol, lphi, vel, px, pm, pvol = _density_inputs(g, dx, 63)
    gres = cp.array(g, dtype=cp.int64)
    buf = CGSolverBuffer(gres)
    s = Dn.DensityCGSolver3D(buf, gres, cp.array([0.0, 0.0, 0.0]), cp.array([n * dx for n in g], dtype=cp.float32))
    pout = _c(px)
    box, restore = _count_calls(Dn, "matvecmul")
    t = time.time()
    s.solve(1000.0, 1.0 / 300, pout, _c(pm), pvol, *[_c(v) for v in vel], _c(sphi), None, _c(lphi), _c(lvol), tol=1e-3)
    restore()
    iters = box["n"] - 1
    print(f"  density3d solve: {iters} iterations, delta={s.delta:.3e}, {time.time()-t:.0f}s")
    _save("density3d_solve_8x10x8", gres=np.array(g), dx=dx, bound_size=np.asarray([n * dx for n in g], dtype=np.float32), sphi=sphi, lvol=lvol,
          lphi=lphi, px=px, pm=pm, pvol=pvol, px_new=pout, x=s.x, wx=s.wx, wy=s.wy, wz=s.wz, m=s.m, vol=s.vol, b=buf.b,
          dispx=s.dx, dispy=s.dy, dispz=s.dz, iterations=iters, delta=s.delta, tol=1e-3)


CASES = {
    "density3d_kernels": case_density3d_kernels,
    "density3d_solve": case_density3d_solve,
    "visc3d_kernels": case_visc3d_kernels,
    "solidfrac3d": case_solidfrac3d,
    "solidfrac2d": case_solidfrac2d,
    "press3d_kernels": case_press3d_kernels,
    "visc2d_kernels": case_visc2d_kernels,
    "press2d_kernels": case_press2d_kernels,
    "visc2d_solve": case_visc2d_solve,
    "press2d_solve": case_press2d_solve,
    "press3d_solve": case_press3d_solve,
    "visc3d_solve": case_visc3d_solve,
    "visc3d_solve_stiff": case_visc3d_solve_stiff,
}

if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit(f"reference tree not found at {REF}: fixtures can only be regenerated in the build container")
    names = sys.argv[1:] or list(CASES)
    for n in names:
        print(f"[{n}]")
        t0 = time.time()
        CASES[n]()
        print(f"  done in {time.time()-t0:.1f}s")
