#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — golden fixtures for the notebook's grid-side kernels.

The kernels live in code cells of ``/root/reference/3D_viscous_fluid_sim.ipynb`` (P2G :cell 2, G2P :cell 3, fluid level
set :cell 4, boundary condition :cell 5, fluid volume :cell 6, extrapolation :cell 7), not in an importable module.  This
script reads those cells' source from the notebook, executes them UNMODIFIED under Numba's CUDA simulator with the
NumPy-backed ``cupy`` shim, drives them on a small seeded scene and writes inputs + outputs to ``tests/golden/nb_*.npz``.

The host launchers of the notebook are used as they are, with one exception: ``apply_boundary_condition`` launches (8,8,8)
blocks over arrays whose shapes are not multiples of 8, and its kernels write ``dv[x,y,z] = 0`` BEFORE their bounds test —
out of bounds for every overhanging thread (undefined behaviour on a GPU, IndexError under the simulator).  Here the three
kernels are launched with an exact-fit grid instead, which pins the in-bounds behaviour; the launcher's three
``g.*.v += g.*.dv`` updates are applied afterwards as in the notebook.

Never run on the GPU box; only the committed ``.npz`` files travel.   Usage: python oracle/refsim/run_notebook_kernels.py
"""
import json
import math
import os
import sys
import types

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("FLUID_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, HERE)      # the cupy shim

import numpy as np  # noqa: E402
import cupy as cp  # noqa: E402  (the shim)
from numba import cuda  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
CELLS = {"p2g": 2, "g2p": 3, "levelset": 4, "boundary": 5, "volume": 6, "extrapolate": 7}


def notebook_namespace():
    """exec the six kernel cells of the notebook, unmodified, into one namespace"""
    with open(os.path.join(REF, "3D_viscous_fluid_sim.ipynb")) as f:
        nb = json.load(f)
    ns = {"cp": cp, "cuda": cuda, "math": math, "__name__": "notebook_cells"}
    for name, idx in CELLS.items():
        src = "".join(nb["cells"][idx]["source"])
        assert nb["cells"][idx]["cell_type"] == "code"
        exec(compile(src, f"<ipynb cell {idx}: {name}>", "exec"), ns)
    return ns


def NS(**kw):
    return types.SimpleNamespace(**kw)


def _c(a, dtype=None):
    return cp.asarray(np.array(a, dtype=dtype, copy=True))


def make_case(g, dx, seed, n_particles):
    """grid / particle objects laid out like the notebook's (ipynb cell 10): fp64 particles, fp32 MAC grids,
    fp64 level-set and volume grids, float32 bound_min / bound_size, cell_size = bound_size / gres (fp64)."""
    rng = np.random.default_rng(seed)
    g = np.asarray(g, dtype=np.int64)
    bound_min = np.array([-0.3, 0.0, -0.25], dtype=np.float32)
    bound_size = (g * dx).astype(np.float32)
    cell = bound_size / g                                    # float32 / int64 -> float64, as in the notebook
    L = g * dx
    # liquid blob: particles in the lower middle of the box, a few right at / beyond the domain faces (index clamping)
    px = np.empty((n_particles, 3))
    lo = bound_min + np.array([1.3, 1.2, 1.4]) * dx
    hi = bound_min + L - np.array([1.3, 2.5, 1.4]) * dx
    px[:] = lo + rng.random((n_particles, 3)) * (hi - lo)
    px[:6] = bound_min + rng.random((6, 3)) * 0.4 * dx                        # inside the first cell
    px[6:10] = bound_min + L - rng.random((4, 3)) * 0.3 * dx                  # inside the last cell
    pdx = dx / 2
    rho = 1000.0
    p = NS(num_particles=n_particles, x=_c(px), m=_c(np.ones(n_particles) * rho * pdx ** 3), v=_c(rng.normal(0, 1.0, (n_particles, 3))),
           cx=_c(rng.normal(0, 2.0, (n_particles, 3))), cy=_c(rng.normal(0, 2.0, (n_particles, 3))), cz=_c(rng.normal(0, 2.0, (n_particles, 3))),
           vol=pdx ** 3)

    def comp(a):
        res = g + np.eye(3, dtype=np.int64)[a]
        bias = np.full(3, 0.5, dtype=np.float32)
        bias[a] = 0.0
        sh = tuple(int(n) for n in res)
        return NS(resolution=_c(res), bias=_c(bias), m=cp.zeros(sh, dtype=cp.float32), v=cp.zeros(sh, dtype=cp.float32), dv=cp.zeros(sh, dtype=cp.float32))

    grid = NS(resolution=_c(g), bound_size=_c(bound_size), bound_min=_c(bound_min), cell_size=_c(cell), x=comp(0), y=comp(1), z=comp(2))
    fine = 2 * g + 1
    fls = NS(resolution=_c(g), bound_size=_c(bound_size), bound_min=_c(bound_min), cell_size=_c(bound_size / g), phi=cp.zeros(tuple(int(n) for n in g)))
    fvol = NS(resolution=_c(fine), bound_size=_c(bound_size), bound_min=_c(bound_min), cell_size=_c(bound_size / (2 * g)),
              vol=cp.zeros(tuple(int(n) for n in fine)))
    # solid: box container inset 1.2 cells (phi > 0 inside), sampled on the fine grid; a moving-wall velocity field
    ax = [bound_min[k] + np.arange(fine[k]) * (dx / 2) for k in range(3)]
    P = np.meshgrid(*ax, indexing="ij")
    inside = np.minimum.reduce([np.minimum(P[k] - (bound_min[k] + 1.2 * dx), (bound_min[k] + L[k] - 1.2 * dx) - P[k]) for k in range(3)])
    sv = rng.normal(0, 0.2, tuple(int(n) for n in fine) + (3,))
    solid = NS(phi=_c(inside.astype(np.float64)), v=_c(sv))
    return p, grid, fls, fvol, solid, dict(gres=g, dx=dx, bound_min=bound_min, bound_size=bound_size, cell_size=cell)


def snapshot_grid(grid, tag, out):
    for a in "xyz":
        c = getattr(grid, a)
        out[f"{tag}_m{a}"] = np.array(c.m)
        out[f"{tag}_v{a}"] = np.array(c.v)


def main():
    ns = notebook_namespace()
    os.makedirs(GOLD, exist_ok=True)
    for name, g, dx, seed, npart in (("nb_kernels_6x7x8", (6, 7, 8), 0.1, 11, 160), ("nb_kernels_9x8x7", (9, 8, 7), 0.05, 12, 220)):
        p, grid, fls, fvol, solid, meta = make_case(g, dx, seed, npart)
        out = dict(meta)
        out.update(px=np.array(p.x), pm=np.array(p.m), pv=np.array(p.v), cx=np.array(p.cx), cy=np.array(p.cy), cz=np.array(p.cz), pvol=p.vol,
                   sphi=np.array(solid.phi), sv=np.array(solid.v))
        for a in "xyz":
            out[f"bias_{a}"] = np.array(getattr(grid, a).bias)
        # ---- fluid level set / fluid volume (ipynb cells 4, 6) ----
        ns["compute_fluid_levelset"](p, fls, dx)
        out["lphi"] = np.array(fls.phi)
        ns["compute_fluid_volume"](p, fvol, p.vol)
        out["lvol"] = np.array(fvol.vol)
        # ---- P2G (cell 2) ----
        ns["p2g"](p, grid)
        snapshot_grid(grid, "p2g", out)
        # ---- extrapolation with mass validity, 2 sweeps as in the time loop (cell 7, ipynb cell 13) ----
        ns["extrapolate"](grid.resolution, 2, grid.x.v, grid.y.v, grid.z.v, grid.x.m, grid.y.m, grid.z.m)
        snapshot_grid(grid, "ext", out)
        # ---- boundary condition (cell 5): kernels launched with an exact-fit grid (see the module docstring) ----
        for a, kern, margs in (("x", "boundary_condition_x", (grid.y.m, grid.z.m)), ("y", "boundary_condition_y", (grid.x.m, grid.z.m)),
                               ("z", "boundary_condition_z", (grid.x.m, grid.y.m))):
            c = getattr(grid, a)
            c.dv[:] = np.nan
            with np.errstate(all="ignore"):
                ns[kern][tuple(int(n) for n in c.dv.shape), (1, 1, 1)](grid.x.v, grid.y.v, grid.z.v, *margs, solid.phi, solid.v, dx, c.dv)
            out[f"bc_dv{a}"] = np.array(c.dv)
        for a in "xyz":
            c = getattr(grid, a)
            c.v += c.dv
        snapshot_grid(grid, "bc", out)
        # ---- G2P (cell 3) ----
        ns["g2p"](p, grid)
        out.update(g2p_pv=np.array(p.v), g2p_cx=np.array(p.cx), g2p_cy=np.array(p.cy), g2p_cz=np.array(p.cz))
        path = os.path.join(GOLD, name + ".npz")
        np.savez_compressed(path, **{k: np.asarray(v) for k, v in out.items()})
        print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)", flush=True)


if __name__ == "__main__":
    main()
