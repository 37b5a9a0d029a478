"""Synthetic scenes for the benchmark configurations of BASELINE.json (SURVEY.md §8 d).

Pure torch (runs on CPU for tests and on the GPU for the large benches), deterministic, and
partition-independent: the velocity noise is a counter-based hash of the GLOBAL face index, so a
z-slab/x-slab decomposition over G ranks sees exactly the same field as one rank.

Every generator returns a dict with exactly the argument set of the reference's ``solve()`` calls
(the ``ml_data`` schema of 3D_viscous_fluid_sim.ipynb:4614-4630): ``vx,vy,vz`` (fp32 MAC faces),
``sphi`` / ``lvol`` (fp64, (2g+1) fine grid), ``lphi`` (fp64, cells), plus ``gres, bound_size, dx, dt,
rho, mu``.  ``sv`` (solid velocity, fine grid x D) is only materialised on request: the viscosity
solve never reads it.

The solid geometry of ``buckling`` is the notebook's five rigid bodies (ipynb:682-689) with the
box signed-distance of ``solver/sdf3D.py:86-109`` restated here in torch (p' = R^T (p - c);
d = |p'| - size/2; sd = |max(d,0)| + min(max_i d_i, 0); negated for the flipped container).
"""
import math

import torch

_M32 = 0xFFFFFFFF


def _mix32(h):
    h = h & _M32
    h = (h ^ (h >> 16)) * 0x7FEB352D & _M32
    h = (h ^ (h >> 15)) * 0x846CA68B & _M32
    return h ^ (h >> 16)


def hash_normal(idx, seed):
    """Standard-normal noise as a pure function of an int64 index tensor (Box–Muller on two hashes)."""
    h1 = _mix32(idx * 2 + 0x9E3779B1 * (seed + 1))
    h2 = _mix32(idx * 2 + 1 + 0x85EBCA77 * (seed + 1))
    u1 = (h1.to(torch.float64) + 0.5) / 4294967296.0
    u2 = (h2.to(torch.float64) + 0.5) / 4294967296.0
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * math.pi * u2)


def _rot(axis, deg):
    a = torch.tensor(axis, dtype=torch.float64)
    a = a / a.norm()
    t = math.radians(deg)
    K = torch.tensor([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]], dtype=torch.float64)
    return torch.eye(3, dtype=torch.float64) + math.sin(t) * K + (1 - math.cos(t)) * (K @ K)


def _box_sd(P, size, center, R, flip):
    """P: (...,3) fp64 positions.  Box SDF (sdf3D.py:86-109)."""
    c = torch.tensor(center, dtype=torch.float64, device=P.device)
    half = torch.tensor(size, dtype=torch.float64, device=P.device) * 0.5
    q = (P - c) @ R.to(P.device)          # rows: R^T (p - c)
    d = q.abs() - half
    sd = d.clamp(min=0).norm(dim=-1) + d.max(dim=-1).values.clamp(max=0)
    return -sd if flip else sd


def _mac_shapes(g):
    return [tuple(n + (1 if i == a else 0) for i, n in enumerate(g)) for a in range(len(g))]


def mac_velocities(g, base, sigma, seed, device, x0=0, gx_total=None):
    """fp32 MAC velocity arrays: base[a] + sigma * noise(global face index).

    ``x0`` / ``gx_total``: this rank's first cell along axis 0 and the global extent, for slab-partitioned
    generation (faces are indexed in the GLOBAL array so every partition sees the same numbers).
    """
    d = len(g)
    gt = list(g)
    if gx_total is not None:
        gt[0] = gx_total
    out = []
    for a, sh in enumerate(_mac_shapes(g)):
        gsh = tuple(n + (1 if i == a else 0) for i, n in enumerate(gt))
        idx = torch.arange(sh[0], device=device, dtype=torch.int64) + x0
        for k in range(1, d):
            idx = idx.unsqueeze(-1) * gsh[k] + torch.arange(sh[k], device=device, dtype=torch.int64)
        v = base[a] + sigma * hash_normal(idx, seed * 8 + a)
        out.append(v.to(torch.float32).contiguous())
    return out


def _fine_chunks(nf0, chunk):
    i = 0
    while i < nf0:
        j = min(nf0, i + chunk)
        yield i, j
        i = j


def _assemble(g, dx, bound_min, device, sphi_fn, liq_fn, chunk=16, x0=0):
    """Evaluate sphi / liquid SDF on the fine grid in x-chunks; returns sphi, lvol, lphi."""
    d = len(g)
    fine = tuple(2 * n + 1 for n in g)
    h = dx / 2
    sphi = torch.empty(fine, dtype=torch.float64, device=device)
    lvol = torch.empty(fine, dtype=torch.float64, device=device)
    lphi = torch.empty(tuple(g), dtype=torch.float64, device=device)
    axes = [bound_min[k] + torch.arange(fine[k], device=device, dtype=torch.float64) * h for k in range(d)]
    axes[0] = axes[0] + 2 * x0 * h
    for i, j in _fine_chunks(fine[0], chunk):
        grids = torch.meshgrid(axes[0][i:j], *axes[1:], indexing="ij")
        P = torch.stack(grids, dim=-1)
        s = sphi_fn(P)
        l = liq_fn(P)
        sphi[i:j] = s
        lvol[i:j] = (0.5 - l / h).clamp(0, 1) * (h ** d) * (s > 0)
        # cell centres are the all-odd fine nodes
        odd = [k for k in range(i, j) if k % 2 == 1]
        if odd:
            sel = l[[k - i for k in odd]]
            for ax in range(1, d):
                sel = sel.index_select(ax, torch.arange(1, fine[ax], 2, device=device))
            lphi[[(k - 1) // 2 for k in odd]] = sel
        del P, grids, s, l
    return sphi, lvol, lphi


def buckling(N, device="cpu", mu=1.0, seed=1234, dx=0.0125, with_sv=False, gres=None):
    """The notebook's buckling scene scaled to an N^3 grid (SURVEY §8 d, configs 1, 3, 4)."""
    g = tuple(gres) if gres is not None else (N, N, N)
    L = g[1] * dx                     # the notebook's domain height is 1.0 -> scale by L
    bound_size = tuple(n * dx for n in g)
    bound_min = (-bound_size[0] / 2, 0.0, -bound_size[2] / 2)
    I = torch.eye(3, dtype=torch.float64)
    bodies = [
        ((0.5 * L, 0.8 * L, 0.5 * L), (0.0, 0.5 * L, 0.0), I, True),
        ((0.67 * L, 0.1 * L, 1.0 * L), (-0.34 * L, 0.7 * L, 0.0), _rot((0, 0, 1), -45), False),
        ((0.67 * L, 0.1 * L, 1.0 * L), (0.34 * L, 0.7 * L, 0.0), _rot((0, 0, 1), 45), False),
        ((1.0 * L, 0.1 * L, 0.7 * L), (0.0, 0.7 * L, -0.3 * L), _rot((1, 0, 0), 45), False),
        ((1.0 * L, 0.1 * L, 0.7 * L), (0.0, 0.7 * L, 0.3 * L), _rot((1, 0, 0), -45), False),
    ]

    def sphi_fn(P):
        sd = None
        for size, c, R, flip in bodies:
            b = _box_sd(P, size, c, R, flip)
            sd = b if sd is None else torch.minimum(sd, b)
        return sd

    def liq_fn(P):
        r = torch.sqrt(P[..., 0] ** 2 + P[..., 2] ** 2)
        col = torch.maximum(r - 2.5 * dx, P[..., 1] - 0.68 * L)
        pool = P[..., 1] - (0.1 * L + 4 * dx)
        return torch.minimum(col, pool)

    sphi, lvol, lphi = _assemble(g, dx, bound_min, device, sphi_fn, liq_fn)
    vx, vy, vz = mac_velocities(g, (0.0, -2.0, 0.0), 0.1, seed, device)
    out = dict(name=f"buckling-{g[0]}x{g[1]}x{g[2]}", gres=g, bound_size=bound_size, bound_min=bound_min, dx=dx,
               dt=1.0 / 300, rho=1000.0, mu=float(mu), sphi=sphi, lvol=lvol, lphi=lphi, vx=vx, vy=vy, vz=vz)
    if with_sv:
        out["sv"] = torch.zeros(tuple(2 * n + 1 for n in g) + (3,), dtype=torch.float64, device=device)
    return out


def viscous_column(g_local, device="cpu", mu=100.0, seed=3456, dx=0.0125, x0=0, gx_total=None):
    """Weak-scaling scene (config 5): liquid column of radius 0.35*L_z along x inside a container with
    2-cell walls; the slab axis is x (storage-slowest), rank r owns cells [x0, x0+g_local[0]).
    """
    g = tuple(g_local)
    gxt = gx_total if gx_total is not None else g[0]
    Ltot = (gxt * dx, g[1] * dx, g[2] * dx)
    bound_min = (0.0, 0.0, 0.0)
    wall = 2 * dx

    def sphi_fn(P):
        ins = None
        for k in range(3):
            t = torch.minimum(P[..., k] - wall, (Ltot[k] - wall) - P[..., k])
            ins = t if ins is None else torch.minimum(ins, t)
        return ins

    def liq_fn(P):
        r = torch.sqrt((P[..., 1] - Ltot[1] / 2) ** 2 + (P[..., 2] - Ltot[2] / 2) ** 2)
        return r - 0.35 * Ltot[2]

    sphi, lvol, lphi = _assemble(g, dx, bound_min, device, sphi_fn, liq_fn, x0=x0)
    vx, vy, vz = mac_velocities(g, (0.0, -2.0, 0.0), 0.1, seed, device, x0=x0, gx_total=gxt)
    return dict(name=f"viscous-column-{g[0]}x{g[1]}x{g[2]}", gres=g, bound_size=tuple(n * dx for n in g), bound_min=bound_min,
                dx=dx, dt=1.0 / 300, rho=1000.0, mu=float(mu), sphi=sphi, lvol=lvol, lphi=lphi, vx=vx, vy=vy, vz=vz,
                x0=x0, gx_total=gxt)


def box2d(W, H=None, device="cpu", mu=1.0, seed=2345, with_sv=True):
    """2-D config 2: unit box, container inset 0.05 L, disc obstacle r=0.1 L at (0.5,0.3) L, liquid y<0.6 L."""
    H = W if H is None else H
    g = (W, H)
    dx = 1.0 / W
    Lx, Ly = W * dx, H * dx

    def sphi_fn(P):
        box = torch.minimum(torch.minimum(P[..., 0] - 0.05 * Lx, 0.95 * Lx - P[..., 0]),
                            torch.minimum(P[..., 1] - 0.05 * Ly, 0.95 * Ly - P[..., 1]))
        disc = torch.sqrt((P[..., 0] - 0.5 * Lx) ** 2 + (P[..., 1] - 0.3 * Ly) ** 2) - 0.1 * Lx
        return torch.minimum(box, disc)

    def liq_fn(P):
        return P[..., 1] - 0.6 * Ly

    sphi, lvol, lphi = _assemble(g, dx, (0.0, 0.0), device, sphi_fn, liq_fn, chunk=256)
    vx, vy = mac_velocities(g, (0.0, -2.0), 0.1, seed, device)
    out = dict(name=f"box2d-{W}x{H}", gres=g, bound_size=(Lx, Ly), dx=dx, dt=1.0 / 300, rho=1000.0, mu=float(mu),
               sphi=sphi, lvol=lvol, lphi=lphi, vx=vx, vy=vy)
    if with_sv:
        out["sv"] = torch.zeros((2 * W + 1, 2 * H + 1, 2), dtype=torch.float64, device=device)
    return out
