"""GPU parity of the notebook's grid-side functions (notebook_kernels.py -> fs_grid_* C ABI -> sm_100a kernels) against the
fixtures the notebook's own cells produced under Numba's simulator, and against the NumPy oracle on a larger seeded case.
Index sets exact; scattered fp32 quantities within 1e-5 of the field maximum (floating-point atomics in the reference)."""
import types

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

CASES = ["nb_kernels_6x7x8", "nb_kernels_9x8x7"]


def NS(**kw):
    return types.SimpleNamespace(**kw)


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


def close(a, b, tol=1e-5):
    a = a.detach().cpu().numpy().astype(np.float64) if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    scale = max(np.max(np.abs(b)), 1e-300)
    err = np.max(np.abs(a - b)) / scale
    assert err < tol, err


def build(f, m_key="p2g_m", v_key="p2g_v", zero=False):
    g = [int(n) for n in f["gres"]]
    comps = {}
    for a in "xyz":
        m = f[m_key + a]
        v = f[v_key + a]
        comps[a] = NS(m=dev(np.zeros_like(m) if zero else m), v=dev(np.zeros_like(v) if zero else v), dv=torch.full(m.shape, float("nan"), dtype=torch.float32, device="cuda"))
    grid = NS(resolution=np.asarray(g), bound_min=f["bound_min"], bound_size=f["bound_size"], cell_size=f["cell_size"], **comps)
    p = NS(num_particles=f["px"].shape[0], x=dev(f["px"]), m=dev(f["pm"]), v=dev(f["pv"]), cx=dev(f["cx"]), cy=dev(f["cy"]), cz=dev(f["cz"]), vol=float(f["pvol"]))
    return g, grid, p


@pytest.mark.parametrize("tag", CASES)
def test_levelset_and_volume_vs_reference(tag):
    import notebook_kernels as K
    f = load_golden(tag)
    g, grid, p = build(f)
    ls = NS(resolution=np.asarray(g), bound_min=f["bound_min"], cell_size=f["bound_size"] / f["gres"], phi=torch.zeros(tuple(g), dtype=torch.float64, device="cuda"))
    K.compute_fluid_levelset(p, ls, float(f["dx"]))
    phi = ls.phi.cpu().numpy()
    assert np.array_equal(phi == 3 * float(f["dx"]), f["lphi"] == 3 * float(f["dx"]))
    close(phi, f["lphi"], 1e-6)
    res = [2 * n + 1 for n in g]
    fv = NS(resolution=np.asarray(res), bound_min=f["bound_min"], cell_size=f["bound_size"] / (2 * f["gres"]), vol=torch.full(tuple(res), 7.0, dtype=torch.float64, device="cuda"))
    K.compute_fluid_volume(p, fv, p.vol)
    vol = fv.vol.cpu().numpy()
    assert np.array_equal(vol > 0, f["lvol"] > 0)
    close(vol, f["lvol"], 1e-6)


@pytest.mark.parametrize("tag", CASES)
def test_time_loop_stages_vs_reference(tag):
    import notebook_kernels as K
    f = load_golden(tag)
    # P2G from zeroed grids
    g, grid, p = build(f, zero=True)
    K.p2g(p, grid)
    for a in "xyz":
        c = getattr(grid, a)
        assert np.array_equal(c.m.cpu().numpy() > 0, f[f"p2g_m{a}"] > 0)
        close(c.m, f[f"p2g_m{a}"])
        close(c.v, f[f"p2g_v{a}"])
    # extrapolation (2 sweeps) from the reference's P2G state
    g, grid, p = build(f)
    K.extrapolate(np.asarray(g), 2, grid.x.v, grid.y.v, grid.z.v, grid.x.m, grid.y.m, grid.z.m)
    for a in "xyz":
        v = getattr(grid, a).v.cpu().numpy()
        assert np.array_equal(v != f[f"p2g_v{a}"], f[f"ext_v{a}"] != f[f"p2g_v{a}"])
        close(v, f[f"ext_v{a}"], 1e-6)
    # boundary condition from the reference's extrapolated state
    g, grid, p = build(f, v_key="ext_v")
    solid = NS(phi=dev(f["sphi"]), v=dev(f["sv"]))
    K.apply_boundary_condition(grid, solid, float(f["dx"]))
    for a in "xyz":
        c = getattr(grid, a)
        dv = c.dv.cpu().numpy()
        assert not np.isnan(dv).any()                                    # every entry of dv is written, nothing outside it
        assert np.array_equal(dv != 0, f[f"bc_dv{a}"] != 0)
        close(dv, f[f"bc_dv{a}"])
        close(c.v, f[f"bc_v{a}"])
    # G2P from the reference's final grid state
    g, grid, p = build(f, v_key="bc_v")
    K.g2p(p, grid)
    close(p.v, f["g2p_pv"])
    for k in ("cx", "cy", "cz"):
        close(getattr(p, k), f["g2p_" + k])


def test_larger_case_vs_oracle_and_conservation():
    import notebook_kernels as K
    from oracle import numpy_oracle_nb as NB
    rng = np.random.default_rng(5)
    g = [40, 36, 44]
    dx = 0.02
    bmin = np.array([-0.4, 0.0, -0.44], dtype=np.float32)
    bsize = (np.asarray(g) * dx).astype(np.float32)
    cell = bsize / np.asarray(g)
    n = 60000
    px = bmin + (0.15 + 0.7 * rng.random((n, 3))) * bsize
    pm = np.full(n, 1000.0 * (dx / 2) ** 3)
    pv = rng.normal(0, 1, (n, 3))
    pc = [rng.normal(0, 2, (n, 3)) for _ in range(3)]
    sh = [tuple(m + (1 if i == a else 0) for i, m in enumerate(g)) for a in range(3)]
    comps = {a: NS(m=torch.zeros(s, dtype=torch.float32, device="cuda"), v=torch.zeros(s, dtype=torch.float32, device="cuda"), dv=torch.zeros(s, dtype=torch.float32, device="cuda"))
             for a, s in zip("xyz", sh)}
    grid = NS(resolution=np.asarray(g), bound_min=bmin, bound_size=bsize, cell_size=cell, **comps)
    p = NS(num_particles=n, x=dev(px), m=dev(pm), v=dev(pv), cx=dev(pc[0]), cy=dev(pc[1]), cz=dev(pc[2]), vol=(dx / 2) ** 3)
    K.p2g(p, grid)
    ms = [np.zeros(s, dtype=np.float32) for s in sh]
    vs = [np.zeros(s, dtype=np.float32) for s in sh]
    NB.p2g(g, bmin, cell, px, pm, pv, pc, ms, vs)
    for a, m, v in zip("xyz", ms, vs):
        c = getattr(grid, a)
        assert np.array_equal(c.m.cpu().numpy() > 0, m > 0)
        close(c.m, m, 2e-5)
        close(c.v, v, 2e-5)
        assert abs(float(c.m.double().sum()) - pm.sum()) < 1e-5 * pm.sum()          # partition of unity: mass is conserved
    K.extrapolate(np.asarray(g), 3, grid.x.v, grid.y.v, grid.z.v, grid.x.m, grid.y.m, grid.z.m)
    NB.extrapolate(3, vs, ms)
    for a, v in zip("xyz", vs):
        close(getattr(grid, a).v, v, 1e-6)
    ls = NS(resolution=np.asarray(g), bound_min=bmin, cell_size=cell, phi=torch.zeros(tuple(g), dtype=torch.float64, device="cuda"))
    K.compute_fluid_levelset(p, ls, dx)
    close(ls.phi, NB.fluid_levelset(g, bmin, cell, px, dx), 1e-6)
    res = [2 * m + 1 for m in g]
    fv = NS(resolution=np.asarray(res), bound_min=bmin, cell_size=bsize / (2 * np.asarray(g)), vol=torch.zeros(tuple(res), dtype=torch.float64, device="cuda"))
    K.compute_fluid_volume(p, fv, p.vol)
    close(fv.vol, NB.fluid_volume(res, bmin, bsize / (2 * np.asarray(g)), px, p.vol), 1e-9)
    pvo, pco = NB.g2p(g, bmin, cell, px, [getattr(grid, a).v.cpu().numpy() for a in "xyz"])
    K.g2p(p, grid)
    close(p.v, pvo, 1e-5)
    for k, c in zip(("cx", "cy", "cz"), pco):
        close(getattr(p, k), c, 1e-5)


def test_host_numpy_arrays_are_updated_in_place():
    """the notebook passes device arrays; host arrays (staged copies) must be written back like every other entry point"""
    import notebook_kernels as K
    f = load_golden(CASES[0])
    g = [int(n) for n in f["gres"]]
    vs = [f[f"p2g_v{a}"].copy() for a in "xyz"]
    ms = [f[f"p2g_m{a}"].copy() for a in "xyz"]
    K.extrapolate(np.asarray(g), 2, *vs, *ms)
    for v, a in zip(vs, "xyz"):
        close(v, f[f"ext_v{a}"], 1e-6)
