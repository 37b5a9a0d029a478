"""GPU: the active-set machinery of the 3-D viscosity CG (segment lists walked by K1/K2/K3).

The CG kernels only visit lattice segments that carry a computed row.  "fluid" visits every row the reference
computes (ViscosityCGSolver3D.py:251-258), "nonzero" (default) drops rows whose seven coefficients are all zero
(all-zero row and column: b = q = r = d = 0 there, x untouched).  Both must reproduce the reference / oracle.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2

pytestmark = pytest.mark.gpu


def _expected_rows(sphi, lvol, g, nonzero):
    """number of computed rows, restated in NumPy from the kernel rules (fluid face, interior, optional non-zero test)"""
    nzf = lvol != 0
    nb = nzf.copy()
    for ax in range(3):
        hi = np.zeros_like(nzf)
        lo = np.zeros_like(nzf)
        sl_a = [slice(None)] * 3
        sl_b = [slice(None)] * 3
        sl_a[ax], sl_b[ax] = slice(1, None), slice(None, -1)
        hi[tuple(sl_b)] = nzf[tuple(sl_a)]
        lo[tuple(sl_a)] = nzf[tuple(sl_b)]
        nb |= hi | lo
    total = 0
    for par in ((0, 1, 1), (1, 0, 1), (1, 1, 0)):
        sl = tuple(slice(p, None, 2) for p in par)
        fl = sphi[sl] >= 0
        interior = np.zeros(fl.shape, bool)
        interior[1:-1, 1:-1, 1:-1] = True
        rows = fl & interior
        if nonzero:
            rows &= nb[sl]
        total += int(rows.sum())
    return total


@pytest.mark.parametrize("N", [20, 33])
def test_active_rows_and_segments_match_numpy(N):
    import scenes
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    sc = scenes.buckling(N, device="cuda", mu=10.0, gres=(N, N + 3, N - 2))
    sphi, lvol = sc["sphi"].cpu().numpy(), sc["lvol"].cpu().numpy()
    seen = {}
    for mode in ("fluid", "nonzero"):
        s = ViscosityCGSolver3D(sc["gres"], sc["bound_size"], active_set=mode)
        v = [sc[k].clone() for k in ("vx", "vy", "vz")]
        s.solve(sc["dt"], 10.0, sc["rho"], *v, sc["sphi"], None, None, sc["lvol"])
        segs, total, rows = s.active_info()
        assert rows == _expected_rows(sphi, lvol, sc["gres"], mode == "nonzero")
        assert 0 < segs <= total and segs * 32 >= rows / 3
        seen[mode] = (segs, rows, s.iterations, [a.clone() for a in v])
    assert seen["nonzero"][0] < seen["fluid"][0] and seen["nonzero"][1] < seen["fluid"][1]
    # identical problem: same iteration count up to reduction-order rounding, same velocities
    assert abs(seen["nonzero"][2] - seen["fluid"][2]) <= 1
    for a, b in zip(seen["nonzero"][3], seen["fluid"][3]):
        assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("mode", ["fluid", "nonzero"])
@pytest.mark.parametrize("tag", ["visc3d_solve_8x10x8", "visc3d_solve_stiff_6x8x6"])
def test_both_modes_vs_reference_fixture(tag, mode):
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    f = load_golden(tag)
    s = ViscosityCGSolver3D(f["gres"], f["bound_size"], active_set=mode)
    v = [torch.as_tensor(f[k]).cuda() for k in ("vx", "vy", "vz")]
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    s.solve(float(f["dt"]), float(f["mu"]), float(f["rho"]), *v, dev(f["sphi"]), None, dev(f["lphi"]), dev(f["lvol"]), tol=float(f["tol"]))
    it_ref = int(f["iterations"])
    assert abs(s.iterations - it_ref) <= max(1, round(0.02 * it_ref)), (s.iterations, it_ref)
    for a, n in zip(v, "xyz"):
        assert rel_l2(a.cpu().numpy(), f[f"v{n}_new"]) < 1e-4
    for nm in ("r", "d", "q", "x", "b"):
        for c in "xyz":
            assert bool(torch.isfinite(getattr(s, f"{nm}_{c}")).all())


def test_no_liquid_at_all():
    """lvol == 0 everywhere: the non-zero active set is empty, delta0 == 0 and the solve returns at once (the reference
    also skips its loop: `if not self.delta < tol ** 2`)."""
    import scenes
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    sc = scenes.buckling(16, device="cuda")
    s = ViscosityCGSolver3D(sc["gres"], sc["bound_size"])
    v = [sc[k].clone() for k in ("vx", "vy", "vz")]
    s.solve(sc["dt"], 1.0, sc["rho"], *v, sc["sphi"], None, None, torch.zeros_like(sc["lvol"]))
    assert s.iterations == 0 and s.delta == 0.0
    assert s.active_info()[0] == 0
    for a, k in zip(v, ("vx", "vy", "vz")):
        assert torch.equal(a, sc[k])


def test_active_set_changes_between_solves():
    """The list is rebuilt by every solve: a second solve on a different liquid region must not reuse stale segments
    or a stale captured graph."""
    import scenes
    from oracle import numpy_oracle as O
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    sc = scenes.buckling(24, device="cuda", mu=10.0)
    s = ViscosityCGSolver3D(sc["gres"], sc["bound_size"])
    lv2 = sc["lvol"].clone()
    lv2[:, : lv2.shape[1] // 3] = 0                      # drain the pool
    for lvol in (sc["lvol"], lv2, sc["lvol"]):
        v = [sc[k].clone() for k in ("vx", "vy", "vz")]
        s.solve(sc["dt"], 10.0, sc["rho"], *v, sc["sphi"], None, None, lvol)
        ref = O.ViscosityCGSolver3D(sc["gres"], sc["bound_size"])
        rv = [sc[k].cpu().numpy().copy() for k in ("vx", "vy", "vz")]
        ref.solve(sc["dt"], 10.0, sc["rho"], *rv, sc["sphi"].cpu().numpy(), None, None, lvol.cpu().numpy())
        assert abs(s.iterations - ref.trace.iterations) <= max(1, round(0.02 * ref.trace.iterations)), (s.iterations, ref.trace.iterations)
        for a, b in zip(v, rv):
            assert rel_l2(a.cpu().numpy(), b) < 1e-4


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_persistent_kernel_matches_three_kernel_path(dtype):
    """cg_mode="persistent" (one cooperative launch, grid barriers), cg_mode="kernels" (K1/K2/K3 from a CUDA graph) and the
    single-reduction forms ("kernels_sr", "persistent_sr"; "auto" resolves to one of them) run the same iteration: same iteration count up to reduction-order rounding, same velocities, on both active-set
    modes, and on a grid large enough for several segments per warp."""
    import scenes
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    sc = scenes.buckling(48, device="cuda", mu=100.0)
    out = {}
    for aset in ("nonzero", "fluid"):
        for mode in ("kernels", "persistent", "kernels_sr", "persistent_sr", "auto"):
            s = ViscosityCGSolver3D(sc["gres"], sc["bound_size"], dtype=dtype, active_set=aset, cg_mode=mode)
            v = [sc[k].clone() for k in ("vx", "vy", "vz")]
            s.solve(sc["dt"], 100.0, sc["rho"], *v, sc["sphi"], None, None, sc["lvol"])
            s.solve(sc["dt"], 100.0, sc["rho"], *v, sc["sphi"], None, None, sc["lvol"])     # second step on the same object
            out[(aset, mode)] = (s.iterations, v, s.delta)
            assert s.delta < 1e-6
    ref_it, ref_v, _ = out[("fluid", "kernels")]
    for key, (it, v, _) in out.items():
        assert abs(it - ref_it) <= max(1, round(0.02 * ref_it)), (key, it, ref_it)
        for a, b in zip(v, ref_v):
            assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) < 1e-5, key


def test_fixed_window_persistent_counts_iterations():
    import scenes
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    sc = scenes.buckling(32, device="cuda", mu=100.0)
    d_after, x_after, deltas = {}, {}, {}
    for mode in ("kernels", "persistent", "kernels_sr", "persistent_sr", "auto"):
        s = ViscosityCGSolver3D(sc["gres"], sc["bound_size"], cg_mode=mode)
        s.max_iter = 150
        v = [sc[k].clone() for k in ("vx", "vy", "vz")]
        with pytest.raises(ValueError, match="Failed to converge!"):
            s.solve(sc["dt"], 100.0, sc["rho"], *v, sc["sphi"], None, None, sc["lvol"], tol=0.0)
        assert s.iterations == 150
        d_after[mode] = [a.clone() for a in (s.d_x, s.d_y, s.d_z)]
        x_after[mode] = [a.clone() for a in (s.x_x, s.x_y, s.x_z)]
        deltas[mode] = s.delta
    # the live search direction ends up in the primary buffer whichever way the launches were cut (150 = 64 + 64 + 22);
    # in the single-reduction forms d holds p_k, the direction the last update used, while the reference's loop has already
    # formed d_{k+1} = r + beta d_k — compare those through x instead
    for mode in ("persistent",):
        for a, b in zip(d_after[mode], d_after["kernels"]):
            assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) < 1e-3, mode
    for mode in x_after:
        assert abs(deltas[mode] - deltas["kernels"]) <= 1e-3 * deltas["kernels"], (mode, deltas)
        for a, b in zip(x_after[mode], x_after["kernels"]):
            assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) < 1e-6, mode


@pytest.mark.parametrize("cap", ["0", "7", None])
def test_extrapolation_worklists_and_overflow_fallback(cap, monkeypatch):
    """Sweeps >= 2 of the extrapolation are driven by the list of faces the previous sweep filled; a list that overflows
    (cap=7) or is disabled (cap=0) falls back to full passes.  All three must reproduce the reference bit for bit."""
    from solver import ViscosityCGSolver3D as V
    if cap is None:
        monkeypatch.delenv("FLUIDSOLVER_B200_EXTRAP_CAP", raising=False)
    else:
        monkeypatch.setenv("FLUIDSOLVER_B200_EXTRAP_CAP", cap)
    V._engines.clear()
    f = load_golden("visc3d_kernels_6x7x8")
    g = tuple(int(n) for n in f["gres"])
    v = [torch.as_tensor(f[k]).cuda() for k in ("vx", "vy", "vz")]
    V.extrapolate(g, 3, *v, torch.as_tensor(f["sphi"]).cuda(), dtype=torch.float64)
    for o, n in zip(v, "xyz"):
        assert np.array_equal(o.cpu().numpy(), f["e" + n])
    # a larger random case: all variants agree with each other (compared against the cap=None run through a module cache)
    gen = torch.Generator(device="cuda").manual_seed(11)
    g2 = (19, 14, 23)
    sphi = torch.randn(tuple(2 * n + 1 for n in g2), dtype=torch.float64, device="cuda", generator=gen) - 0.8
    vel = [torch.randn(sh, dtype=torch.float64, device="cuda", generator=gen) for sh in ((20, 14, 23), (19, 15, 23), (19, 14, 24))]
    for sweeps in (1, 2, 5):
        out = [a.clone() for a in vel]
        V.extrapolate(g2, sweeps, *out, sphi, dtype=torch.float64)
        key = ("extrap", sweeps)
        if key not in _EXTRAP_CACHE:                        # the NumPy oracle (Jacobi with full copies), once per sweep count
            from oracle import numpy_oracle as O
            ref = [a.cpu().numpy().copy() for a in vel]
            O.visc3d_extrapolate(g2, sweeps, *ref, sphi.cpu().numpy())
            _EXTRAP_CACHE[key] = ref
        for a, b in zip(out, _EXTRAP_CACHE[key]):
            assert np.array_equal(a.cpu().numpy(), b)
    V._engines.clear()


_EXTRAP_CACHE = {}


@pytest.mark.parametrize("N,mu", [(128, 100.0), (256, 100.0)])
def test_true_residual_at_full_size(N, mu):
    """Size-independent end-to-end property at BASELINE.json's full size: after solve() on the benchmark scene, the TRUE
    residual b - A x — recomputed with the dense masked operator kernel (fs_visc3d_apply: one thread per lattice point, no
    lists, no persistent kernel) — matches the recursively updated residual the CG stopped on, and x is untouched outside
    the active set.  This ties the active-set / persistent path to the dense operator on 50 million faces."""
    import ctypes
    import scenes
    from solver import _native as N_
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    sc = scenes.buckling(N, device="cuda", mu=mu)
    s = ViscosityCGSolver3D(sc["gres"], sc["bound_size"])
    v = [sc[k].clone() for k in ("vx", "vy", "vz")]
    s.solve(sc["dt"], mu, sc["rho"], *v, sc["sphi"], None, None, sc["lvol"], tol=1e-3)
    assert s.delta < 1e-6 and s.iterations > 50
    segs, total, rows = s.active_info()
    assert 0 < segs < total
    lib, e = N_.load(), s._e
    scale = sc["dt"] / s.cell_vol / sc["rho"]
    b = [a.clone() for a in (s.b_x, s.b_y, s.b_z)]
    r = [a.clone() for a in (s.r_x, s.r_y, s.r_z)]
    x = [a.clone() for a in (s.x_x, s.x_y, s.x_z)]
    N_.check(lib.fs_visc3d_apply(e.h, scale, mu, N_.VEC_X, N_.VEC_Q, 0), "apply")       # q = A x, dense masked kernel
    torch.cuda.synchronize()
    true_rr = 0.0
    rec_rr = 0.0
    bnorm = 0.0
    for bb, rr, qq in zip(b, r, (s.q_x, s.q_y, s.q_z)):
        t = bb - qq
        true_rr += float((t * t).sum())
        rec_rr += float((rr * rr).sum())
        bnorm += float((bb * bb).sum())
        assert float((t - rr).abs().max()) < 1e-7 * max(1.0, float(bb.abs().max()))
    assert abs(rec_rr - s.delta) <= 1e-9 * max(s.delta, 1e-300)
    assert true_rr < 4.0 * s.delta + 1e-10 * bnorm
    # rows outside the active set: the solution equals the (extrapolated) start value, i.e. the caller's arrays are unchanged
    changed = sum(int((a != b0).sum()) for a, b0 in zip(v, (sc["vx"], sc["vy"], sc["vz"])))
    assert 0 < changed <= rows
    del x
