"""GPU: the implementation switches behind fs_set_option select between result-equivalent kernels — the persistent
single-reduction CG through global memory or as the shared-memory resident kernel; the stand-alone K1s
with interleaved or blocked warp -> segment mapping, or as the shared-memory tiled kernel for dense lattices.  Every form
must give the iterates of the others and the reference's iteration counts (ViscosityCGSolver3D.py:588-612)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2

pytestmark = pytest.mark.gpu

FORMS = [0, 2]             # resident_form: persistent CG through global memory / shared-memory resident kernel


@pytest.fixture(autouse=True)
def _restore_options():
    from solver import _native as N
    yield
    for k in ("resident_form", "k1_block", "k1_tile", "sparse_setup"):
        N.set_option(k, -1)


def _solve(sc, mu, max_iter=None, tol=1e-3, cg_mode="persistent_sr", active_set="nonzero"):
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    s = ViscosityCGSolver3D(sc["gres"], sc["bound_size"], cg_mode=cg_mode, active_set=active_set)
    if max_iter is not None:
        s.max_iter = max_iter
    v = [sc[k].clone() for k in ("vx", "vy", "vz")]
    try:
        s.solve(sc["dt"], mu, sc["rho"], *v, sc["sphi"], None, None, sc["lvol"], tol=tol)
    except ValueError:
        assert max_iter is not None
    x = [a.double().cpu().numpy().copy() for a in (s.x_x, s.x_y, s.x_z)]
    return s.iterations, float(s.delta), x, [a.cpu().numpy() for a in v]


@pytest.mark.parametrize("grid", [None, "3"])
def test_forms_agree_on_fixed_window(grid, monkeypatch):
    """100 fixed iterations on a 40x44x36 buckling scene: every form reproduces the global-memory form's delta and iterate.
    grid="3" forces three CTAs, so most of a CTA's run lies beyond the resident slots (the global-memory tail of the
    resident kernels is exercised together with the resident part)."""
    import scenes
    from solver import _native as N
    if grid:
        monkeypatch.setenv("FLUIDSOLVER_B200_PERSIST_GRID", grid)
    sc = scenes.buckling(40, device="cuda", mu=50.0, gres=(40, 44, 36))
    out = {}
    for form in FORMS:
        N.set_option("resident_form", form)
        out[form] = _solve(sc, 50.0, max_iter=100, tol=0.0)
    it0, d0, x0, _ = out[0]
    assert it0 == 100
    for key, (it, d, x, _) in out.items():
        assert it == it0, key
        # same recurrence, different reduction trees (block -> segment maps differ): rounding-level differences that the
        # iteration amplifies, same bars as test_active_set_gpu.test_fixed_window_persistent_counts_iterations
        assert abs(d - d0) <= 1e-3 * abs(d0), (key, d, d0)
        for a, b in zip(x, x0):
            assert rel_l2(a, b) < 1e-6, key


@pytest.mark.parametrize("form", FORMS)
@pytest.mark.parametrize("tag", ["visc3d_solve_8x10x8", "visc3d_solve_stiff_6x8x6"])
def test_forms_vs_reference_fixture(tag, form):
    from solver import _native as N
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    N.set_option("resident_form", form)
    f = load_golden(tag)
    s = ViscosityCGSolver3D(f["gres"], f["bound_size"], cg_mode="persistent_sr")
    v = [torch.as_tensor(f[k]).cuda() for k in ("vx", "vy", "vz")]
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    s.solve(float(f["dt"]), float(f["mu"]), float(f["rho"]), *v, dev(f["sphi"]), None, dev(f["lphi"]), dev(f["lvol"]), tol=float(f["tol"]))
    it_ref = int(f["iterations"])
    assert abs(s.iterations - it_ref) <= max(2, round(0.02 * it_ref)), (s.iterations, it_ref)
    for a, n in zip(v, "xyz"):
        assert rel_l2(a.cpu().numpy(), f[f"v{n}_new"]) < 1e-4


@pytest.mark.parametrize("form", [2])
def test_resident_forms_converged_and_repeated_solves(form):
    """Converged solves (several cooperative launches each) repeated on the same handle, against the NumPy oracle."""
    import scenes
    from oracle import numpy_oracle as O
    from solver import _native as N
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    N.set_option("resident_form", form)
    sc = scenes.buckling(24, device="cuda", mu=10.0)
    s = ViscosityCGSolver3D(sc["gres"], sc["bound_size"], cg_mode="persistent_sr")
    ref = O.ViscosityCGSolver3D(sc["gres"], sc["bound_size"])
    rv = [sc[k].cpu().numpy().copy() for k in ("vx", "vy", "vz")]
    ref.solve(sc["dt"], 10.0, sc["rho"], *rv, sc["sphi"].cpu().numpy(), None, None, sc["lvol"].cpu().numpy())
    for _ in range(3):
        v = [sc[k].clone() for k in ("vx", "vy", "vz")]
        s.solve(sc["dt"], 10.0, sc["rho"], *v, sc["sphi"], None, None, sc["lvol"])
        assert abs(s.iterations - ref.trace.iterations) <= max(1, round(0.02 * ref.trace.iterations)), (s.iterations, ref.trace.iterations)
        for a, b in zip(v, rv):
            assert rel_l2(a.cpu().numpy(), b) < 1e-4


def _solve_dtype(sc, mu, dtype, **kw):
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    s = ViscosityCGSolver3D(sc["gres"], sc["bound_size"], cg_mode="kernels_sr", active_set="fluid", dtype=dtype)
    s.max_iter = kw.get("max_iter", 60)
    v = [sc[k].clone() for k in ("vx", "vy", "vz")]
    try:
        s.solve(sc["dt"], mu, sc["rho"], *v, sc["sphi"], None, None, sc["lvol"], tol=0.0)
    except ValueError:
        pass
    return s.iterations, float(s.delta), [a.double().cpu().numpy().copy() for a in (s.x_x, s.x_y, s.x_z)]


@pytest.mark.parametrize("blk", [1, 3, 8])
def test_k1_blocked_mapping_agrees(blk):
    """Blocked warp -> segment mapping of the stand-alone K1s (1000 + n forces it on a list below the HBM-size threshold):
    every active segment is visited exactly once — same delta and iterate as the interleaved mapping, up to the rounding of
    a different reduction tree."""
    import scenes
    from solver import _native as N
    sc = scenes.buckling(40, device="cuda", mu=50.0, gres=(36, 40, 44))
    N.set_option("k1_block", 0)
    N.set_option("k1_tile", 0)
    it0, d0, x0 = _solve_dtype(sc, 50.0, torch.float64)
    N.set_option("k1_block", 1000 + blk)
    it, d, x = _solve_dtype(sc, 50.0, torch.float64)
    assert it0 == it == 60
    assert abs(d - d0) <= 1e-6 * abs(d0), (d, d0)
    for a, b in zip(x, x0):
        assert rel_l2(a, b) < 1e-9


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("rows,planes", [(0, 0), (2, 3), (3, 16), (5, 1)])
@pytest.mark.parametrize("gres", [(36, 40, 44), (21, 17, 30)])
def test_k1_tiled_kernel_agrees_with_list_kernel(gres, rows, planes, dtype):
    """The shared-memory tiled K1s (forced on: mode 2; row-block height and planes per work item varied, incl. blocks that
    overhang the last lattice row and single-plane items) computes the same w = A r and dot products as the list kernel:
    identical CG trajectory up to the rounding of the reduction tree."""
    import scenes
    from solver import _native as N
    sc = scenes.buckling(gres[0], device="cuda", mu=50.0, gres=gres)
    N.set_option("k1_block", 0)
    N.set_option("k1_tile", 0)
    it0, d0, x0 = _solve_dtype(sc, 50.0, dtype)
    N.set_option("k1_tile", 1000 * planes + 10 * rows + 2)
    it, d, x = _solve_dtype(sc, 50.0, dtype)
    assert it0 == it == 60
    tol_d, tol_x = (1e-6, 1e-6) if dtype == torch.float64 else (1e-2, 1e-4)     # (rounding of the reduction tree, amplified over 60 stiff iterations)
    assert abs(d - d0) <= tol_d * abs(d0), (d, d0)
    for a, b in zip(x, x0):
        assert rel_l2(a, b) < tol_x


@pytest.mark.parametrize("gres", [(24, 24, 24), (36, 40, 44)])
@pytest.mark.parametrize("cap", [None, "7"])
def test_sparse_setup_equals_dense_setup(gres, cap, monkeypatch):
    """fs_visc3d_solve with the sparse set-up (velocities loaded / extrapolated around the active set only) against the
    dense set-up: RHS, first residual, iteration count, delta and the written-back velocities are IDENTICAL BITS — on
    a sequence of solves with changing liquid regions on the same object (the lattice vector keeps stale values outside
    the region of each solve), with work lists large enough and overflowing (cap=7: sweeps 2 and 3 fall back to full
    passes)."""
    import scenes
    from solver import _native as N
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    if cap is not None:
        monkeypatch.setenv("FLUIDSOLVER_B200_EXTRAP_CAP", cap)
    sc = scenes.buckling(gres[0], device="cuda", mu=10.0, gres=gres)
    lv2 = sc["lvol"].clone()
    lv2[:, : lv2.shape[1] // 3] = 0                      # drain the pool
    lv3 = sc["lvol"].clone()
    lv3[: lv3.shape[0] // 2] = 0                         # keep the liquid of one half only
    N.set_option("sparse_setup", 1)
    sparse = ViscosityCGSolver3D(sc["gres"], sc["bound_size"])
    torch.manual_seed(5)
    for lvol in (sc["lvol"], lv2, lv3, sc["lvol"]):
        vin = [sc[k] + 0.05 * torch.randn_like(sc[k]) for k in ("vx", "vy", "vz")]      # new velocities every step
        N.set_option("sparse_setup", 0)
        dense = ViscosityCGSolver3D(sc["gres"], sc["bound_size"])
        vd = [a.clone() for a in vin]
        dense.solve(sc["dt"], 10.0, sc["rho"], *vd, sc["sphi"], None, None, lvol)
        N.set_option("sparse_setup", 1)
        vs = [a.clone() for a in vin]
        sparse.solve(sc["dt"], 10.0, sc["rho"], *vs, sc["sphi"], None, None, lvol)
        assert sparse.iterations == dense.iterations and sparse.delta == dense.delta
        for nm in ("b", "r"):
            for c in "xyz":
                assert torch.equal(getattr(sparse, f"{nm}_{c}"), getattr(dense, f"{nm}_{c}")), (nm, c)
        for a, b in zip(vs, vd):
            assert torch.equal(a, b)
        assert any(not torch.equal(a, b) for a, b in zip(vs, vin))          # (the solve did change something)
