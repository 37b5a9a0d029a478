"""GPU: the implementation switches of the persistent single-reduction CG (fs_set_option) select between result-equivalent
kernels — global-memory form, first and second shared-memory resident form, counter-based or flag-in-data grid reduction,
L2 prefetch in the stand-alone K1s.  Every combination must give the iterates of the default path and the reference's
iteration counts (ViscosityCGSolver3D.py:588-612)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2

pytestmark = pytest.mark.gpu

FORMS = [(0, 0), (1, 0), (2, 0), (2, 1)]          # (resident_form, llred)


@pytest.fixture(autouse=True)
def _restore_options():
    from solver import _native as N
    yield
    for k in ("resident_form", "llred", "k1_prefetch"):
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
    for form, ll in FORMS:
        N.set_option("resident_form", form)
        N.set_option("llred", ll)
        out[(form, ll)] = _solve(sc, 50.0, max_iter=100, tol=0.0)
    it0, d0, x0, _ = out[(0, 0)]
    assert it0 == 100
    for key, (it, d, x, _) in out.items():
        assert it == it0, key
        # same recurrence, different reduction trees (block -> segment maps differ): rounding-level differences that the
        # iteration amplifies, same bars as test_active_set_gpu.test_fixed_window_persistent_counts_iterations
        assert abs(d - d0) <= 1e-3 * abs(d0), (key, d, d0)
        for a, b in zip(x, x0):
            assert rel_l2(a, b) < 1e-6, key


@pytest.mark.parametrize("form,ll", FORMS)
@pytest.mark.parametrize("tag", ["visc3d_solve_8x10x8", "visc3d_solve_stiff_6x8x6"])
def test_forms_vs_reference_fixture(tag, form, ll):
    from solver import _native as N
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    N.set_option("resident_form", form)
    N.set_option("llred", ll)
    f = load_golden(tag)
    s = ViscosityCGSolver3D(f["gres"], f["bound_size"], cg_mode="persistent_sr")
    v = [torch.as_tensor(f[k]).cuda() for k in ("vx", "vy", "vz")]
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    s.solve(float(f["dt"]), float(f["mu"]), float(f["rho"]), *v, dev(f["sphi"]), None, dev(f["lphi"]), dev(f["lvol"]), tol=float(f["tol"]))
    it_ref = int(f["iterations"])
    assert abs(s.iterations - it_ref) <= max(2, round(0.02 * it_ref)), (s.iterations, it_ref)
    for a, n in zip(v, "xyz"):
        assert rel_l2(a.cpu().numpy(), f[f"v{n}_new"]) < 1e-4


@pytest.mark.parametrize("form,ll", [(2, 0), (2, 1)])
def test_second_form_converged_solve_and_repeated_solves(form, ll):
    """Converged solves (several cooperative launches each, sequence numbers carried from launch to launch and from solve to
    solve on the same handle) against the NumPy oracle."""
    import scenes
    from oracle import numpy_oracle as O
    from solver import _native as N
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    N.set_option("resident_form", form)
    N.set_option("llred", ll)
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


def test_k1_prefetch_changes_no_bit():
    """The L2 prefetch of the stand-alone K1s is a hint: bit-identical iterates with and without it (value 2 forces it on
    lists below the HBM-size threshold)."""
    import scenes
    from solver import _native as N
    sc = scenes.buckling(40, device="cuda", mu=50.0, gres=(36, 40, 44))
    res = {}
    for pf in (0, 2):
        N.set_option("k1_prefetch", pf)
        res[pf] = _solve(sc, 50.0, max_iter=60, tol=0.0, cg_mode="kernels_sr", active_set="fluid")
    assert res[0][0] == res[2][0] == 60
    assert res[0][1] == res[2][1]
    for a, b in zip(res[0][2], res[2][2]):
        assert np.array_equal(a, b)
