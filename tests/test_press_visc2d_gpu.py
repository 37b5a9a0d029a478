"""GPU parity tests: pressure (3-D / 2-D), solid fractions and 2-D viscosity against the reference's golden
vectors and the NumPy oracle.  fp64 kernels follow the reference's association without FMA contraction, so the
single-kernel results are expected to be bit-exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2, rel_max

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def _nanfull(ref, dtype=None):
    return torch.full(ref.shape, float("nan"), dtype=dtype or torch.float64, device="cuda")


def _same_written(out, ref, exact=True, tol=0.0):
    out = out.cpu().numpy()
    assert np.array_equal(np.isnan(out), np.isnan(ref)), "set of written entries differs from the reference"
    m = ~np.isnan(ref)
    if exact:
        assert np.array_equal(out[m], ref[m])
    else:
        assert rel_max(out[m], ref[m]) < tol


def test_solidfrac3d_bit_exact():
    from solver.SolidFraction3D import compute_solid_frac
    f = load_golden("solidfrac3d_6x7x8")
    for nm in ("rand", "smooth"):
        w = [_nanfull(f[f"w{c}_{nm}"]) for c in "xyz"]
        compute_solid_frac(f["gres"], _dev(f["sphi_" + nm]), *w)
        for a, c in zip(w, "xyz"):
            _same_written(a, f[f"w{c}_{nm}"])


def test_solidfrac2d_bit_exact():
    from solver.SolidFraction2D import compute_solid_frac
    f = load_golden("solidfrac2d_9x7")
    w = [_nanfull(f["wx"]), _nanfull(f["wy"])]
    compute_solid_frac(f["gres"], _dev(f["sphi"]), *w)
    _same_written(w[0], f["wx"])
    _same_written(w[1], f["wy"])


def test_solidfrac3d_large_vs_oracle():
    from oracle import numpy_oracle as O
    from solver.SolidFraction3D import compute_solid_frac
    import scenes
    sc = scenes.buckling(40, device="cuda")
    g = sc["gres"]
    w = [torch.zeros(s, dtype=torch.float64, device="cuda") for s in ((g[0] + 1, g[1], g[2]), (g[0], g[1] + 1, g[2]), (g[0], g[1], g[2] + 1))]
    compute_solid_frac(g, sc["sphi"], *w)
    ref = [np.zeros(tuple(a.shape)) for a in w]
    O.solidfrac3d(g, sc["sphi"].cpu().numpy(), *ref)
    for a, b in zip(w, ref):
        assert np.array_equal(a.cpu().numpy(), b)
        assert set(np.unique(b)) <= {0.0, 0.5, 0.75, 1.0}


@pytest.mark.parametrize("tag,dims", [("press3d_kernels_7x6x8", "xyz"), ("press2d_kernels_9x7", "xy")])
def test_pressure_kernels_bit_exact(tag, dims):
    if len(dims) == 3:
        from solver import PressureCGSolver3D as P
    else:
        from solver import PressureCGSolver2D as P
    f = load_golden(tag)
    g = f["gres"]
    ws = [_dev(f["w" + c]) for c in dims]
    vel = [_dev(f["v" + c]) for c in dims]           # fp32
    q = _nanfull(f["q"])
    P.matvecmul(g, _dev(f["pv"]), q, *ws, _dev(f["lphi"]))
    _same_written(q, f["q"])
    b = _nanfull(f["b"])
    P.initialize_solver(f["cell_size"], g, *vel, None, _dev(f["sv"]), _dev(f["lphi"]), b, *ws)
    _same_written(b, f["b"])
    u = [v.clone() for v in vel]
    P.apply_pressure(g, f["cell_size"], *u, _dev(f["pv"]), *ws, _dev(f["sv"]), _dev(f["lphi"]))
    for a, c in zip(u, dims):
        assert a.dtype == torch.float32
        assert np.array_equal(a.cpu().numpy(), f["u" + c])


def test_press3d_solve_vs_reference():
    from solver.CGSolverBuffer import CGSolverBuffer
    from solver.PressureCGSolver3D import PressureCGSolver3D
    f = load_golden("press3d_solve_8x10x8")
    buf = CGSolverBuffer(f["gres"])
    s = PressureCGSolver3D(buf, f["gres"], float(f["bound_size"]))      # scalar GDX like the notebook
    v = [_dev(f[k]) for k in ("vx", "vy", "vz")]
    s.solve(*v, _dev(f["sphi"]), _dev(f["sv"]), _dev(f["lphi"]), tol=float(f["tol"]))
    it_ref = int(f["iterations"])
    assert abs(s.iterations - it_ref) <= max(1, round(0.02 * it_ref)), (s.iterations, it_ref)
    for a, k in zip((s.wx, s.wy, s.wz), ("wx", "wy", "wz")):
        assert np.array_equal(a.cpu().numpy(), f[k])
    assert np.array_equal(buf.b.cpu().numpy(), f["b"])
    assert rel_l2(s.x.cpu().numpy(), f["x"]) < 1e-4
    for a, n in zip(v, "xyz"):
        assert rel_l2(a.cpu().numpy(), f[f"v{n}_new"]) < 1e-4
    assert s.delta < float(f["tol"]) ** 2
    # caller-supplied weights path (as the notebook does with DensitySolver.wx/wy/wz)
    v2 = [_dev(f[k]) for k in ("vx", "vy", "vz")]
    s.solve(*v2, _dev(f["sphi"]), _dev(f["sv"]), _dev(f["lphi"]), wx=_dev(f["wx"]), wy=_dev(f["wy"]), wz=_dev(f["wz"]))
    for a, b in zip(v, v2):
        assert torch.equal(a, b)


def test_press2d_solve_vs_reference():
    from solver.CGSolverBuffer import CGSolverBuffer
    from solver.PressureCGSolver2D import PressureCGSolver2D
    f = load_golden("press2d_solve_14x12")
    buf = CGSolverBuffer(f["gres"])
    s = PressureCGSolver2D(buf, f["gres"], f["bound_size"])
    v = [_dev(f[k]) for k in ("vx", "vy")]
    s.solve(*v, _dev(f["sphi"]), _dev(f["sv"]), _dev(f["lphi"]), tol=float(f["tol"]))
    it_ref = int(f["iterations"])
    assert abs(s.iterations - it_ref) <= max(1, round(0.02 * it_ref)), (s.iterations, it_ref)
    assert np.array_equal(s.wx.cpu().numpy(), f["wx"]) and np.array_equal(s.wy.cpu().numpy(), f["wy"])
    for a, n in zip(v, "xy"):
        assert rel_l2(a.cpu().numpy(), f[f"v{n}_new"]) < 1e-4
    # 2-D loop exits silently at max_iter (no raise)
    s.max_iter = 2
    v = [_dev(f[k]) for k in ("vx", "vy")]
    s.solve(*v, _dev(f["sphi"]), _dev(f["sv"]), _dev(f["lphi"]), tol=0.0)
    assert s.iterations == 2


def test_press3d_raises_on_exhaustion():
    from solver.CGSolverBuffer import CGSolverBuffer
    from solver.PressureCGSolver3D import PressureCGSolver3D
    f = load_golden("press3d_solve_8x10x8")
    s = PressureCGSolver3D(CGSolverBuffer(f["gres"]), f["gres"], float(f["bound_size"]))
    s.max_iter = 3
    v = [_dev(f[k]) for k in ("vx", "vy", "vz")]
    with pytest.raises(ValueError, match="Failed to converge!"):
        s.solve(*v, _dev(f["sphi"]), _dev(f["sv"]), _dev(f["lphi"]), tol=0.0)
    assert s.iterations == 3
    for a, k in zip(v, ("vx", "vy", "vz")):
        assert np.array_equal(a.cpu().numpy(), f[k])           # no update applied


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_visc2d_kernels_vs_reference(dtype):
    from solver import ViscosityCGSolver2D as V
    f = load_golden("visc2d_kernels_9x7")
    g = tuple(int(n) for n in f["gres"])
    sc, mu = float(f["scale"]), float(f["mu"])
    v = [_dev(f[k]) for k in ("vx", "vy")]
    sphi, vol = _dev(f["sphi"]), _dev(f["vol"])
    q = [_nanfull(f["qx"]), _nanfull(f["qy"])]
    V.matvecmul(g, sc, mu, *v, *q, sphi, vol, dtype=dtype)
    b = [_nanfull(f["bx"]), _nanfull(f["by"])]
    V.initialize_solver(g, sc, mu, *v, sphi, None, vol, *b, dtype=dtype)
    for c, n in enumerate("xy"):
        _same_written(q[c], f["q" + n], exact=(dtype == torch.float64), tol=1e-5)
        _same_written(b[c], f["b" + n], exact=(dtype == torch.float64), tol=1e-5)
    wb = [_nanfull(f["wbx"], torch.float32), _nanfull(f["wby"], torch.float32)]
    V.apply_viscosity(g, *wb, *v, sphi, None)
    _same_written(wb[0], f["wbx"])
    _same_written(wb[1], f["wby"])


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_visc2d_solve_vs_reference(dtype):
    from solver.ViscosityCGSolver2D import ViscosityCGSolver2D
    f = load_golden("visc2d_solve_14x12")
    s = ViscosityCGSolver2D(f["gres"], f["bound_size"], dtype=dtype)
    assert s.cell_vol == float(f["cell_vol"])
    v = [_dev(f[k]) for k in ("vx", "vy")]
    s.solve(float(f["dt"]), float(f["mu"]), float(f["rho"]), *v, _dev(f["sphi"]), None, _dev(f["lphi"]), _dev(f["lvol"]))
    it_ref = int(f["iterations"])
    assert abs(s.iterations - it_ref) <= max(1, round(0.02 * it_ref)), (s.iterations, it_ref)
    for a, n in zip(v, "xy"):
        assert rel_l2(a.cpu().numpy(), f[f"v{n}_new"]) < 1e-4


def test_2d_scene_vs_oracle():
    """Config 2 at a size the oracle finishes in seconds: Visc2D then SolidFraction2D + Press2D."""
    import scenes
    from oracle import numpy_oracle as O
    from solver.CGSolverBuffer import CGSolverBuffer
    from solver.PressureCGSolver2D import PressureCGSolver2D
    from solver.ViscosityCGSolver2D import ViscosityCGSolver2D
    sc = scenes.box2d(64, device="cuda")
    g = sc["gres"]
    ov = O.ViscosityCGSolver2D(g, sc["bound_size"])
    rv = [sc[k].cpu().numpy().copy() for k in ("vx", "vy")]
    ov.solve(sc["dt"], sc["mu"], sc["rho"], *rv, sc["sphi"].cpu().numpy(), None, None, sc["lvol"].cpu().numpy())
    op = O.PressureCGSolver2D(O.CGSolverBuffer(g), g, sc["bound_size"])
    op.solve(*rv, sc["sphi"].cpu().numpy(), sc["sv"].cpu().numpy(), sc["lphi"].cpu().numpy())
    sv_ = ViscosityCGSolver2D(g, sc["bound_size"])
    v = [sc[k].clone() for k in ("vx", "vy")]
    sv_.solve(sc["dt"], sc["mu"], sc["rho"], *v, sc["sphi"], None, None, sc["lvol"])
    sp = PressureCGSolver2D(CGSolverBuffer(g), g, sc["bound_size"])
    sp.solve(*v, sc["sphi"], sc["sv"], sc["lphi"])
    assert abs(sv_.iterations - ov.trace.iterations) <= max(1, round(0.02 * ov.trace.iterations))
    assert abs(sp.iterations - op.trace.iterations) <= max(1, round(0.02 * op.trace.iterations))
    for a, b in zip(v, rv):
        assert rel_l2(a.cpu().numpy(), b) < 1e-4


def test_full_step_3d_vs_oracle():
    """Config 3's pipeline (SolidFraction3D -> Viscosity -> Pressure) at 24^3 against the oracle."""
    import scenes
    from oracle import numpy_oracle as O
    from solver.CGSolverBuffer import CGSolverBuffer
    from solver.PressureCGSolver3D import PressureCGSolver3D
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    sc = scenes.buckling(24, device="cuda", with_sv=True)
    g, dx = sc["gres"], sc["dx"]
    ov = O.ViscosityCGSolver3D(g, sc["bound_size"])
    rv = [sc[k].cpu().numpy().copy() for k in ("vx", "vy", "vz")]
    sphi, lvol, lphi, sv = (sc[k].cpu().numpy() for k in ("sphi", "lvol", "lphi", "sv"))
    ov.solve(sc["dt"], sc["mu"], sc["rho"], *rv, sphi, sv, lphi, lvol)
    op = O.PressureCGSolver3D(O.CGSolverBuffer(g), g, dx)
    op.solve(*rv, sphi, sv, lphi)
    v = [sc[k].clone() for k in ("vx", "vy", "vz")]
    s1 = ViscosityCGSolver3D(g, sc["bound_size"])
    s1.solve(sc["dt"], sc["mu"], sc["rho"], *v, sc["sphi"], sc["sv"], sc["lphi"], sc["lvol"])
    s2 = PressureCGSolver3D(CGSolverBuffer(g), g, dx)
    s2.solve(*v, sc["sphi"], sc["sv"], sc["lphi"])
    assert abs(s1.iterations - ov.trace.iterations) <= max(1, round(0.02 * ov.trace.iterations))
    assert abs(s2.iterations - op.trace.iterations) <= max(1, round(0.02 * op.trace.iterations)), (s2.iterations, op.trace.iterations)
    for a, b in zip(v, rv):
        assert rel_l2(a.cpu().numpy(), b) < 1e-4
