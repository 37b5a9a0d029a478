"""CPU: the NumPy oracle against golden vectors produced by the reference's own kernels
(oracle/refsim/run_reference.py, Numba CUDA simulator).  This is what pins the oracle."""
import numpy as np
import pytest

from conftest import load_golden, rel_l2
from oracle import numpy_oracle as O


def _eq(a, b):
    """bit-exact on the entries the reference wrote; untouched (NaN-poisoned) entries stay untouched"""
    a = np.asarray(a)
    b = np.asarray(b)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(b)
    assert np.array_equal(a[m], b[m])


def test_visc3d_kernels_bit_exact():
    f = load_golden("visc3d_kernels_6x7x8")
    g, sc, mu = f["gres"], float(f["scale"]), float(f["mu"])
    shapes = [f["vx"].shape, f["vy"].shape, f["vz"].shape]
    q = [np.full(s, np.nan) for s in shapes]
    O.visc3d_matvecmul(g, sc, mu, f["vx"], f["vy"], f["vz"], *q, f["sphi"], f["vol"])
    b = [np.full(s, np.nan) for s in shapes]
    O.visc3d_initialize_solver(g, sc, mu, f["vx"], f["vy"], f["vz"], f["sphi"], None, f["vol"], *b)
    e = [f["vx"].copy(), f["vy"].copy(), f["vz"].copy()]
    O.visc3d_extrapolate(g, 3, *e, f["sphi"])
    wb = [np.full(s, np.nan, dtype=np.float32) for s in shapes]
    O.visc3d_apply_viscosity(g, *wb, f["vx"], f["vy"], f["vz"], f["sphi"], None)
    for c, n in enumerate("xyz"):
        _eq(q[c], f["q" + n])
        _eq(b[c], f["b" + n])
        _eq(e[c], f["e" + n])
        _eq(wb[c], f["wb" + n])


def test_visc2d_kernels_bit_exact():
    f = load_golden("visc2d_kernels_9x7")
    g, sc, mu = f["gres"], float(f["scale"]), float(f["mu"])
    shapes = [f["vx"].shape, f["vy"].shape]
    q = [np.full(s, np.nan) for s in shapes]
    O.visc2d_matvecmul(g, sc, mu, f["vx"], f["vy"], *q, f["sphi"], f["vol"])
    b = [np.full(s, np.nan) for s in shapes]
    O.visc2d_initialize_solver(g, sc, mu, f["vx"], f["vy"], f["sphi"], None, f["vol"], *b)
    wb = [np.full(s, np.nan, dtype=np.float32) for s in shapes]
    O.visc2d_apply_viscosity(g, *wb, f["vx"], f["vy"], f["sphi"], None)
    for c, n in enumerate("xy"):
        _eq(q[c], f["q" + n])
        _eq(b[c], f["b" + n])
        _eq(wb[c], f["wb" + n])


def test_solidfrac_bit_exact_and_quantised():
    f = load_golden("solidfrac3d_6x7x8")
    for nm in ("rand", "smooth"):
        w = [np.full(f[f"w{c}_{nm}"].shape, np.nan) for c in "xyz"]
        O.solidfrac3d(f["gres"], f["sphi_" + nm], *w)
        for a, c in zip(w, "xyz"):
            _eq(a, f[f"w{c}_{nm}"])
            vals = np.unique(a[~np.isnan(a)])
            assert set(vals) <= {0.0, 0.5, 0.75, 1.0}
        # far planes never written (B-8)
        assert np.isnan(w[0][-1]).all() and np.isnan(w[1][:, -1]).all() and np.isnan(w[2][:, :, -1]).all()
    f = load_golden("solidfrac2d_9x7")
    w = [np.full(f["wx"].shape, np.nan), np.full(f["wy"].shape, np.nan)]
    O.solidfrac2d(f["gres"], f["sphi"], *w)
    _eq(w[0], f["wx"])
    _eq(w[1], f["wy"])


@pytest.mark.parametrize("tag,dims", [("press3d_kernels_7x6x8", "xyz"), ("press2d_kernels_9x7", "xy")])
def test_pressure_kernels_bit_exact(tag, dims):
    f = load_golden(tag)
    g = f["gres"]
    ws = [f["w" + c] for c in dims]
    vel = [f["v" + c] for c in dims]
    q = np.full(tuple(g), np.nan)
    O.press_matvecmul(g, f["pv"], q, ws, f["lphi"])
    _eq(q, f["q"])
    b = np.full(tuple(g), np.nan)
    O.press_initialize_solver(f["cell_size"], g, vel, None, f["sv"], f["lphi"], b, ws)
    _eq(b, f["b"])
    u = [v.copy() for v in vel]
    O.press_apply_pressure(g, f["cell_size"], u, f["pv"], ws, f["sv"], f["lphi"])
    for a, c in zip(u, dims):
        _eq(a, f["u" + c])


@pytest.mark.parametrize("tag", ["visc3d_solve_8x10x8", "visc3d_solve_stiff_6x8x6"])
def test_visc3d_solve_matches_reference(tag):
    f = load_golden(tag)
    s = O.ViscosityCGSolver3D(f["gres"], f["bound_size"])
    v = [f["vx"].copy(), f["vy"].copy(), f["vz"].copy()]
    s.solve(float(f["dt"]), float(f["mu"]), float(f["rho"]), *v, f["sphi"], None, f["lphi"], f["lvol"], tol=float(f["tol"]))
    assert s.trace.iterations == int(f["iterations"])
    assert s.cell_vol == float(f["cell_vol"])
    for a, n in zip(v, "xyz"):
        assert np.array_equal(a, f[f"v{n}_new"])                 # fp32 write-back, bit-exact
    for a, n in zip((s.x_x, s.x_y, s.x_z), "xyz"):
        assert rel_l2(a, f["x_" + n]) < 1e-13
    assert abs(s.delta - float(f["delta"])) <= 1e-9 * float(f["delta"])


def test_visc2d_solve_matches_reference():
    f = load_golden("visc2d_solve_14x12")
    s = O.ViscosityCGSolver2D(f["gres"], f["bound_size"])
    v = [f["vx"].copy(), f["vy"].copy()]
    s.solve(float(f["dt"]), float(f["mu"]), float(f["rho"]), *v, f["sphi"], None, f["lphi"], f["lvol"])
    assert s.trace.iterations == int(f["iterations"])
    for a, n in zip(v, "xy"):
        assert np.array_equal(a, f[f"v{n}_new"])


def test_press3d_solve_matches_reference():
    f = load_golden("press3d_solve_8x10x8")
    buf = O.CGSolverBuffer(f["gres"])
    s = O.PressureCGSolver3D(buf, f["gres"], float(f["bound_size"]))
    v = [f["vx"].copy(), f["vy"].copy(), f["vz"].copy()]
    s.solve(*v, f["sphi"], f["sv"], f["lphi"], tol=float(f["tol"]))
    assert s.trace.iterations == int(f["iterations"])
    assert np.array_equal(s.wx, f["wx"]) and np.array_equal(s.wy, f["wy"]) and np.array_equal(s.wz, f["wz"])
    assert rel_l2(s.x, f["x"]) < 1e-12
    for a, n in zip(v, "xyz"):
        assert rel_l2(a, f[f"v{n}_new"]) < 1e-6
    assert rel_l2(buf.b, f["b"]) < 1e-14


def test_press2d_solve_matches_reference():
    f = load_golden("press2d_solve_14x12")
    buf = O.CGSolverBuffer(f["gres"])
    s = O.PressureCGSolver2D(buf, f["gres"], f["bound_size"])
    v = [f["vx"].copy(), f["vy"].copy()]
    s.solve(*v, f["sphi"], f["sv"], f["lphi"], tol=float(f["tol"]))
    assert s.trace.iterations == int(f["iterations"])
    assert np.array_equal(s.wx, f["wx"]) and np.array_equal(s.wy, f["wy"])
    for a, n in zip(v, "xy"):
        assert rel_l2(a, f[f"v{n}_new"]) < 1e-6


def test_operator_properties():
    """Symmetry on interior fluid rows and constants in the null space of the viscous part."""
    rng = np.random.default_rng(7)
    g = (6, 7, 8)
    fine = tuple(2 * n + 1 for n in g)
    sphi = rng.standard_normal(fine)
    vol = rng.random(fine)
    shapes = [(g[0] + 1, g[1], g[2]), (g[0], g[1] + 1, g[2]), (g[0], g[1], g[2] + 1)]

    def rand_fluid_interior():
        out = []
        for c, s in enumerate(shapes):
            a = np.zeros(s)
            a[1:-1, 1:-1, 1:-1] = rng.standard_normal(tuple(n - 2 for n in s))
            off = [(0, 1, 1), (1, 0, 1), (1, 1, 0)][c]
            m = sphi[off[0]::2, off[1]::2, off[2]::2][: s[0], : s[1], : s[2]] >= 0
            out.append(a * m)
        return out

    a, b = rand_fluid_interior(), rand_fluid_interior()
    Aa = [np.zeros(s) for s in shapes]
    Ab = [np.zeros(s) for s in shapes]
    O.visc3d_matvecmul(g, 0.7, 1.3, *a, *Aa, sphi, vol)
    O.visc3d_matvecmul(g, 0.7, 1.3, *b, *Ab, sphi, vol)
    lhs = sum(float(np.sum(x * y)) for x, y in zip(b, Aa))
    rhs = sum(float(np.sum(x * y)) for x, y in zip(a, Ab))
    assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), 1.0)
    ones = [np.full(s, 3.0) for s in shapes]
    q = [np.zeros(s) for s in shapes]
    O.visc3d_matvecmul(g, 0.7, 1.3, *ones, *q, np.ones(fine), np.ones(fine))
    for c, s in enumerate(shapes):
        np.testing.assert_allclose(q[c][1:-1, 1:-1, 1:-1], 3.0, rtol=1e-13)
        assert q[c][0].max() == 0 and q[c][-1].max() == 0


def test_density3d_kernels_vs_reference():
    """§8 f-1: DensityCGSolver3D kernels.  The particle scatter accumulates with atomics, so mass/volume agree to rounding;
    every per-cell kernel downstream is checked on the reference's own inputs and is bit-exact."""
    f = load_golden("density3d_kernels_7x8x6")
    g = f["gres"]
    cell = np.full(3, float(f["dx"]))
    ws = [f["wx"], f["wy"], f["wz"]]
    gm, gvol = np.zeros(tuple(g)), np.zeros(tuple(g))
    O.density_initialize_density(np.zeros(3), cell, g, f["px"], f["pm"], float(f["pvol"]), gm, gvol)
    assert rel_l2(gm, f["gm"]) < 1e-14 and rel_l2(gvol, f["gvol"]) < 1e-14
    assert abs(gm.sum() - f["pm"].sum()) < 1e-12 * f["pm"].sum()            # partition of unity
    fixed = f["gvol"].copy()
    O.density_fix_volume(cell, g, f["lvol"], fixed, f["sphi"], f["lphi"], ws)
    assert np.array_equal(fixed, f["gvol_fixed"])
    b = np.full(tuple(g), np.nan)
    O.density_initialize_solver(1000.0, 1.0 / 300, g, cell, f["gm"], f["gvol_fixed"], f["lphi"], ws, b)
    _eq(b, f["b"])
    q = np.full(tuple(g), np.nan)
    O.density_matvecmul(g, f["pv"], q, ws, f["lphi"])
    _eq(q, f["q"])
    disp = [np.full(f["disp" + c].shape, np.nan) for c in "xyz"]
    O.density_compute_displacement(g, 1.0 / 300, cell, disp, f["pv"], f["lphi"])
    for a, c in zip(disp, "xyz"):
        _eq(a, f["disp" + c])
    px = f["px"].copy()
    bias = ((0, 0.5, 0.5), (0.5, 0, 0.5), (0.5, 0.5, 0))
    for a, c in enumerate("xyz"):
        O.density_apply_displacement(px, f["df" + c], np.zeros(3), cell, bias[a], a)
    assert rel_l2(px, f["pmoved"]) < 1e-15


def test_density3d_solve_matches_reference():
    f = load_golden("density3d_solve_8x10x8")
    buf = O.CGSolverBuffer(f["gres"])
    s = O.DensityCGSolver3D(buf, f["gres"], np.zeros(3), f["bound_size"])
    px = f["px"].copy()
    s.solve(1000.0, 1.0 / 300, px, f["pm"], float(f["pvol"]), None, None, None, f["sphi"], None, f["lphi"], f["lvol"], tol=float(f["tol"]))
    assert s.trace.iterations == int(f["iterations"])
    assert rel_l2(s.m, f["m"]) < 1e-13 and rel_l2(s.vol, f["vol"]) < 1e-13
    assert rel_l2(buf.b, f["b"]) < 1e-10
    assert rel_l2(s.x, f["x"]) < 1e-8
    assert rel_l2(px, f["px_new"]) < 1e-12
    assert np.array_equal(s.wx, f["wx"])
