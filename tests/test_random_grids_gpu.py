"""GPU parity on randomized and degenerate grids (edge cases: extents 1..3, z extents that are not multiples of 4,
strongly non-cubic shapes) against the NumPy oracle.  fp64 single-kernel results are expected bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES_3D = [(1, 1, 1), (2, 3, 1), (3, 3, 3), (4, 3, 5), (3, 9, 4), (5, 4, 13), (11, 3, 7), (2, 2, 17), (7, 8, 9)]
SHAPES_2D = [(1, 1), (2, 2), (3, 3), (3, 8), (9, 4), (5, 13), (16, 3)]


def _mac(g):
    return [tuple(n + (1 if i == a else 0) for i, n in enumerate(g)) for a in range(len(g))]


def _fine(g):
    return tuple(2 * n + 1 for n in g)


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("g", SHAPES_3D)
def test_visc3d_kernels_random(g):
    from oracle import numpy_oracle as O
    from solver import ViscosityCGSolver3D as V
    rng = np.random.default_rng(hash(g) % 2 ** 31)
    sphi = rng.standard_normal(_fine(g))
    sphi[rng.random(_fine(g)) < 0.1] = 0.0
    vol = rng.random(_fine(g))
    vel = [rng.standard_normal(s) for s in _mac(g)]
    sc, mu = 0.37, 2.1
    for fn_gpu, fn_ref in ((lambda o: V.matvecmul(g, sc, mu, *[_dev(v) for v in vel], *o, _dev(sphi), _dev(vol)),
                            lambda o: O.visc3d_matvecmul(g, sc, mu, *vel, *o, sphi, vol)),
                           (lambda o: V.initialize_solver(g, sc, mu, *[_dev(v) for v in vel], _dev(sphi), None, _dev(vol), *o),
                            lambda o: O.visc3d_initialize_solver(g, sc, mu, *vel, sphi, None, vol, *o))):
        out = [torch.full(s, float("nan"), dtype=torch.float64, device="cuda") for s in _mac(g)]
        ref = [np.full(s, np.nan) for s in _mac(g)]
        fn_gpu(out)
        fn_ref(ref)
        for a, b in zip(out, ref):
            a = a.cpu().numpy()
            assert np.array_equal(np.isnan(a), np.isnan(b))
            assert np.array_equal(a[~np.isnan(b)], b[~np.isnan(b)])
    ex = [_dev(v) for v in vel]
    V.extrapolate(g, 3, *ex, _dev(sphi))
    er = [v.copy() for v in vel]
    O.visc3d_extrapolate(g, 3, *er, sphi)
    for a, b in zip(ex, er):
        np.testing.assert_allclose(a.cpu().numpy(), b, rtol=1e-15, atol=0)


@pytest.mark.parametrize("g", SHAPES_2D)
def test_visc2d_kernels_random(g):
    from oracle import numpy_oracle as O
    from solver import ViscosityCGSolver2D as V
    rng = np.random.default_rng(hash(g) % 2 ** 31)
    sphi = rng.standard_normal(_fine(g))
    sphi[rng.random(_fine(g)) < 0.1] = 0.0
    vol = rng.random(_fine(g))
    vel = [rng.standard_normal(s) for s in _mac(g)]
    out = [torch.full(s, float("nan"), dtype=torch.float64, device="cuda") for s in _mac(g)]
    ref = [np.full(s, np.nan) for s in _mac(g)]
    V.matvecmul(g, 0.8, 1.7, *[_dev(v) for v in vel], *out, _dev(sphi), _dev(vol))
    O.visc2d_matvecmul(g, 0.8, 1.7, *vel, *ref, sphi, vol)
    for a, b in zip(out, ref):
        a = a.cpu().numpy()
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.array_equal(a[~np.isnan(b)], b[~np.isnan(b)])


@pytest.mark.parametrize("g", SHAPES_3D + SHAPES_2D)
def test_pressure_and_solidfrac_random(g):
    from oracle import numpy_oracle as O
    d = len(g)
    if d == 3:
        from solver import PressureCGSolver3D as P
        from solver.SolidFraction3D import compute_solid_frac
        sf_ref = O.solidfrac3d
    else:
        from solver import PressureCGSolver2D as P
        from solver.SolidFraction2D import compute_solid_frac
        sf_ref = O.solidfrac2d
    rng = np.random.default_rng(hash(g) % 2 ** 31 + 5)
    sphi = rng.standard_normal(_fine(g))
    sphi[rng.random(_fine(g)) < 0.1] = 0.0
    w = [torch.full(s, float("nan"), dtype=torch.float64, device="cuda") for s in _mac(g)]
    wr = [np.full(s, np.nan) for s in _mac(g)]
    compute_solid_frac(g, _dev(sphi), *w)
    sf_ref(g, sphi, *wr)
    for a, b in zip(w, wr):
        a = a.cpu().numpy()
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.array_equal(a[~np.isnan(b)], b[~np.isnan(b)])
    ws = [np.nan_to_num(x, nan=0.0) for x in wr]
    lphi = rng.standard_normal(g) - 0.2
    pv = rng.standard_normal(g)
    vel = [rng.standard_normal(s).astype(np.float32) for s in _mac(g)]
    sv = rng.standard_normal(_fine(g) + (d,))
    cell = rng.random(d) + 0.5
    q = torch.full(g, float("nan"), dtype=torch.float64, device="cuda")
    qr = np.full(g, np.nan)
    P.matvecmul(g, _dev(pv), q, *[_dev(x) for x in ws], _dev(lphi))
    O.press_matvecmul(g, pv, qr, ws, lphi)
    b = torch.full(g, float("nan"), dtype=torch.float64, device="cuda")
    br = np.full(g, np.nan)
    P.initialize_solver(cell, g, *[_dev(v) for v in vel], None, _dev(sv), _dev(lphi), b, *[_dev(x) for x in ws])
    O.press_initialize_solver(cell, g, vel, None, sv, lphi, br, ws)
    for a, r in ((q, qr), (b, br)):
        a = a.cpu().numpy()
        assert np.array_equal(np.isnan(a), np.isnan(r))
        assert np.array_equal(a[~np.isnan(r)], r[~np.isnan(r)])
    u = [_dev(v) for v in vel]
    ur = [v.copy() for v in vel]
    P.apply_pressure(g, cell, *u, _dev(pv), *[_dev(x) for x in ws], _dev(sv), _dev(lphi))
    O.press_apply_pressure(g, cell, ur, pv, ws, sv, lphi)
    for a, r in zip(u, ur):
        assert np.array_equal(a.cpu().numpy(), r)


@pytest.mark.parametrize("g", [(3, 3, 3), (4, 5, 6), (2, 2, 2), (1, 4, 4)])
def test_degenerate_solves_do_not_crash(g):
    """grids with no interior rows: delta0 == 0 -> zero iterations, velocities untouched (reference: loop skipped)"""
    from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
    rng = np.random.default_rng(3)
    s = ViscosityCGSolver3D(g, (0.1 * g[0], 0.1 * g[1], 0.1 * g[2]))
    v = [_dev(rng.standard_normal(sh).astype(np.float32)) for sh in _mac(g)]
    v0 = [a.clone() for a in v]
    sphi = torch.ones(_fine(g), dtype=torch.float64, device="cuda")
    lvol = torch.full(_fine(g), 1e-6, dtype=torch.float64, device="cuda")
    if min(g) < 3:
        s.solve(0.01, 1.0, 1000.0, *v, sphi, None, None, lvol)
        assert s.iterations == 0
        for a, b in zip(v, v0):
            assert torch.equal(a, b)
    else:
        s.solve(0.01, 1.0, 1000.0, *v, sphi, None, None, lvol)
        assert s.delta < 1e-6


def test_graph_and_plain_launch_paths_agree():
    """iterations replayed from a CUDA graph give the same bits as launch-by-launch execution"""
    import os
    import subprocess
    import sys
    code = r'''
import sys, torch, hashlib
sys.path.insert(0, "python-fluid-simulation_b200")
import scenes
from solver.ViscosityCGSolver3D import ViscosityCGSolver3D
sc = scenes.buckling(32, device="cuda", mu=100.0)
s = ViscosityCGSolver3D(sc["gres"], sc["bound_size"])
v = [sc[k].clone() for k in ("vx", "vy", "vz")]
s.solve(sc["dt"], 100.0, sc["rho"], *v, sc["sphi"], None, None, sc["lvol"])
h = hashlib.sha256(b"".join(a.cpu().numpy().tobytes() for a in v)).hexdigest()
print(s.iterations, h)
'''
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for g in ("1", "0"):
        env = dict(os.environ, FLUIDSOLVER_B200_GRAPH=g)
        r = subprocess.run([sys.executable, "-c", code], cwd=repo, env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1], outs
    assert int(outs[0].split()[0]) > 16          # more than one graph batch was replayed
