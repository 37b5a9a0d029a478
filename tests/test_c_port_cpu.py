"""CPU: the C/OpenMP restatement (oracle/c_port) against the NumPy oracle and the reference's golden vectors."""
import numpy as np

from conftest import load_golden, rel_l2
from oracle import c_port
from oracle import numpy_oracle as O


def test_c_port_apply_bit_exact_vs_reference():
    f = load_golden("visc3d_kernels_6x7x8")
    shapes = [f["vx"].shape, f["vy"].shape, f["vz"].shape]
    q = [np.full(s, np.nan) for s in shapes]
    c_port.matvecmul(f["gres"], float(f["scale"]), float(f["mu"]), f["vx"], f["vy"], f["vz"], *q, f["sphi"], f["vol"])
    for a, n in zip(q, "xyz"):
        ref = f["q" + n]
        assert np.array_equal(np.isnan(a), np.isnan(ref))
        assert np.array_equal(a[~np.isnan(ref)], ref[~np.isnan(ref)])


def test_c_port_solve_vs_reference():
    for tag in ("visc3d_solve_8x10x8", "visc3d_solve_stiff_6x8x6"):
        f = load_golden(tag)
        s = c_port.ViscosityCGSolver3D(f["gres"], f["bound_size"])
        v = [f["vx"].copy(), f["vy"].copy(), f["vz"].copy()]
        s.solve(float(f["dt"]), float(f["mu"]), float(f["rho"]), *v, f["sphi"], None, f["lphi"], f["lvol"], tol=float(f["tol"]))
        it_ref = int(f["iterations"])
        assert abs(s.iterations - it_ref) <= max(1, round(0.02 * it_ref)), (s.iterations, it_ref)
        for a, n in zip(v, "xyz"):
            assert rel_l2(a, f[f"v{n}_new"]) < 1e-4


def test_c_port_vs_numpy_oracle_scene():
    import scenes
    sc = scenes.buckling(20, mu=10.0)
    args = (sc["dt"], sc["mu"], sc["rho"])
    a = O.ViscosityCGSolver3D(sc["gres"], sc["bound_size"])
    va = [sc[k].numpy().copy() for k in ("vx", "vy", "vz")]
    a.solve(*args, *va, sc["sphi"].numpy(), None, None, sc["lvol"].numpy())
    b = c_port.ViscosityCGSolver3D(sc["gres"], sc["bound_size"])
    vb = [sc[k].numpy().copy() for k in ("vx", "vy", "vz")]
    b.solve(*args, *vb, sc["sphi"].numpy(), None, None, sc["lvol"].numpy())
    assert abs(a.trace.iterations - b.iterations) <= max(1, round(0.02 * a.trace.iterations))
    for x, y in zip(va, vb):
        assert rel_l2(x, y) < 1e-4
