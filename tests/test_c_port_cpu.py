"""CPU: the C/OpenMP restatement (oracle/c_port) against the NumPy oracle and the reference's golden vectors."""
import numpy as np
import pytest

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


def test_c_port_rhs_and_extrapolate_bit_exact_vs_reference():
    """the C restatements of initialize_solver (:504-513) and extrapolate (:472-502) against the reference's own outputs"""
    f = load_golden("visc3d_kernels_6x7x8")
    b = [np.full(f[k].shape, np.nan) for k in ("vx", "vy", "vz")]
    c_port.initialize_solver(f["gres"], float(f["scale"]), float(f["mu"]), f["vx"], f["vy"], f["vz"], f["sphi"], None, f["vol"], *b)
    for a, n in zip(b, "xyz"):
        ref = f["b" + n]
        assert np.array_equal(np.isnan(a), np.isnan(ref))
        assert np.array_equal(a[~np.isnan(ref)], ref[~np.isnan(ref)])
    v = [f[k].copy() for k in ("vx", "vy", "vz")]
    c_port.extrapolate(f["gres"], 3, *v, f["sphi"])
    for a, n in zip(v, "xyz"):
        assert np.array_equal(a, f["e" + n])


def test_c_port_setup_vs_numpy_oracle_scene():
    """same two functions on the benchmark scene (non-trivial solid geometry), bit for bit against the NumPy oracle"""
    import scenes
    sc = scenes.buckling(20, mu=10.0)
    g = sc["gres"]
    sphi = sc["sphi"].numpy()
    vol = np.ascontiguousarray(sc["lvol"].numpy() / 3.0e-7)
    va = [sc[k].numpy().astype(np.float64) for k in ("vx", "vy", "vz")]
    vb = [a.copy() for a in va]
    O.visc3d_extrapolate(g, 3, *va, sphi)
    c_port.extrapolate(g, 3, *vb, sphi)
    for a, b in zip(va, vb):
        assert np.array_equal(a, b)
    ba = [np.zeros_like(a) for a in va]
    bb = [np.zeros_like(a) for a in va]
    O.visc3d_initialize_solver(g, 1.7, 10.0, *va, sphi, None, vol, *ba)
    c_port.initialize_solver(g, 1.7, 10.0, *vb, sphi, None, vol, *bb)
    for a, b in zip(ba, bb):
        assert np.array_equal(a, b)
    assert any(np.any(a != 0) for a in ba)


def test_c_port_thread_control():
    n = c_port.use_all_cores()
    assert n >= 1 and n == c_port.num_threads()


@pytest.mark.parametrize("threads", [1, 3, None])
def test_c_port_solve_vs_reference(threads):
    """The port's dot products follow NumPy's pairwise summation order (what the fixtures' cp.sum evaluated to), whatever the
    OpenMP team: the reference's iteration counts are reproduced EXACTLY and run to run, also on the stiff fixture whose
    count moves by three with the summation order."""
    if threads is None:
        c_port.use_all_cores()
    else:
        c_port.load().port_set_num_threads(threads)
    try:
        for tag in ("visc3d_solve_8x10x8", "visc3d_solve_stiff_6x8x6"):
            f = load_golden(tag)
            s = c_port.ViscosityCGSolver3D(f["gres"], f["bound_size"])
            v = [f["vx"].copy(), f["vy"].copy(), f["vz"].copy()]
            s.solve(float(f["dt"]), float(f["mu"]), float(f["rho"]), *v, f["sphi"], None, f["lphi"], f["lvol"], tol=float(f["tol"]))
            assert s.iterations == int(f["iterations"]), (s.iterations, int(f["iterations"]))
            for a, n in zip(v, "xyz"):
                assert rel_l2(a, f[f"v{n}_new"]) < 1e-4
    finally:
        c_port.use_all_cores()


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
    # same kernels bit for bit, same summation order: the two oracles agree in every bit of the solution
    assert a.trace.iterations == b.iterations
    for x, y in zip(va, vb):
        assert np.array_equal(x, y)
