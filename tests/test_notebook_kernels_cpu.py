"""CPU: the NumPy restatement of the notebook's grid-side kernels (oracle/numpy_oracle_nb.py) against fixtures produced by
the notebook's own code cells under Numba's CUDA simulator (oracle/refsim/run_notebook_kernels.py).

Index sets are exact.  Scattered fp32 quantities agree to 1e-5 of the field's maximum: the reference accumulates with
floating-point atomics into fp32 arrays in thread order, the restatement in fp64 with one final rounding."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import numpy_oracle_nb as NB

CASES = ["nb_kernels_6x7x8", "nb_kernels_9x8x7"]


def close(a, b, tol=1e-5):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(b)
    scale = max(np.max(np.abs(b[m])), 1e-300) if m.any() else 1.0
    err = np.max(np.abs(a[m] - b[m])) / scale if m.any() else 0.0
    assert err < tol, err


@pytest.mark.parametrize("tag", CASES)
def test_levelset_and_volume(tag):
    f = load_golden(tag)
    g = [int(n) for n in f["gres"]]
    phi = NB.fluid_levelset(g, f["bound_min"], f["bound_size"] / f["gres"], f["px"], float(f["dx"]))
    assert np.array_equal(phi == 3 * float(f["dx"]), f["lphi"] == 3 * float(f["dx"]))      # cells no particle reaches
    close(phi, f["lphi"], 1e-6)
    res = [2 * n + 1 for n in g]
    vol = NB.fluid_volume(res, f["bound_min"], f["bound_size"] / (2 * f["gres"]), f["px"], float(f["pvol"]))
    assert np.array_equal(vol > 0, f["lvol"] > 0)
    close(vol, f["lvol"], 1e-6)


@pytest.mark.parametrize("tag", CASES)
def test_p2g_extrapolate_boundary_g2p(tag):
    f = load_golden(tag)
    g = [int(n) for n in f["gres"]]
    cell = f["cell_size"]
    shapes = [f[f"p2g_m{a}"].shape for a in "xyz"]
    ms = [np.zeros(s, dtype=np.float32) for s in shapes]
    vs = [np.zeros(s, dtype=np.float32) for s in shapes]
    NB.p2g(g, f["bound_min"], cell, f["px"], f["pm"], f["pv"], (f["cx"], f["cy"], f["cz"]), ms, vs)
    for m, v, a in zip(ms, vs, "xyz"):
        assert np.array_equal(m > 0, f[f"p2g_m{a}"] > 0)
        close(m, f[f"p2g_m{a}"])
        close(v, f[f"p2g_v{a}"])
    # continue from the fixture's own state so that every stage is checked in isolation
    ms = [f[f"p2g_m{a}"].copy() for a in "xyz"]
    vs = [f[f"p2g_v{a}"].copy() for a in "xyz"]
    NB.extrapolate(2, vs, ms)
    for v, a in zip(vs, "xyz"):
        assert np.array_equal(v != f[f"p2g_v{a}"], f[f"ext_v{a}"] != f[f"p2g_v{a}"])      # same faces filled
        close(v, f[f"ext_v{a}"], 1e-6)
    vs = [f[f"ext_v{a}"].copy() for a in "xyz"]
    dv = NB.boundary_condition(g, float(f["dx"]), vs, ms, f["sphi"], f["sv"])
    for d, a in zip(dv, "xyz"):
        ref = f[f"bc_dv{a}"]
        assert not np.isnan(ref).any(), "every in-bounds entry is written by the reference"
        assert np.array_equal(d != 0, ref != 0)
        close(d, ref)
    vgrid = [f[f"bc_v{a}"] for a in "xyz"]
    pv, pc = NB.g2p(g, f["bound_min"], cell, f["px"], vgrid)
    close(pv, f["g2p_pv"])
    for c, k in zip(pc, ("g2p_cx", "g2p_cy", "g2p_cz")):
        close(c, f[k])
