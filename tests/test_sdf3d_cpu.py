"""CPU: the NumPy restatement of sdf3D.evaluate / project against fixtures produced by the UNMODIFIED reference module
under Numba's simulator (oracle/refsim/run_sdf3d.py)."""
import numpy as np

from conftest import load_golden
from oracle import numpy_oracle_sdf as S


def test_evaluate_vs_reference():
    f = load_golden("sdf3d_bodies")
    sd, vel = S.evaluate(f["rb_d"], f["pos"])
    assert np.max(np.abs(sd - f["sd"])) <= 1e-15
    assert np.array_equal(vel, f["vel"])
    assert (f["sd"] <= 0).any() and (np.abs(f["vel"]).sum(axis=1) > 0).any()       # the fixture exercises moving bodies
    gsd, gvel = S.evaluate(f["rb_d"], f["grid_pos"])
    assert gsd.shape == f["grid_sd"].shape and np.max(np.abs(gsd - f["grid_sd"])) <= 1e-15
    assert np.array_equal(gvel, f["grid_vel"])
    sd2, _ = S.evaluate(f["rb2"], f["pos2"])                                      # flipped sphere + cylinder (outside its end planes)
    assert np.max(np.abs(sd2 - f["sd2"])) <= 1e-15


def test_project_vs_reference():
    f = load_golden("sdf3d_bodies")
    p = S.project(f["rb_d"], f["pos"])
    assert np.max(np.abs(p - f["projected"])) <= 1e-15
    moved = np.abs(f["projected"] - f["pos"]).max(axis=1) > 1e-12
    assert moved.sum() > 100                                                      # outside the flipped container / inside the bodies
    sd_after, _ = S.evaluate(f["rb_d"], p)
    assert (sd_after > -1e-9).mean() > 0.9                                        # sequential projection leaves (almost) nothing inside
