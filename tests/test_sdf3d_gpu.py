"""GPU parity of solver/sdf3D.py (fs_sdf3d_* kernels) against fixtures produced by the unmodified reference module, and
against the NumPy oracle for the cylinder's repaired between-the-caps case."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def test_evaluate_and_project_vs_reference():
    from solver import sdf3D as sdf
    f = load_golden("sdf3d_bodies")
    rb = dev(f["rb_d"])
    n = f["pos"].shape[0]
    sd = torch.full((n,), 7.0, dtype=torch.float64, device="cuda")
    vel = torch.full((n, 3), 7.0, dtype=torch.float64, device="cuda")
    sdf.evaluate(rb, sd, vel, dev(f["pos"]))
    assert np.max(np.abs(sd.cpu().numpy() - f["sd"])) <= 1e-15
    assert np.array_equal(vel.cpu().numpy(), f["vel"])                      # zero outside, the nearest body's velocity inside
    gsd = torch.zeros(f["grid_sd"].shape, dtype=torch.float64, device="cuda")
    gvel = torch.zeros(f["grid_vel"].shape, dtype=torch.float64, device="cuda")
    sdf.evaluate(rb, gsd, gvel, dev(f["grid_pos"]))                         # grid-shaped call, like the notebook's solid level set
    assert np.max(np.abs(gsd.cpu().numpy() - f["grid_sd"])) <= 1e-15
    assert np.array_equal(gvel.cpu().numpy(), f["grid_vel"])
    pos = dev(f["pos"])
    sdf.project(rb, pos)
    assert np.max(np.abs(pos.cpu().numpy() - f["projected"])) <= 1e-15
    sd2 = torch.zeros(f["pos2"].shape[0], dtype=torch.float64, device="cuda")
    sdf.evaluate(dev(f["rb2"]), sd2, torch.zeros(f["pos2"].shape, dtype=torch.float64, device="cuda"), dev(f["pos2"]))
    assert np.max(np.abs(sd2.cpu().numpy() - f["sd2"])) <= 1e-15


def test_generate_rb_and_cylinder_vs_oracle():
    from oracle import numpy_oracle_sdf as S
    from solver import sdf3D as sdf
    f = load_golden("sdf3d_bodies")
    rb, m = None, {}
    rb, m = sdf.generate_rb(rb, m, "cube", ["box", 0.5, 0.8, 0.5], flip=True, center=[0, 0.5, 0], axis=[0, 1, 0], angle=0)
    rb, m = sdf.generate_rb(rb, m, "cube1", ["box", 0.67, 0.1, 1.0], flip=False, center=[-0.34, 0.7, 0], axis=[0, 0, 1], angle=-45)
    assert m == {"cube": 0, "cube1": 1}
    assert np.allclose(rb.cpu().numpy(), f["rb_d"][:2], atol=1e-15)        # same table as the reference's generate_rb
    rb, m = sdf.generate_rb(rb, m, "can", ["cylinder", 0.1, 0.25], flip=False, center=[0.0, 0.3, 0.05], axis=[1, 0, 1], angle=35)
    rb, m = sdf.generate_rb(rb, m, "pipe", ["cylinder", 0.3, 0.9], flip=True, center=[0.0, 0.5, 0.0], axis=[0, 0, 1], angle=5)
    sdf.set_vel_rb(rb, m["can"], [0.1, 0.2, 0.3])
    rng = np.random.default_rng(3)
    pos = np.array([-0.4, -0.1, -0.4]) + rng.random((20000, 3)) * np.array([0.8, 1.2, 0.8])
    sd = torch.zeros(pos.shape[0], dtype=torch.float64, device="cuda")
    vel = torch.zeros(pos.shape, dtype=torch.float64, device="cuda")
    sdf.evaluate(rb, sd, vel, dev(pos))
    osd, ovel = S.evaluate(rb.cpu().numpy(), pos)
    assert np.max(np.abs(sd.cpu().numpy() - osd)) <= 1e-14
    assert np.array_equal(vel.cpu().numpy(), ovel)
    p = dev(pos)
    sdf.project(rb, p)
    assert np.max(np.abs(p.cpu().numpy() - S.project(rb.cpu().numpy(), pos))) <= 1e-13
    host = pos.copy()                                                      # host arrays are updated in place too
    sdf.project(rb, host)
    assert np.array_equal(host, p.cpu().numpy())
