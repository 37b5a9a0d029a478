"""B200-native drop-in for the reference's ``solver/sdf3D.py``: rigid-body signed distance fields.

Same public surface — ``evaluate(rb_d, sd, vel, position)`` (:255-266), ``project(rb_d, position)`` (:268-273),
``get_T`` / ``get_R`` / ``generate_rb`` / ``transform_rb`` / ``set_vel_rb`` (:275-337) with the reference's body-table layout
(``n x 10 x 4`` fp64) — but the table lives in a torch CUDA tensor and the two kernels are hand-written sm_100a code behind
the C ABI (``fs_sdf3d_evaluate`` / ``fs_sdf3d_project``).  No CuPy / Numba / matplotlib, no CPU fallback.
"""
import numpy as np
import torch

from . import _arrays as A
from . import _native as N


def _table(rb_d):
    t = A.as_arg(rb_d, "rb_d", dtypes=(torch.float64,), want=torch.float64)
    if t.t.dim() != 3 or tuple(t.t.shape[1:]) != (10, 4):
        raise ValueError(f"rb_d must have shape (n, 10, 4), got {tuple(t.t.shape)}")
    return t


def evaluate(rb_d, sd, vel, position):
    """sd <- signed distance to the nearest body, vel <- that body's velocity where sd <= 0 (0 elsewhere); in place."""
    pos = A.as_arg(position, "position", dtypes=(torch.float64,), want=torch.float64)
    if pos.t.shape[-1] != 3:
        raise ValueError("position must have a trailing axis of 3")
    s = A.as_arg(sd, "sd", shape=tuple(pos.t.shape[:-1]), dtypes=(torch.float64,), want=torch.float64)
    v = A.as_arg(vel, "vel", dtypes=(torch.float64,), want=torch.float64)
    if v.t.shape[-1] != 3 or v.t.numel() != pos.t.numel():
        raise ValueError("vel must have the shape of position")
    tab = _table(rb_d)
    lib = N.load()
    N.check(lib.fs_sdf3d_evaluate(tab.ptr, int(tab.t.shape[0]), pos.t.numel() // 3, pos.ptr, s.ptr, v.ptr, A.stream_ptr()), "fs_sdf3d_evaluate")
    s.sync_back()
    v.sync_back()


def project(rb_d, position):
    """Move every position out of the solid bodies (into the flipped ones), body by body in table order; in place."""
    pos = A.as_arg(position, "position", dtypes=(torch.float64,), want=torch.float64)
    if pos.t.shape[-1] != 3:
        raise ValueError("position must have a trailing axis of 3")
    tab = _table(rb_d)
    N.check(N.load().fs_sdf3d_project(tab.ptr, int(tab.t.shape[0]), pos.t.numel() // 3, pos.ptr, A.stream_ptr()), "fs_sdf3d_project")
    pos.sync_back()


def get_T(position):
    t = torch.eye(4, dtype=torch.float64, device=A.device())
    t[0:3, 3] = torch.as_tensor(np.asarray(position, dtype=np.float64), device=A.device())
    return t


def get_R(axis, angle):
    r = torch.eye(4, dtype=torch.float64, device=A.device())
    if angle:
        from scipy.spatial.transform import Rotation as R
        axis = np.asarray(axis, dtype=np.float64)
        m = R.from_rotvec(axis / np.linalg.norm(axis) * angle * np.pi / 180).as_matrix()
        r[:3, :3] = torch.as_tensor(m, device=A.device())
    return r


def generate_rb(rb_d, rb_map, name, rbparam, flip=False, center=[0, 0, 0], axis=[0, 1, 0], angle=0):
    """Append a body (``['sphere', r]``, ``['box', sx, sy, sz]`` or ``['cylinder', r, h]``) to the table; returns (rb_d, rb_map)."""
    rb = torch.zeros((1, 10, 4), dtype=torch.float64, device=A.device())
    if rbparam[0] == "sphere":
        rb[:, 0, 0] = 1 if flip else 0
        rb[:, 0, 1] = rbparam[1]
    elif rbparam[0] == "box":
        rb[:, 0, 0] = 3 if flip else 2
        rb[:, 0, 1:] = torch.as_tensor(np.asarray(rbparam[1:], dtype=np.float64), device=A.device())
    elif rbparam[0] == "cylinder":
        rb[:, 0, 0] = 5 if flip else 4
        rb[:, 0, 1:3] = torch.as_tensor(np.asarray(rbparam[1:], dtype=np.float64), device=A.device())
    else:
        return rb_d
    rb[:, 1:5, :] = get_T(center)
    rb[:, 5:9, :] = get_R(axis, angle)
    n = 0 if rb_d is None else int(np.prod(tuple(rb_d.shape))) and int(rb_d.shape[0])
    rb_map[name] = n
    if n == 0:
        rb_d = rb
    else:
        rb_d = torch.cat([torch.as_tensor(rb_d, device=A.device()).to(torch.float64), rb], dim=0)
    return rb_d, rb_map


def transform_rb(rb_d, index, center=None, axis=None, angle=None):
    if center:
        rb_d[index, 1:5, :] = get_T(center)
    if axis and angle:
        rb_d[index, 5:9, :] = get_R(axis, angle)


def set_vel_rb(rb_d, index, vel):
    rb_d[index, -1, :3] = torch.as_tensor(np.asarray(vel, dtype=np.float64), device=rb_d.device)
