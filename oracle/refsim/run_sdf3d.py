#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — golden fixtures for sdf3D.evaluate / sdf3D.project.

Executes the UNMODIFIED ``/root/reference/solver/sdf3D.py`` under Numba's CUDA simulator.  Two shims, both outside the
reference's source: (1) its helper functions (box_eval, inv_rigid, matvecmul4, ...) are decorated ``@cuda.jit`` without
``device=True``; real Numba-CUDA compiles such callees as device functions, the simulator treats them as kernels and refuses
to call them without a launch configuration.  ``numba.cuda.jit`` is therefore wrapped so that a decorated function called
WITHOUT a launch configuration runs as the plain function (what the GPU does), while ``kernel[blocks, threads](...)`` still
goes through the simulator.  (2) ``matplotlib`` (imported by the module, unused on this path, not installed) is stubbed.

``cylinder_eval`` reads ``y_clip`` before assignment whenever the point lies between the cylinder's end planes
(sdf3D.py:161-167: UnboundLocalError in Python, an uninitialised value on a GPU), so the cylinder in the fixtures is only
sampled above / below its end planes; the product implements the evident intent (y_clip = y, as cylinder_project does)
and is checked against the NumPy oracle there.

Never run on the GPU box.  Usage: python oracle/refsim/run_sdf3d.py
"""
import os
import sys
import types

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("FLUID_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, HERE)      # the cupy shim
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import cupy as cp  # noqa: E402
from numba import cuda  # noqa: E402

for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

_orig_jit = cuda.jit


class _DeviceCallable:
    """kernel[cfg](...) -> simulator launch; plain call -> the undecorated function (device-function semantics)"""

    def __init__(self, fn, kernel):
        self._fn, self._kernel = fn, kernel

    def __getitem__(self, cfg):
        return self._kernel[cfg]

    def __call__(self, *a, **k):
        return self._fn(*a, **k)


def _jit(fn=None, **kw):
    if fn is None:
        return lambda f: _jit(f, **kw)
    return _DeviceCallable(fn, _orig_jit(fn, **kw))


cuda.jit = _jit
if not hasattr(cp, "identity"):
    cp.identity = lambda n: cp.asarray(np.identity(n))
if not hasattr(cp, "append"):
    cp.append = lambda a, b, axis=None: cp.asarray(np.append(np.asarray(a), np.asarray(b), axis=axis))

from solver import sdf3D as sdf  # noqa: E402  (the reference module, unmodified)

GOLD = os.path.join(REPO, "tests", "golden")


def main():
    rng = np.random.default_rng(21)
    rb_d, rb_map = cp.asarray([]), {}
    # the notebook's five bodies (ipynb:682-689) ...
    rb_d, rb_map = sdf.generate_rb(rb_d, rb_map, "cube", ["box", 0.5, 0.8, 0.5], flip=True, center=[0, 0.5, 0], axis=[0, 1, 0], angle=0)
    rb_d, rb_map = sdf.generate_rb(rb_d, rb_map, "cube1", ["box", 0.67, 0.1, 1.0], flip=False, center=[-0.34, 0.7, 0], axis=[0, 0, 1], angle=-45)
    rb_d, rb_map = sdf.generate_rb(rb_d, rb_map, "cube2", ["box", 0.67, 0.1, 1.0], flip=False, center=[0.34, 0.7, 0], axis=[0, 0, 1], angle=45)
    rb_d, rb_map = sdf.generate_rb(rb_d, rb_map, "cube3", ["box", 1.0, 0.1, 0.7], flip=False, center=[0, 0.7, -0.3], axis=[1, 0, 0], angle=45)
    rb_d, rb_map = sdf.generate_rb(rb_d, rb_map, "cube4", ["box", 1.0, 0.1, 0.7], flip=False, center=[0, 0.7, 0.3], axis=[1, 0, 0], angle=-45)
    # ... plus a moving sphere and a small moving box inside the container
    rb_d, rb_map = sdf.generate_rb(rb_d, rb_map, "ball", ["sphere", 0.07], flip=False, center=[0.05, 0.3, -0.04])
    rb_d, rb_map = sdf.generate_rb(rb_d, rb_map, "brick", ["box", 0.1, 0.06, 0.08], flip=False, center=[-0.1, 0.25, 0.1], axis=[1, 1, 0], angle=30)
    sdf.set_vel_rb(rb_d, rb_map["ball"], cp.asarray([0.3, -0.2, 0.1]))
    sdf.set_vel_rb(rb_d, rb_map["brick"], cp.asarray([-0.5, 0.0, 0.25]))
    n = 3000
    pos = np.array([-0.35, -0.05, -0.35]) + rng.random((n, 3)) * np.array([0.7, 1.0, 0.7])
    pos[:200] = np.array([0.05, 0.3, -0.04]) + rng.normal(0, 0.05, (200, 3))          # around / inside the sphere
    pos[200:400] = np.array([-0.1, 0.25, 0.1]) + rng.normal(0, 0.05, (200, 3))        # around / inside the brick
    sd = cp.zeros(n)
    vel = cp.asarray(rng.normal(0, 1, (n, 3)))            # evaluate() zeroes it first
    sdf.evaluate(rb_d, sd, vel, cp.asarray(pos.copy()))
    proj = cp.asarray(pos.copy())
    sdf.project(rb_d, proj)
    out = dict(rb_d=np.array(rb_d), pos=pos, sd=np.array(sd), vel=np.array(vel), projected=np.array(proj))
    # a (3-D shaped) grid evaluation like the notebook's solid level set (ipynb:791)
    gp = np.stack(np.meshgrid(np.linspace(-0.3, 0.3, 9), np.linspace(0.0, 1.0, 11), np.linspace(-0.3, 0.3, 7), indexing="ij"), axis=-1)
    gsd, gvel = cp.zeros(gp.shape[:-1]), cp.zeros(gp.shape)
    sdf.evaluate(rb_d, gsd, gvel, cp.asarray(gp.copy()))
    out.update(grid_pos=gp, grid_sd=np.array(gsd), grid_vel=np.array(gvel))
    # flipped sphere + cylinder sampled outside its end planes only (see the module docstring)
    rb2, m2 = cp.asarray([]), {}
    rb2, m2 = sdf.generate_rb(rb2, m2, "dome", ["sphere", 0.4], flip=True, center=[0, 0.4, 0])
    rb2, m2 = sdf.generate_rb(rb2, m2, "can", ["cylinder", 0.1, 0.2], flip=False, center=[0.0, 0.3, 0.0], axis=[0, 0, 1], angle=20)
    p2 = np.array([-0.5, -0.1, -0.5]) + rng.random((1500, 3)) * np.array([1.0, 1.0, 1.0])
    R = np.array(rb2[1, 5:8, :3])
    local_y = (p2 - np.array(rb2[1, 1:4, 3])) @ R[:, 1]
    p2 = p2[np.abs(local_y) > 0.1 + 1e-9]
    sd2, v2 = cp.zeros(p2.shape[0]), cp.zeros(p2.shape)
    sdf.evaluate(rb2, sd2, v2, cp.asarray(p2.copy()))
    out.update(rb2=np.array(rb2), pos2=p2, sd2=np.array(sd2))
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, "sdf3d_bodies.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
