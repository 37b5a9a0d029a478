"""B200-native drop-ins for the grid-side functions the reference notebook defines inline
(``3D_viscous_fluid_sim.ipynb`` code cells 2-7): ``p2g``, ``g2p``, ``compute_fluid_levelset``, ``compute_fluid_volume``,
``extrapolate`` and ``apply_boundary_condition`` — same names, same argument lists (the notebook's ``edict`` objects: any
object with the same attributes works), same in-place mutation.  Replacing those cells by

    from notebook_kernels import p2g, g2p, compute_fluid_levelset, compute_fluid_volume, extrapolate, apply_boundary_condition

keeps ``lvol`` / ``lphi`` / the MAC velocities on the device between the particle steps and the implicit solves, which is
what removes the per-step PCIe upload of the solver inputs from a real time step.  Hand-written sm_100a kernels behind the
C ABI (``fs_grid_*`` in ``include/fluidsolver_b200.h``); no CuPy / Numba, no CPU fallback.

Array types are the notebook's (ipynb cell 10): particle arrays fp64, MAC grids fp32, level-set / volume grids fp64; other
dtypes are converted on the way in and written back.  One deliberate difference: the reference's boundary-condition kernels
write ``dv[x,y,z] = 0`` before their bounds test, i.e. out of bounds for every thread of a launch that overhangs the array;
here only the array is written.
"""
import ctypes
import math

import numpy as np
import torch

from solver import _arrays as A
from solver import _native as N

_D3 = ctypes.c_double * 3


def _vec3(v):
    a = A.to_host_f64(v, 3)
    return _D3(*[float(x) for x in a])


def _get(obj, name):
    return obj[name] if isinstance(obj, dict) else getattr(obj, name)


def _arr(a, name, dtype, shape=None):
    return A.as_arg(a, name, shape=shape, dtypes=(dtype,), want=dtype)


def _mac_shapes(g):
    return [tuple(n + (1 if i == a else 0) for i, n in enumerate(g)) for a in range(3)]


def _particles(p):
    n = int(_get(p, "num_particles"))
    x = _arr(_get(p, "x"), "p.x", torch.float64, (n, 3))
    return n, x


def p2g(p, g):
    """Particles -> MAC grids (ipynb :279-344).  ``g.{x,y,z}.m`` / ``.v`` must be zeroed by the caller, as in the notebook."""
    lib = N.load()
    gres = A.to_host_ints(_get(g, "resolution"))
    n, x = _particles(p)
    pm = _arr(_get(p, "m"), "p.m", torch.float64, (n,))
    pv = _arr(_get(p, "v"), "p.v", torch.float64, (n, 3))
    c = [_arr(_get(p, k), "p." + k, torch.float64, (n, 3)) for k in ("cx", "cy", "cz")]
    comps = [_get(g, a) for a in "xyz"]
    m = [_arr(_get(cc, "m"), f"g.{a}.m", torch.float32, s) for cc, a, s in zip(comps, "xyz", _mac_shapes(gres))]
    v = [_arr(_get(cc, "v"), f"g.{a}.v", torch.float32, s) for cc, a, s in zip(comps, "xyz", _mac_shapes(gres))]
    N.check(lib.fs_grid_p2g(*gres, _vec3(_get(g, "bound_min")), _vec3(_get(g, "cell_size")), n, x.ptr, pm.ptr, pv.ptr, c[0].ptr, c[1].ptr, c[2].ptr,
                            m[0].ptr, v[0].ptr, m[1].ptr, v[1].ptr, m[2].ptr, v[2].ptr, A.stream_ptr()), "fs_grid_p2g")
    for a in m + v:
        a.sync_back()


def g2p(p, g):
    """MAC grids -> particle velocity and affine matrices (ipynb :352-393)."""
    lib = N.load()
    gres = A.to_host_ints(_get(g, "resolution"))
    n, x = _particles(p)
    pv = _arr(_get(p, "v"), "p.v", torch.float64, (n, 3))
    c = [_arr(_get(p, k), "p." + k, torch.float64, (n, 3)) for k in ("cx", "cy", "cz")]
    v = [_arr(_get(_get(g, a), "v"), f"g.{a}.v", torch.float32, s) for a, s in zip("xyz", _mac_shapes(gres))]
    N.check(lib.fs_grid_g2p(*gres, _vec3(_get(g, "bound_min")), _vec3(_get(g, "cell_size")), n, x.ptr, pv.ptr, c[0].ptr, c[1].ptr, c[2].ptr,
                            v[0].ptr, v[1].ptr, v[2].ptr, A.stream_ptr()), "fs_grid_g2p")
    for a in [pv] + c:
        a.sync_back()


def compute_fluid_levelset(p, ls, gdx):
    """Fluid signed distance on the cell grid (ipynb :94-136): phi = 3*gdx, then the union of particle spheres."""
    lib = N.load()
    gres = A.to_host_ints(_get(ls, "resolution"))
    n, x = _particles(p)
    phi = _arr(_get(ls, "phi"), "ls.phi", torch.float64, tuple(gres))
    r = gdx * 0.5 * math.sqrt(3.0) * 1.02
    N.check(lib.fs_grid_levelset(*gres, _vec3(_get(ls, "bound_min")), _vec3(_get(ls, "cell_size")), n, x.ptr, float(r), float(gdx * 3), phi.ptr,
                                 A.stream_ptr()), "fs_grid_levelset")
    phi.sync_back()


def compute_fluid_volume(p, fv, pvol):
    """Liquid volume per fine-grid node (ipynb :224-268): trilinear splat of the particle volume, clamped to the node volume."""
    lib = N.load()
    res = A.to_host_ints(_get(fv, "resolution"))
    n, x = _particles(p)
    vol = _arr(_get(fv, "vol"), "fv.vol", torch.float64, tuple(res))
    cs = A.to_host_f64(_get(fv, "cell_size"), 3)
    N.check(lib.fs_grid_fluid_volume(*res, _vec3(_get(fv, "bound_min")), _vec3(cs), n, x.ptr, float(pvol), float(np.prod(cs)), vol.ptr, A.stream_ptr()),
            "fs_grid_fluid_volume")
    vol.sync_back()


def extrapolate(gres, num_iter, vx, vy, vz, mx, my, mz):
    """Extrapolate the MAC velocities into faces without mass (ipynb :501-557), in place."""
    lib = N.load()
    g = A.to_host_ints(gres)
    v = [_arr(a, n, torch.float32, s) for a, n, s in zip((vx, vy, vz), ("vx", "vy", "vz"), _mac_shapes(g))]
    m = [_arr(a, n, torch.float32, s) for a, n, s in zip((mx, my, mz), ("mx", "my", "mz"), _mac_shapes(g))]
    nbytes = lib.fs_grid_extrapolate_workspace_bytes(*g)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=A.device())
    N.check(lib.fs_grid_extrapolate(*g, int(num_iter), v[0].ptr, v[1].ptr, v[2].ptr, m[0].ptr, m[1].ptr, m[2].ptr, ws.data_ptr(), nbytes, A.stream_ptr()),
            "fs_grid_extrapolate")
    for a in v:
        a.sync_back()


def apply_boundary_condition(g, solid, dx):
    """Remove the velocity component pointing into the solid near its surface (ipynb cell 5); fills ``g.*.dv``, updates ``g.*.v``."""
    lib = N.load()
    gres = A.to_host_ints(_get(g, "resolution"))
    comps = [_get(g, a) for a in "xyz"]
    sh = _mac_shapes(gres)
    v = [_arr(_get(c, "v"), f"g.{a}.v", torch.float32, s) for c, a, s in zip(comps, "xyz", sh)]
    m = [_arr(_get(c, "m"), f"g.{a}.m", torch.float32, s) for c, a, s in zip(comps, "xyz", sh)]
    dv = [_arr(_get(c, "dv"), f"g.{a}.dv", torch.float32, s) for c, a, s in zip(comps, "xyz", sh)]
    fine = tuple(2 * n + 1 for n in gres)
    sphi = _arr(_get(solid, "phi"), "solid.phi", torch.float64, fine)
    sv = _arr(_get(solid, "v"), "solid.v", torch.float64, fine + (3,))
    N.check(lib.fs_grid_boundary(*gres, float(dx), v[0].ptr, v[1].ptr, v[2].ptr, m[0].ptr, m[1].ptr, m[2].ptr, sphi.ptr, sv.ptr,
                                 dv[0].ptr, dv[1].ptr, dv[2].ptr, A.stream_ptr()), "fs_grid_boundary")
    for a in v + dv:
        a.sync_back()
