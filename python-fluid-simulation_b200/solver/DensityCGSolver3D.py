"""B200-native drop-in for the reference's ``solver/DensityCGSolver3D.py`` (SURVEY §8 "next" row f-1).

Particle-density (volume-conservation) solve: scatter particle mass / volume to the cells, fix the cell volumes,
build the RHS, solve the 7-point ghost-fluid system with plain CG on ``CGSolverBuffer``'s arrays and move the particles
by the resulting displacement field.  Module functions keep the reference's names and argument orders
(``initialize_density`` :250, ``fix_volume`` :257, ``initialize_solver`` :266, ``matvecmul`` :274,
``compute_displacement`` :280, ``apply_displacement`` :286); ``class DensityCGSolver3D(buf, gres, bound_min,
bound_size)`` / ``solve(rho0, dt, px, pm, pvol, vx, vy, vz, sphi, sv, lphi, lvol, wx=None, wy=None, wz=None,
tol=1e-3)`` keeps its surface (:283-350): ``px`` is moved in place, ``wx, wy, wz`` (the face open fractions the
notebook hands on to the pressure solver, ipynb:4648), ``m, vol, x, dx, dy, dz`` are attributes.

Everything runs in hand-written sm_100a kernels behind the C ABI (``fs_dens3d_*`` and ``fs_press_*`` with
``FS_OP_DENSITY``): the CG iterates on the active cell set with the persistent whole-iteration kernel.  fp64 state.
The particle scatter uses fp64 atomics like the reference, so ``m`` / ``vol`` are reproducible to summation-order
rounding; every other kernel is bit-exact in fp64.  No CPU fallback.
"""
import ctypes

import numpy as np
import torch

from . import _arrays as A
from . import _native as N
from . import _pressure as P
from .SolidFraction3D import compute_solid_frac, edge_in_fraction  # noqa: F401  (same re-exports as the reference :6)

_BIAS = ((0.0, 0.5, 0.5), (0.5, 0.0, 0.5), (0.5, 0.5, 0.0))          # :288-290


def _vec3(v):
    a = A.to_host_f64(v, 3)
    return (ctypes.c_double * 3)(*a)


def _g3(gres):
    g = A.to_host_ints(gres)
    if len(g) != 3:
        raise ValueError("DensityCGSolver3D needs a 3-entry gres")
    return g


def _cells(a, name, g, inout=False):
    return A.as_arg(a, name, shape=g, dtypes=(torch.float64,)) if inout else A.as_arg(a, name, shape=g, want=torch.float64)


def _faces(ws, g, inout=False, names=("wx", "wy", "wz")):
    shapes = P.mac_shapes(g)
    if inout:
        return [A.as_arg(w, n, shape=s, dtypes=(torch.float64,)) for w, n, s in zip(ws, names, shapes)]
    return [A.as_arg(w, n, shape=s, want=torch.float64) for w, n, s in zip(ws, names, shapes)]


def _particles(px, inout):
    a = A.as_arg(px, "px")
    if a.t.dim() != 2 or a.t.shape[1] != 3:
        raise ValueError(f"px: expected shape (P, 3), got {tuple(a.t.shape)}")
    return a


def initialize_density(bound_min, cell_size, gres, px, pm, pvol, gm, gvol, sphi=None, lphi=None):
    """Particle -> cell scatter of mass and volume (reference :250-255; ``sphi`` / ``lphi`` are accepted and unused there too)."""
    g = _g3(gres)
    p = _particles(px, False)
    m = A.as_arg(pm, "pm", shape=(p.t.shape[0],))
    mm, vv = _cells(gm, "gm", g, True), _cells(gvol, "gvol", g, True)
    lib = N.load()
    N.check(lib.fs_dens3d_scatter(*g, _vec3(bound_min), _vec3(cell_size), p.ptr, p.code, m.ptr, m.code, int(p.t.shape[0]), float(pvol),
                                  mm.ptr, vv.ptr, A.stream_ptr()), "fs_dens3d_scatter")
    mm.sync_back()
    vv.sync_back()


def fix_volume(cell_size, gres, lvol, gvol, sphi, lphi, wx, wy, wz):
    """Reference :257-264.  ``lvol`` is unused by the reference kernel (its use is commented out, :44-61)."""
    g = _g3(gres)
    vv = _cells(gvol, "gvol", g, True)
    s = A.as_arg(sphi, "sphi", shape=P.fine_shape(g), want=torch.float64)
    lp = _cells(lphi, "lphi", g)
    w = _faces((wx, wy, wz), g)
    lib = N.load()
    N.check(lib.fs_dens3d_fix_volume(*g, _vec3(cell_size), vv.ptr, s.ptr, lp.ptr, w[0].ptr, w[1].ptr, w[2].ptr, A.stream_ptr()), "fs_dens3d_fix_volume")
    vv.sync_back()


def initialize_solver(rho0, dt, gres, cell_size, gm, gvol, lphi, wx, wy, wz, b):
    """RHS b = (1 - clamp(density / rho0, 0.5, 1.5)) / dt on interior fluid cells (reference :266-272)."""
    g = _g3(gres)
    mm, vv, lp = _cells(gm, "gm", g), _cells(gvol, "gvol", g), _cells(lphi, "lphi", g)
    w = _faces((wx, wy, wz), g)
    bb = _cells(b, "b", g, True)
    lib = N.load()
    N.check(lib.fs_dens3d_rhs(*g, float(rho0), float(dt), _vec3(cell_size), mm.ptr, vv.ptr, lp.ptr, w[0].ptr, w[1].ptr, w[2].ptr, bb.ptr,
                              A.stream_ptr()), "fs_dens3d_rhs")
    bb.sync_back()


def matvecmul(gres, v, out, wx, wy, wz, lphi):
    """out = A v with the density operator (reference :274-278)."""
    g = _g3(gres)
    e = P.engine(g, "density")
    vv = _cells(v, "v", g)
    oo = _cells(out, "out", g, True)
    w = _faces((wx, wy, wz), g)
    lp = _cells(lphi, "lphi", g)
    N.check(e.lib.fs_press_apply(e.h, vv.ptr, oo.ptr, w[0].ptr, w[1].ptr, w[2].ptr, lp.ptr, A.stream_ptr()), "fs_press_apply")
    oo.sync_back()


def compute_displacement(gres, dt, cell_size, dx, dy, dz, pv, lphi):
    """Face displacements from the solved potential, indices 1..g-1 on all axes (reference :280-284)."""
    g = _g3(gres)
    d = _faces((dx, dy, dz), g, True, ("dx", "dy", "dz"))
    p, lp = _cells(pv, "pv", g), _cells(lphi, "lphi", g)
    lib = N.load()
    N.check(lib.fs_dens3d_displacement(*g, float(dt), _vec3(cell_size), d[0].ptr, d[1].ptr, d[2].ptr, p.ptr, lp.ptr, A.stream_ptr()),
            "fs_dens3d_displacement")
    for a in d:
        a.sync_back()


def apply_displacement(px, dx, bound_min, cell_size, grid_bias, axis):
    """Trilinear gather of one displacement component onto the particles, in place on ``px[:, axis]`` (reference :286-291)."""
    p = _particles(px, True)
    d = A.as_arg(dx, "dx", want=torch.float64)
    if d.t.dim() != 3:
        raise ValueError("dx must be a 3-D array")
    lib = N.load()
    N.check(lib.fs_dens3d_gather(p.ptr, p.code, int(p.t.shape[0]), d.ptr, *[int(n) for n in d.t.shape], _vec3(bound_min), _vec3(cell_size),
                                 _vec3(grid_bias), int(axis), A.stream_ptr()), "fs_dens3d_gather")
    p.sync_back()


class DensityCGSolver3D:
    """Reference :283-350."""

    def __init__(self, buf, gres, bound_min, bound_size):
        self.gres = gres
        self._g = _g3(gres)
        self.bound_min = bound_min
        self._bmin = A.to_host_f64(bound_min, 3)
        self.cell_size = A.to_host_f64(bound_size, 3) / np.asarray(self._g, dtype=np.float64)           # :297
        self.buf = buf
        dev = A.device()
        g = self._g
        self.m = torch.zeros(g, dtype=torch.float64, device=dev)
        self.vol = torch.zeros(g, dtype=torch.float64, device=dev)
        self.x = torch.zeros(g, dtype=torch.float64, device=dev)
        for a, s in enumerate(P.mac_shapes(g)):
            setattr(self, "w" + "xyz"[a], torch.zeros(s, dtype=torch.float64, device=dev))
            setattr(self, "d" + "xyz"[a], torch.zeros(s, dtype=torch.float64, device=dev))
        self.alpha = 0.0
        self.beta = 0.0
        self.delta = 0.0
        self.iterations = 0
        self.max_iter = int(np.prod(np.asarray(g, dtype=np.int64)))                                    # :318
        self._e = P.Engine(g, "density")

    def solve(self, rho0, dt, px, pm, pvol, vx, vy, vz, sphi, sv, lphi, lvol, wx=None, wy=None, wz=None, tol=1e-3):
        """In-place particle displacement towards rest density (reference :320-350).  ``vx, vy, vz, sv, lvol`` are
        accepted and unused, exactly as in the reference.  Raises ``ValueError("Failed to converge!")`` like its
        ``for ... else``."""
        g, e, lib = self._g, self._e, self._e.lib
        if wx is None or wy is None or wz is None:                                                      # :321-325
            compute_solid_frac(g, sphi, self.wx, self.wy, self.wz)
            wx, wy, wz = self.wx, self.wy, self.wz
        stream = A.stream_ptr()
        p = _particles(px, True)
        m = A.as_arg(pm, "pm", shape=(p.t.shape[0],))
        s = A.as_arg(sphi, "sphi", shape=P.fine_shape(g), want=torch.float64)
        lp = _cells(lphi, "lphi", g)
        w = _faces((wx, wy, wz), g)
        bufs = {k: A.as_arg(getattr(self.buf, k), "buf." + k, shape=g, dtypes=(torch.float64,)) for k in "drqb"}
        cs, bmin = _vec3(self.cell_size), _vec3(self._bmin)
        npart = int(p.t.shape[0])
        self.m.zero_()                                                                                 # :326-328 (x is zeroed by the CG)
        self.vol.zero_()
        N.check(lib.fs_dens3d_scatter(*g, bmin, cs, p.ptr, p.code, m.ptr, m.code, npart, float(pvol), self.m.data_ptr(), self.vol.data_ptr(), stream),
                "fs_dens3d_scatter")
        N.check(lib.fs_dens3d_fix_volume(*g, cs, self.vol.data_ptr(), s.ptr, lp.ptr, w[0].ptr, w[1].ptr, w[2].ptr, stream), "fs_dens3d_fix_volume")
        N.check(lib.fs_dens3d_rhs(*g, float(rho0), float(dt), cs, self.m.data_ptr(), self.vol.data_ptr(), lp.ptr, w[0].ptr, w[1].ptr, w[2].ptr,
                                  bufs["b"].ptr, stream), "fs_dens3d_rhs")
        st = N.CgStats()
        status = N.check(lib.fs_press_cg(e.h, self.x.data_ptr(), bufs["d"].ptr, bufs["r"].ptr, bufs["q"].ptr, bufs["b"].ptr,
                                         w[0].ptr, w[1].ptr, w[2].ptr, lp.ptr, float(tol), int(self.max_iter), ctypes.byref(st), stream),
                         "fs_press_cg")
        self.delta, self.alpha, self.beta, self.iterations = st.delta, st.alpha, st.beta, int(st.iterations)
        for b in bufs.values():
            b.sync_back()
        if status == N.FS_NOT_CONVERGED:
            raise ValueError("Failed to converge!")                                                    # :342-343
        N.check(lib.fs_dens3d_displacement(*g, float(dt), cs, self.dx.data_ptr(), self.dy.data_ptr(), self.dz.data_ptr(), self.x.data_ptr(),
                                           lp.ptr, stream), "fs_dens3d_displacement")                  # :345
        for axis, d in enumerate((self.dx, self.dy, self.dz)):                                          # :346-348
            N.check(lib.fs_dens3d_gather(p.ptr, p.code, npart, d.data_ptr(), *[int(n) for n in d.shape], bmin, cs, _vec3(_BIAS[axis]), axis, stream),
                    "fs_dens3d_gather")
        torch.cuda.current_stream().synchronize()
        p.sync_back()
