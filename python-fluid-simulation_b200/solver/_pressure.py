"""Dimension-generic host side of the pressure solvers (PressureCGSolver3D.py / PressureCGSolver2D.py)."""
import ctypes

import numpy as np
import torch

from . import _arrays as A
from . import _native as N


def mac_shapes(g):
    return [tuple(n + (1 if i == a else 0) for i, n in enumerate(g)) for a in range(len(g))]


def fine_shape(g):
    return tuple(2 * n + 1 for n in g)


class Engine:
    """One native fs_press handle (+ its small device workspace) per grid resolution."""

    def __init__(self, g, op="pressure"):
        self.lib = N.load()
        self.g = tuple(g)
        self.op = op
        nz = self.g[2] if len(self.g) == 3 else 0
        self.dims = (self.g[0], self.g[1], nz)
        nbytes = self.lib.fs_press_workspace_bytes(*self.dims)
        if nbytes == 0:
            raise ValueError(f"invalid grid resolution {self.g}")
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=A.device())
        h = ctypes.c_void_p()
        N.check(self.lib.fs_press_create(ctypes.byref(h), *self.dims, self.ws.data_ptr(), nbytes), "fs_press_create")
        self.h = h
        if op != "pressure":       # the density solver shares the CG machinery with a different operator (fs_press_set_operator)
            N.check(self.lib.fs_press_set_operator(h, {"density": N.OP_DENSITY}[op]), "fs_press_set_operator")

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.fs_press_destroy(self.h)
                self.h = None
        except Exception:
            pass


_engines = {}


def engine(g, op="pressure"):
    key = (tuple(g), op, torch.cuda.current_device())
    e = _engines.get(key)
    if e is None:
        if len(_engines) > 8:
            _engines.clear()
        e = _engines[key] = Engine(g, op)
    return e


def _cs3(cell_size, d):
    cs = A.to_host_f64(cell_size, d)
    arr = (ctypes.c_double * 3)(*(list(cs) + [1.0] * (3 - d)))
    return arr


def _p(a):
    return a.ptr if a is not None else None


def _vel_args(g, vel, names):
    v = [A.as_arg(a, n, shape=s) for a, n, s in zip(vel, names, mac_shapes(g))]
    if len({a.code for a in v}) != 1:
        raise TypeError("velocity components must share one dtype")
    return v


def _w_args(g, ws):
    return [A.as_arg(w, f"w{'xyz'[a]}", shape=s, want=torch.float64) for a, (w, s) in enumerate(zip(ws, mac_shapes(g)))]


def _pad3(lst):
    return list(lst) + [None] * (3 - len(lst))


def initialize_solver(cell_size, gres, vel, sphi, sv, lphi, b, ws):
    g = A.to_host_ints(gres)
    d = len(g)
    e = engine(g)
    v = _vel_args(g, vel, ["vx", "vy", "vz"][:d])
    w = _w_args(g, ws)
    s = A.as_arg(sv, "sv", shape=fine_shape(g) + (d,), want=torch.float64)
    lp = A.as_arg(lphi, "lphi", shape=g, want=torch.float64)
    bb = A.as_arg(b, "b", shape=g, dtypes=(torch.float64,))
    vv, ww = _pad3(v), _pad3(w)
    N.check(e.lib.fs_press_rhs(e.h, _cs3(cell_size, d), _p(vv[0]), _p(vv[1]), _p(vv[2]), v[0].code, s.ptr, lp.ptr, bb.ptr,
                               _p(ww[0]), _p(ww[1]), _p(ww[2]), A.stream_ptr()), "fs_press_rhs")
    bb.sync_back()


def matvecmul(gres, v, out, ws, lphi):
    g = A.to_host_ints(gres)
    e = engine(g)
    vv = A.as_arg(v, "v", shape=g, want=torch.float64)
    oo = A.as_arg(out, "out", shape=g, dtypes=(torch.float64,))
    w = _pad3(_w_args(g, ws))
    lp = A.as_arg(lphi, "lphi", shape=g, want=torch.float64)
    N.check(e.lib.fs_press_apply(e.h, vv.ptr, oo.ptr, _p(w[0]), _p(w[1]), _p(w[2]), lp.ptr, A.stream_ptr()), "fs_press_apply")
    oo.sync_back()


def apply_pressure(gres, cell_size, vel, pv, ws, sv, lphi):
    g = A.to_host_ints(gres)
    d = len(g)
    e = engine(g)
    v = _vel_args(g, vel, ["vx", "vy", "vz"][:d])
    w = _w_args(g, ws)
    p = A.as_arg(pv, "pv", shape=g, want=torch.float64)
    s = A.as_arg(sv, "sv", shape=fine_shape(g) + (d,), want=torch.float64)
    lp = A.as_arg(lphi, "lphi", shape=g, want=torch.float64)
    vv, ww = _pad3(v), _pad3(w)
    N.check(e.lib.fs_press_update(e.h, _cs3(cell_size, d), _p(vv[0]), _p(vv[1]), _p(vv[2]), v[0].code, p.ptr,
                                  _p(ww[0]), _p(ww[1]), _p(ww[2]), s.ptr, lp.ptr, A.stream_ptr()), "fs_press_update")
    for a in v:
        a.sync_back()


class PressureSolverBase:
    """Shared implementation of PressureCGSolver3D (:173-226) and PressureCGSolver2D (:140-179)."""

    _dim = 3
    _raise_on_fail = True

    def __init__(self, buf, gres, bound_size):
        self.gres = gres
        self._g = A.to_host_ints(gres)
        if len(self._g) != self._dim:
            raise ValueError(f"{type(self).__name__} needs a {self._dim}-entry gres")
        # bound_size may be the scalar GDX, as the notebook passes it (ipynb:778): cell_size = GDX / gres per axis
        self.cell_size = A.to_host_f64(bound_size, self._dim) / np.asarray(self._g, dtype=np.float64)
        self.buf = buf
        dev = A.device()
        self.x = torch.zeros(self._g, dtype=torch.float64, device=dev)
        for a, s in enumerate(mac_shapes(self._g)):
            setattr(self, "w" + "xyz"[a], torch.zeros(s, dtype=torch.float64, device=dev))
        self.alpha = 0.0
        self.beta = 0.0
        self.delta = 0.0
        self.iterations = 0
        self.max_iter = int(np.prod(np.asarray(self._g, dtype=np.int64)))
        self._e = Engine(self._g)

    def _own_weights(self):
        return [getattr(self, "w" + "xyz"[a]) for a in range(self._dim)]

    def _solve(self, vel, sphi, sv, lphi, ws, tol):
        g, d, e = self._g, self._dim, self._e
        if ws is None:
            if d == 3:
                from .SolidFraction3D import compute_solid_frac
            else:
                from .SolidFraction2D import compute_solid_frac
            compute_solid_frac(g, sphi, *self._own_weights())
            ws = self._own_weights()
        v = _vel_args(g, vel, ["vx", "vy", "vz"][:d])
        w = _w_args(g, ws)
        s = A.as_arg(sv, "sv", shape=fine_shape(g) + (d,), want=torch.float64)
        lp = A.as_arg(lphi, "lphi", shape=g, want=torch.float64)
        bufs = {k: A.as_arg(getattr(self.buf, k), "buf." + k, shape=g, dtypes=(torch.float64,)) for k in "drqb"}
        vv, ww = _pad3(v), _pad3(w)
        cs = _cs3(self.cell_size, d)
        stream = A.stream_ptr()
        N.check(e.lib.fs_press_rhs(e.h, cs, _p(vv[0]), _p(vv[1]), _p(vv[2]), v[0].code, s.ptr, lp.ptr, bufs["b"].ptr,
                                   _p(ww[0]), _p(ww[1]), _p(ww[2]), stream), "fs_press_rhs")
        st = N.CgStats()
        status = N.check(e.lib.fs_press_cg(e.h, self.x.data_ptr(), bufs["d"].ptr, bufs["r"].ptr, bufs["q"].ptr, bufs["b"].ptr,
                                           _p(ww[0]), _p(ww[1]), _p(ww[2]), lp.ptr, float(tol), int(self.max_iter),
                                           ctypes.byref(st), stream), "fs_press_cg")
        self.delta, self.alpha, self.beta, self.iterations = st.delta, st.alpha, st.beta, int(st.iterations)
        for b in bufs.values():
            b.sync_back()
        if status == N.FS_NOT_CONVERGED and self._raise_on_fail:
            raise ValueError("Failed to converge!")
        N.check(e.lib.fs_press_update(e.h, cs, _p(vv[0]), _p(vv[1]), _p(vv[2]), v[0].code, self.x.data_ptr(),
                                      _p(ww[0]), _p(ww[1]), _p(ww[2]), s.ptr, lp.ptr, stream), "fs_press_update")
        torch.cuda.current_stream().synchronize()
        for a in v:
            a.sync_back()
