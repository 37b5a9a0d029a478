"""B200-native drop-in for the reference's ``solver/ViscosityCGSolver2D.py``.

Module functions ``initialize_solver``, ``matvecmul``, ``apply_viscosity`` keep the reference's argument order
(:222, :231, :240); ``class ViscosityCGSolver2D(gres, bound_size)`` / ``solve(dt, mu, rho, vx, vy, sphi, sv,
lphi, lvol, tol=1e-4, save=False)`` keeps its surface (:246-318).  Differences from 3-D that are kept: solid
test ``sphi <= 0`` (fluid iff ``> 0``), no extrapolation step, default ``tol=1e-4``, unused ``save`` kwarg,
``vol = lvol / (cell_vol * 0.125)`` with the 3-D factor 0.125 (:278)."""
import ctypes

import numpy as np
import torch

from . import _arrays as A
from . import _native as N

_DT = {torch.float32: N.FS_F32, torch.float64: N.FS_F64, "float32": N.FS_F32, "float64": N.FS_F64,
       np.float32: N.FS_F32, np.float64: N.FS_F64}
_TORCH = {N.FS_F32: torch.float32, N.FS_F64: torch.float64}


def _mac_shapes(g):
    return [(g[0] + 1, g[1]), (g[0], g[1] + 1)]


def _fine_shape(g):
    return (2 * g[0] + 1, 2 * g[1] + 1)


class _Engine:
    def __init__(self, g, code):
        self.lib = N.load()
        self.g = tuple(g)
        self.code = code
        self.tdtype = _TORCH[code]
        nbytes = self.lib.fs_visc2d_workspace_bytes(*self.g, code)
        if nbytes == 0:
            raise ValueError(f"invalid grid resolution {self.g}")
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=A.device())
        h = ctypes.c_void_p()
        N.check(self.lib.fs_visc2d_create(ctypes.byref(h), *self.g, code, self.ws.data_ptr(), nbytes), "fs_visc2d_create")
        self.h = h
        X, Yp, NL = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
        N.check(self.lib.fs_visc2d_lattice(h, X, Yp, NL), "fs_visc2d_lattice")
        self.lat = (X.value, Yp.value)
        self.NL = NL.value

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.fs_visc2d_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def view(self, vec, comp):
        p = self.lib.fs_visc2d_vector_ptr(self.h, vec, comp)
        esz = 4 if self.code == N.FS_F32 else 8
        off = p - self.ws.data_ptr()
        flat = self.ws[off: off + self.NL * esz].view(self.tdtype)
        sh = _mac_shapes(self.g)[comp]
        return flat.view(*self.lat)[: sh[0], : sh[1]]

    def pack(self, sphi, lvol, vol_norm):
        N.check(self.lib.fs_visc2d_pack(self.h, sphi.ptr, lvol.ptr, float(vol_norm), A.stream_ptr()), "fs_visc2d_pack")

    def load(self, vec, v):
        if v[0].code != v[1].code:
            raise TypeError("velocity components must share one dtype")
        N.check(self.lib.fs_visc2d_load(self.h, vec, v[0].ptr, v[1].ptr, v[0].code, A.stream_ptr()), "fs_visc2d_load")

    def store(self, vec, v, mode):
        if v[0].code != v[1].code:
            raise TypeError("velocity components must share one dtype")
        N.check(self.lib.fs_visc2d_store(self.h, vec, v[0].ptr, v[1].ptr, v[0].code, mode, A.stream_ptr()), "fs_visc2d_store")
        for a in v:
            a.sync_back()


_engines = {}


def _engine(g, code):
    key = (tuple(g), code, torch.cuda.current_device())
    e = _engines.get(key)
    if e is None:
        if len(_engines) > 4:
            _engines.clear()
        e = _engines[key] = _Engine(g, code)
    return e


def _mac_args(g, arrs, names):
    return [A.as_arg(a, n, shape=s) for a, n, s in zip(arrs, names, _mac_shapes(g))]


def _code(dtype):
    return N.FS_F64 if dtype is None else _DT[dtype]


def initialize_solver(gres, scale, mu, vx, vy, sphi, sv, vol, b_x, b_y, dtype=None):
    g = A.to_host_ints(gres)
    e = _engine(g, _code(dtype))
    v = _mac_args(g, (vx, vy), ("vx", "vy"))
    b = _mac_args(g, (b_x, b_y), ("b_x", "b_y"))
    s = A.as_arg(sphi, "sphi", shape=_fine_shape(g), want=torch.float64)
    vl = A.as_arg(vol, "vol", shape=_fine_shape(g), want=torch.float64)
    e.pack(s, vl, 1.0)
    e.load(N.VEC_X, v)
    N.check(e.lib.fs_visc2d_rhs(e.h, float(scale), float(mu), N.VEC_X, N.VEC_B, A.stream_ptr()), "fs_visc2d_rhs")
    e.store(N.VEC_B, b, N.STORE_INTERIOR)


def matvecmul(gres, scale, mu, vx, vy, out_x, out_y, sphi, vol, dtype=None):
    g = A.to_host_ints(gres)
    e = _engine(g, _code(dtype))
    v = _mac_args(g, (vx, vy), ("vx", "vy"))
    o = _mac_args(g, (out_x, out_y), ("out_x", "out_y"))
    s = A.as_arg(sphi, "sphi", shape=_fine_shape(g), want=torch.float64)
    vl = A.as_arg(vol, "vol", shape=_fine_shape(g), want=torch.float64)
    e.pack(s, vl, 1.0)
    e.load(N.VEC_D, v)
    N.check(e.lib.fs_visc2d_apply(e.h, float(scale), float(mu), N.VEC_D, N.VEC_Q, A.stream_ptr()), "fs_visc2d_apply")
    e.store(N.VEC_Q, o, N.STORE_INTERIOR)


def apply_viscosity(gres, vx, vy, out_x, out_y, sphi, sv, dtype=None):
    g = A.to_host_ints(gres)
    e = _engine(g, _code(dtype))
    v = _mac_args(g, (vx, vy), ("vx", "vy"))
    o = _mac_args(g, (out_x, out_y), ("out_x", "out_y"))
    s = A.as_arg(sphi, "sphi", shape=_fine_shape(g), want=torch.float64)
    e.pack(s, s, 1.0)
    e.load(N.VEC_X, o)
    e.store(N.VEC_X, v, N.STORE_FLUID)


class ViscosityCGSolver2D:
    def __init__(self, gres, bound_size, dtype=torch.float64):
        self.gres = gres
        self._g = A.to_host_ints(gres)
        if len(self._g) != 2:
            raise ValueError("ViscosityCGSolver2D needs a 2-entry gres")
        self.cell_size = A.to_host_f64(bound_size, 2) / np.asarray(self._g, dtype=np.float64)
        self.cell_vol = float(np.prod(self.cell_size))
        self._code = _DT[dtype]
        self._e = _Engine(self._g, self._code)
        for vec, nm in ((N.VEC_D, "d"), (N.VEC_R, "r"), (N.VEC_Q, "q"), (N.VEC_X, "x"), (N.VEC_B, "b")):
            for c, ax in enumerate("xy"):
                setattr(self, f"{nm}_{ax}", self._e.view(vec, c))
        self.alpha = 0.0
        self.beta = 0.0
        self.delta = 0.0
        self.iterations = 0
        self.max_iter = int(np.prod(np.asarray(self._g, dtype=np.int64)))

    def solve(self, dt, mu, rho, vx, vy, sphi, sv, lphi, lvol, tol=1e-4, save=False):
        g, e = self._g, self._e
        v = _mac_args(g, (vx, vy), ("vx", "vy"))
        if v[0].code != v[1].code:
            raise TypeError("vx, vy must share one dtype")
        s = A.as_arg(sphi, "sphi", shape=_fine_shape(g), want=torch.float64)
        vl = A.as_arg(lvol, "lvol", shape=_fine_shape(g), want=torch.float64)
        st = N.CgStats()
        status = N.check(
            e.lib.fs_visc2d_solve(e.h, float(dt), float(mu), float(rho), self.cell_vol, v[0].ptr, v[1].ptr, v[0].code,
                                  s.ptr, vl.ptr, float(tol), int(self.max_iter), ctypes.byref(st), A.stream_ptr()),
            "fs_visc2d_solve")
        self.delta, self.alpha, self.beta, self.iterations = st.delta, st.alpha, st.beta, int(st.iterations)
        if status == N.FS_NOT_CONVERGED:
            raise ValueError("Failed to converge!")
        for a in v:
            a.sync_back()
