"""TEST INFRASTRUCTURE ONLY — build + ctypes wrapper of the C/OpenMP restatement (``visc3d_port.c``).

Used by tests (as a faster checker than the NumPy oracle) and by ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs.  The reference has no CPU implementation and no compiled sources, so
``kind`` is "port".  The built ``.so`` lands in ``oracle/_build/`` (git-ignored, travels to the GPU box).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "visc3d_port.c")
OUT_DIR = os.path.join(os.path.dirname(HERE), "_build")
LIB = os.path.join(OUT_DIR, "libvisc3d_port.so")

_lib = None
_P = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", SRC, "-o", LIB, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed:\n" + r.stderr)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()                      # no-op when the .so is newer than the source
        lib = ctypes.CDLL(LIB)
        lib.port_visc3d_matvecmul.restype = None
        lib.port_visc3d_matvecmul.argtypes = [ctypes.c_int] * 3 + [ctypes.c_double] * 2 + [_P] * 8
        lib.port_visc3d_cg.restype = ctypes.c_int64
        lib.port_visc3d_cg.argtypes = [ctypes.c_int] * 3 + [ctypes.c_double] * 2 + [_P] * 14 + [ctypes.c_double, ctypes.c_int64, _P]
        lib.port_num_threads.restype = ctypes.c_int
        lib.port_set_num_threads.restype = None
        lib.port_set_num_threads.argtypes = [ctypes.c_int]
        lib.port_visc3d_rhs.restype = None
        lib.port_visc3d_rhs.argtypes = [ctypes.c_int] * 3 + [ctypes.c_double] * 2 + [_P] * 8
        lib.port_visc3d_extrapolate.restype = None
        lib.port_visc3d_extrapolate.argtypes = [ctypes.c_int] * 4 + [_P] * 4 + [_P, ctypes.POINTER(ctypes.c_ubyte)]
        _lib = lib
    return _lib


def _p(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_P)


def num_threads():
    return int(load().port_num_threads())


def use_all_cores():
    """Size the OpenMP team to the cores this process may run on.  torchrun exports OMP_NUM_THREADS=1 to its workers,
    which would silently turn the multi-core baseline into a single-core one; returns the thread count now in use."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    load().port_set_num_threads(int(n))
    return num_threads()


def initialize_solver(gres, scale, mu, vx, vy, vz, sphi, sv, vol, bx, by, bz):
    """RHS (reference :504-513); ``sv`` unused like in the reference."""
    g = [int(n) for n in gres]
    load().port_visc3d_rhs(*g, float(scale), float(mu), _p(vx), _p(vy), _p(vz), _p(bx), _p(by), _p(bz), _p(sphi), _p(vol))


def extrapolate(gres, num_iter, vx, vy, vz, sphi):
    """In-place Jacobi extrapolation (reference :472-502)."""
    g = [int(n) for n in gres]
    n = max(a.size for a in (vx, vy, vz))
    tmp_v = np.empty(n, dtype=np.float64)
    tmp_m = np.empty(2 * n, dtype=np.uint8)
    load().port_visc3d_extrapolate(*g, int(num_iter), _p(vx), _p(vy), _p(vz), _p(sphi), _p(tmp_v),
                                   tmp_m.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)))


def matvecmul(gres, scale, mu, vx, vy, vz, ox, oy, oz, sphi, vol):
    g = [int(n) for n in gres]
    load().port_visc3d_matvecmul(*g, float(scale), float(mu), _p(vx), _p(vy), _p(vz), _p(ox), _p(oy), _p(oz), _p(sphi), _p(vol))


def cg(gres, scale, mu, x, r, d, q, sphi, vol, tol2, max_iter, delta):
    """Run CG iterations in place on lists of 3 arrays; returns (iterations, delta)."""
    g = [int(n) for n in gres]
    dl = ctypes.c_double(float(delta))
    it = load().port_visc3d_cg(*g, float(scale), float(mu), *[_p(a) for a in x], *[_p(a) for a in r], *[_p(a) for a in d],
                               *[_p(a) for a in q], _p(sphi), _p(vol), float(tol2), int(max_iter), ctypes.byref(dl))
    return int(it), dl.value


class ViscosityCGSolver3D:
    """Same flow as the reference's solve() (:566-613): setup steps by the NumPy oracle, the CG loop by the C port."""

    def __init__(self, gres, bound_size):
        self.gres = np.asarray(gres, dtype=np.int64)                       # reference :533-536, :564
        self.cell_size = np.asarray(bound_size) / self.gres
        self.cell_vol = float(np.prod(self.cell_size))
        self.max_iter = int(np.prod(self.gres))
        self.iterations = 0
        self.delta = 0.0

    def prepare(self, dt, mu, rho, vx, vy, vz, sphi, lvol):
        """everything before the loop: returns the CG state (x, r, d, q, vol, scale, delta0)"""
        o = self
        scale = dt / o.cell_vol / rho
        vol = np.ascontiguousarray(lvol / (o.cell_vol * 0.125))
        sphi = np.ascontiguousarray(sphi, dtype=np.float64)
        x = [np.ascontiguousarray(a, dtype=np.float64).copy() for a in (vx, vy, vz)]
        extrapolate(o.gres, 3, *x, sphi)
        b = [np.zeros_like(a) for a in x]
        initialize_solver(o.gres, scale, mu, *x, sphi, None, vol, *b)
        q = [np.zeros_like(a) for a in x]
        matvecmul(o.gres, scale, mu, *x, *q, sphi, vol)
        d = [bb - qq for bb, qq in zip(b, q)]
        r = [a.copy() for a in d]
        delta = float(sum(np.sum(a ** 2) for a in r))
        return dict(x=x, r=r, d=d, q=q, b=b, vol=vol, scale=scale, delta=delta, sphi=sphi)

    def solve(self, dt, mu, rho, vx, vy, vz, sphi, sv, lphi, lvol, tol=1e-3):
        from oracle import numpy_oracle as O
        st = self.prepare(dt, mu, rho, vx, vy, vz, sphi, lvol)
        self.iterations, self.delta = 0, st["delta"]
        if not st["delta"] < tol ** 2:
            self.iterations, self.delta = cg(self.gres, st["scale"], mu, st["x"], st["r"], st["d"], st["q"], st["sphi"], st["vol"],
                                             tol ** 2, self.max_iter, st["delta"])
            if not self.delta < tol ** 2:
                raise ValueError("Failed to converge!")
        O.visc3d_apply_viscosity(self.gres, vx, vy, vz, *st["x"], sphi, sv)
        self.x = st["x"]
