"""B200-native drop-in for the reference's ``solver/ViscosityCGSolver3D.py``.

Same public surface (module functions ``extrapolate``, ``initialize_solver``, ``matvecmul``,
``apply_viscosity`` with the reference's argument order — ViscosityCGSolver3D.py:472, :504, :515, :526 —
and ``class ViscosityCGSolver3D`` with ``__init__(gres, bound_size)`` / ``solve(dt, mu, rho, vx, vy, vz,
sphi, sv, lphi, lvol, tol=1e-3)`` — :532-613), but every array operation runs in hand-written sm_100a
kernels behind the C ABI (``include/fluidsolver_b200.h``).  No CuPy / Numba, no CPU fallback.

Differences a caller can observe (all documented in DESIGN.md):
  * solver vectors live in a padded SoA lattice; ``x_x`` … ``b_z`` are strided torch views into it;
  * ``dtype=torch.float32`` selects fp32 vector/coefficient storage with fp64 dot products (the
    reference is fp64 throughout; fp64 is the default here);
  * a NaN residual aborts with the same ``ValueError("Failed to converge!")`` instead of spinning
    for ``prod(gres)`` iterations;
  * ``solve()`` loads and extrapolates the velocities only where it reads them (within five lattice layers
    of a row the CG computes): ``vx, vy, vz``, ``b_*``, ``r_*``, ``delta`` and ``iterations`` are bit-identical to the
    whole-lattice set-up, but ``x_*`` keeps stale values far from the liquid
    (``solver._native.set_option("sparse_setup", 0)`` restores the reference's behaviour for ``x_*``; the module-level
    ``extrapolate`` is always dense).
"""
import ctypes

import numpy as np
import torch

from . import _arrays as A
from . import _native as N

_DT = {torch.float32: N.FS_F32, torch.float64: N.FS_F64, "float32": N.FS_F32, "float64": N.FS_F64,
       np.float32: N.FS_F32, np.float64: N.FS_F64}
_TORCH = {N.FS_F32: torch.float32, N.FS_F64: torch.float64}


def _mac_shapes(g):
    return [tuple(n + (1 if i == a else 0) for i, n in enumerate(g)) for a in range(len(g))]


def _fine_shape(g):
    return tuple(2 * n + 1 for n in g)


class _RawDeviceBuffer:
    """Lets torch alias a device allocation owned by the native library (``fs_shared_alloc``)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class _Engine:
    """Owns one native fs_visc3d handle + its device workspace.

    ``shared=True`` allocates the workspace with ``fs_shared_alloc`` (plain cudaMalloc, exportable through CUDA IPC) so
    that neighbouring ranks can map it; otherwise it comes from torch's caching allocator."""

    def __init__(self, g, dtype_code, shared=False):
        self.lib = N.load()
        self.g = tuple(g)
        self.code = dtype_code
        self.tdtype = _TORCH[dtype_code]
        nbytes = self.lib.fs_visc3d_workspace_bytes(*self.g, dtype_code)
        if nbytes == 0:
            raise ValueError(f"invalid grid resolution {self.g}")
        self.nbytes = nbytes
        self.shared_ptr = None
        if shared:
            A.device()
            self.shared_ptr = self.lib.fs_shared_alloc(nbytes)
            if not self.shared_ptr:
                raise N.NativeError("fs_shared_alloc failed: " + (self.lib.fs_last_error() or b"?").decode())
            self.ws = torch.as_tensor(_RawDeviceBuffer(self.shared_ptr, nbytes), device=A.device())
        else:
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=A.device())
        h = ctypes.c_void_p()
        N.check(self.lib.fs_visc3d_create(ctypes.byref(h), *self.g, dtype_code, self.ws.data_ptr(), nbytes), "fs_visc3d_create")
        self.h = h
        X, Y, Zp, NL = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
        N.check(self.lib.fs_visc3d_lattice(h, X, Y, Zp, NL), "fs_visc3d_lattice")
        self.lat = (X.value, Y.value, Zp.value)
        self.NL = NL.value

    def close(self):
        """Destroy the native handle and release the workspace (views into it become invalid)."""
        if getattr(self, "h", None):
            self.lib.fs_visc3d_destroy(self.h)
            self.h = None
        if getattr(self, "shared_ptr", None):
            self.ws = None
            self.lib.fs_shared_free(self.shared_ptr)
            self.shared_ptr = None

    def leak(self):
        """Drop the handle but keep an IPC-exported workspace allocated: peers may still have it mapped."""
        if getattr(self, "h", None):
            self.lib.fs_visc3d_destroy(self.h)
            self.h = None
        self.shared_ptr = None

    def __del__(self):
        try:
            if getattr(self, "shared_ptr", None):
                self.leak()              # never free exported memory without the collective close() of the owning solver
            else:
                self.close()
        except Exception:
            pass

    def view(self, vec, comp):
        """torch view (reference shape) of component `comp` of solver vector `vec` inside the workspace."""
        p = self.lib.fs_visc3d_vector_ptr(self.h, vec, comp)
        esz = 4 if self.code == N.FS_F32 else 8
        off = p - self.ws.data_ptr()
        flat = self.ws[off: off + self.NL * esz].view(self.tdtype)
        X, Y, Zp = self.lat
        sh = _mac_shapes(self.g)[comp]
        return flat.view(X, Y, Zp)[: sh[0], : sh[1], : sh[2]]

    # thin wrappers -----------------------------------------------------------------------------
    def set_active_mode(self, mode):
        code = {"nonzero": N.ACTIVE_NONZERO, "fluid": N.ACTIVE_FLUID}.get(mode)
        if code is None:
            raise ValueError("active_set must be 'nonzero' or 'fluid'")
        N.check(self.lib.fs_visc3d_set_active_mode(self.h, code), "fs_visc3d_set_active_mode")
        self.active_mode_name = mode

    def set_cg_mode(self, mode):
        code = {"auto": N.CG_AUTO, "kernels": N.CG_KERNELS, "persistent": N.CG_PERSISTENT, "kernels_sr": N.CG_KERNELS_SR,
                "persistent_sr": N.CG_PERSISTENT_SR}.get(mode)
        if code is None:
            raise ValueError("cg_mode must be 'auto', 'kernels', 'persistent', 'kernels_sr' or 'persistent_sr'")
        N.check(self.lib.fs_visc3d_set_cg_mode(self.h, code), "fs_visc3d_set_cg_mode")

    def active_info(self):
        a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        N.check(self.lib.fs_visc3d_active_info(self.h, a, b, c, A.stream_ptr()), "fs_visc3d_active_info")
        return a.value, b.value, c.value

    def pack(self, sphi, lvol, vol_norm):
        N.check(self.lib.fs_visc3d_pack(self.h, sphi.ptr, lvol.ptr, float(vol_norm), A.stream_ptr()), "fs_visc3d_pack")

    def load(self, vec, v):
        if len({a.code for a in v}) != 1:
            raise TypeError("velocity components must share one dtype")
        N.check(self.lib.fs_visc3d_load(self.h, vec, v[0].ptr, v[1].ptr, v[2].ptr, v[0].code, A.stream_ptr()), "fs_visc3d_load")

    def store(self, vec, v, mode):
        if len({a.code for a in v}) != 1:
            raise TypeError("velocity components must share one dtype")
        N.check(self.lib.fs_visc3d_store(self.h, vec, v[0].ptr, v[1].ptr, v[2].ptr, v[0].code, mode, A.stream_ptr()), "fs_visc3d_store")
        for a in v:
            a.sync_back()


_engines = {}


def _engine(g, code):
    key = (tuple(g), code, torch.cuda.current_device())
    e = _engines.get(key)
    if e is None:
        while len(_engines) >= 4:            # module-level functions reuse engines per grid; keep the 4 most recent
            _engines.pop(next(iter(_engines)))
        e = _engines[key] = _Engine(g, code)
    else:
        _engines[key] = _engines.pop(key)    # most recently used last
    return e


def _mac_args(g, arrs, names):
    return [A.as_arg(a, n, shape=s) for a, n, s in zip(arrs, names, _mac_shapes(g))]


def _code_of(dtype, arrs):
    if dtype is not None:
        return _DT[dtype]
    return N.FS_F64


# ------------------------------------------------------------------------------------------------
# module-level functions with the reference's signatures
# ------------------------------------------------------------------------------------------------

def extrapolate(gres, num_iter, vx, vy, vz, sphi, dtype=None):
    """In-place extrapolation of vx,vy,vz into solid faces (reference :472-502)."""
    g = A.to_host_ints(gres)
    e = _engine(g, _code_of(dtype, (vx, vy, vz)))
    v = _mac_args(g, (vx, vy, vz), ("vx", "vy", "vz"))
    s = A.as_arg(sphi, "sphi", shape=_fine_shape(g), want=torch.float64)
    e.pack(s, s, 1.0)
    e.load(N.VEC_X, v)
    N.check(e.lib.fs_visc3d_extrapolate(e.h, N.VEC_X, int(num_iter), A.stream_ptr()), "fs_visc3d_extrapolate")
    e.store(N.VEC_X, v, N.STORE_ALL)


def initialize_solver(gres, scale, mu, vx, vy, vz, sphi, sv, vol, b_x, b_y, b_z, dtype=None):
    """RHS build (reference :504-513).  ``sv`` is accepted and ignored exactly like the reference."""
    g = A.to_host_ints(gres)
    e = _engine(g, _code_of(dtype, (vx, vy, vz)))
    v = _mac_args(g, (vx, vy, vz), ("vx", "vy", "vz"))
    b = _mac_args(g, (b_x, b_y, b_z), ("b_x", "b_y", "b_z"))
    s = A.as_arg(sphi, "sphi", shape=_fine_shape(g), want=torch.float64)
    vl = A.as_arg(vol, "vol", shape=_fine_shape(g), want=torch.float64)
    e.pack(s, vl, 1.0)
    e.load(N.VEC_X, v)
    N.check(e.lib.fs_visc3d_rhs(e.h, float(scale), float(mu), N.VEC_X, N.VEC_B, A.stream_ptr()), "fs_visc3d_rhs")
    e.store(N.VEC_B, b, N.STORE_INTERIOR)


def matvecmul(gres, scale, mu, vx, vy, vz, out_x, out_y, out_z, sphi, vol, dtype=None):
    """out = A v (reference :515-524); boundary layer of out_* is left untouched."""
    g = A.to_host_ints(gres)
    e = _engine(g, _code_of(dtype, (vx, vy, vz)))
    v = _mac_args(g, (vx, vy, vz), ("vx", "vy", "vz"))
    o = _mac_args(g, (out_x, out_y, out_z), ("out_x", "out_y", "out_z"))
    s = A.as_arg(sphi, "sphi", shape=_fine_shape(g), want=torch.float64)
    vl = A.as_arg(vol, "vol", shape=_fine_shape(g), want=torch.float64)
    e.pack(s, vl, 1.0)
    e.load(N.VEC_D, v)
    N.check(e.lib.fs_visc3d_apply(e.h, float(scale), float(mu), N.VEC_D, N.VEC_Q, A.stream_ptr()), "fs_visc3d_apply")
    e.store(N.VEC_Q, o, N.STORE_INTERIOR)


def apply_viscosity(gres, vx, vy, vz, out_x, out_y, out_z, sphi, sv, dtype=None):
    """Masked write-back v <- out on fluid faces, indices 1..g-1 (reference :526-530, :458-470)."""
    g = A.to_host_ints(gres)
    e = _engine(g, _code_of(dtype, (out_x, out_y, out_z)))
    v = _mac_args(g, (vx, vy, vz), ("vx", "vy", "vz"))
    o = _mac_args(g, (out_x, out_y, out_z), ("out_x", "out_y", "out_z"))
    s = A.as_arg(sphi, "sphi", shape=_fine_shape(g), want=torch.float64)
    e.pack(s, s, 1.0)
    e.load(N.VEC_X, o)
    e.store(N.VEC_X, v, N.STORE_FLUID)


# ------------------------------------------------------------------------------------------------
# the solver class
# ------------------------------------------------------------------------------------------------

class ViscosityCGSolver3D:
    """Implicit variational viscosity step on a 3-D MAC grid, plain CG (reference :532-613)."""

    def __init__(self, gres, bound_size, dtype=torch.float64, active_set="nonzero", cg_mode="auto"):
        """``active_set`` (extra, not in the reference): which rows the CG kernels visit — "nonzero" (default; rows
        with at least one non-zero coefficient) or "fluid" (every row the reference computes).  ``cg_mode``: "auto"
        (single-reduction CG; persistent kernel for small working sets, stand-alone kernels otherwise), "kernels" /
        "persistent" (the reference's two-reduction recurrence as three kernels per iteration from a CUDA graph / as one
        cooperative launch) or "kernels_sr" / "persistent_sr" (the single-reduction form, explicitly).  Results are
        identical up to rounding."""
        self.gres = gres
        self._g = A.to_host_ints(gres)
        if len(self._g) != 3:
            raise ValueError("ViscosityCGSolver3D needs a 3-entry gres")
        self.cell_size = A.to_host_f64(bound_size, 3) / np.asarray(self._g, dtype=np.float64)   # :535
        self.cell_vol = float(np.prod(self.cell_size))                                            # :536
        self._code = _DT[dtype]
        self._e = _Engine(self._g, self._code)
        self._e.set_active_mode(active_set)
        self._e.set_cg_mode(cg_mode)
        for vec, nm in ((N.VEC_D, "d"), (N.VEC_R, "r"), (N.VEC_Q, "q"), (N.VEC_X, "x"), (N.VEC_B, "b")):
            for c, ax in enumerate("xyz"):
                setattr(self, f"{nm}_{ax}", self._e.view(vec, c))
        self.alpha = 0.0
        self.beta = 0.0
        self.delta = 0.0
        self.iterations = 0                      # extra: CG iterations of the last solve
        self.max_iter = int(np.prod(np.asarray(self._g, dtype=np.int64)))                         # :564

    @property
    def dtype(self):
        return _TORCH[self._code]

    def active_info(self):
        """(active 32-point segments, segments in the lattice, computed rows) of the last solve's operator."""
        return self._e.active_info()

    def solve(self, dt, mu, rho, vx, vy, vz, sphi, sv, lphi, lvol, tol=1e-3):
        """In-place implicit viscosity update of vx,vy,vz (reference :566-613).

        ``sv`` and ``lphi`` are accepted and unused, exactly as in the reference (every ``sv`` use is
        commented out there).  Raises ``ValueError("Failed to converge!")`` like the reference's
        ``for … else``.
        """
        g = self._g
        e = self._e
        v = _mac_args(g, (vx, vy, vz), ("vx", "vy", "vz"))
        if len({a.code for a in v}) != 1:
            raise TypeError("vx, vy, vz must share one dtype")
        s = A.as_arg(sphi, "sphi", shape=_fine_shape(g), want=torch.float64)
        vl = A.as_arg(lvol, "lvol", shape=_fine_shape(g), want=torch.float64)
        st = N.CgStats()
        status = N.check(
            e.lib.fs_visc3d_solve(e.h, float(dt), float(mu), float(rho), self.cell_vol,
                                  v[0].ptr, v[1].ptr, v[2].ptr, v[0].code, s.ptr, vl.ptr,
                                  float(tol), int(self.max_iter), ctypes.byref(st), A.stream_ptr()),
            "fs_visc3d_solve")
        self.delta, self.alpha, self.beta, self.iterations = st.delta, st.alpha, st.beta, int(st.iterations)
        if status == N.FS_NOT_CONVERGED:
            raise ValueError("Failed to converge!")
        for a in v:
            a.sync_back()
