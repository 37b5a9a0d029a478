"""ctypes binding of ``libfluidsolver_b200.so`` (the C ABI in ``include/fluidsolver_b200.h``).

The product path has NO fallback: if the library is missing or a call fails, this module raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libfluidsolver_b200.so")

FS_OK = 0
FS_NOT_CONVERGED = 1
FS_F32 = 0
FS_F64 = 1
VEC_X, VEC_R, VEC_D, VEC_Q, VEC_B = range(5)
STORE_ALL, STORE_INTERIOR, STORE_FLUID = range(3)
ACTIVE_FLUID, ACTIVE_NONZERO = range(2)
CG_AUTO, CG_KERNELS, CG_PERSISTENT, CG_KERNELS_SR, CG_PERSISTENT_SR = range(5)
OP_PRESSURE, OP_DENSITY = range(2)


class CgStats(Structure):
    _fields_ = [
        ("iterations", c_int64),
        ("delta", c_double),
        ("alpha", c_double),
        ("beta", c_double),
        ("delta0", c_double),
        ("converged", c_int32),
        ("reserved", c_int32),
    ]


_SIGS = {
    "fs_abi_version": (c_int, []),
    "fs_last_error": (c_char_p, []),
    "fs_launch_count": (c_int64, []),
    "fs_set_option": (c_int, [c_char_p, c_int]),
    # viscosity 3-D
    "fs_visc3d_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "fs_visc3d_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_void_p, c_size_t]),
    "fs_visc3d_destroy": (None, [c_void_p]),
    "fs_visc3d_lattice": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int64)]),
    "fs_visc3d_vector_ptr": (c_void_p, [c_void_p, c_int, c_int]),
    "fs_visc3d_set_active_mode": (c_int, [c_void_p, c_int]),
    "fs_visc3d_set_cg_mode": (c_int, [c_void_p, c_int]),
    "fs_visc3d_cg_mode_in_use": (c_int, [c_void_p]),
    "fs_visc3d_debug_read": (c_int, [c_void_p, c_int, c_void_p, c_size_t]),
    "fs_visc3d_active_info": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), c_void_p]),
    "fs_visc3d_pack": (c_int, [c_void_p, c_void_p, c_void_p, c_double, c_void_p]),
    "fs_visc3d_load": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "fs_visc3d_store": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "fs_visc3d_extrapolate": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "fs_visc3d_rhs": (c_int, [c_void_p, c_double, c_double, c_int, c_int, c_void_p]),
    "fs_visc3d_apply": (c_int, [c_void_p, c_double, c_double, c_int, c_int, c_void_p]),
    "fs_visc3d_cg": (c_int, [c_void_p, c_double, c_double, c_double, c_int64, POINTER(CgStats), c_void_p]),
    "fs_visc3d_cg_enqueue": (c_int, [c_void_p, c_double, c_double, c_int64, c_void_p]),
    "fs_visc3d_kernel_enqueue": (c_int, [c_void_p, c_int, c_double, c_double, c_int64, c_void_p]),
    "fs_visc3d_read_stats": (c_int, [c_void_p, POINTER(CgStats), c_void_p]),
    "fs_visc3d_solve": (c_int, [c_void_p, c_double, c_double, c_double, c_double, c_void_p, c_void_p, c_void_p, c_int,
                                c_void_p, c_void_p, c_double, c_int64, POINTER(CgStats), c_void_p]),
    # multi-GPU
    "fs_comm_unique_id": (c_int, [c_void_p]),
    "fs_comm_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_void_p]),
    "fs_comm_destroy": (None, [c_void_p]),
    "fs_comm_rank": (c_int, [c_void_p]),
    "fs_comm_size": (c_int, [c_void_p]),
    "fs_visc3d_set_slab": (c_int, [c_void_p, c_void_p, c_int, c_int]),
    "fs_shared_alloc": (c_void_p, [c_size_t]),
    "fs_shared_free": (None, [c_void_p]),
    "fs_shared_get_handle": (c_int, [c_void_p, c_void_p]),
    "fs_shared_open": (c_void_p, [c_void_p]),
    "fs_shared_close": (None, [c_void_p]),
    "fs_visc3d_set_peers": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, POINTER(c_void_p)]),
    "fs_visc3d_peer_error": (c_int, [c_void_p]),
    # gathered multi-GPU solve / windowed inputs
    "fs_visc3d_set_window": (c_int, [c_void_p, c_int, c_int]),
    "fs_visc3d_gather_record_bytes": (c_size_t, [c_void_p]),
    "fs_visc3d_gather_export": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_double, c_int, c_int,
                                        c_void_p, c_int64, POINTER(c_int64), c_void_p]),
    "fs_visc3d_gather_reexport": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "fs_visc3d_gather_import": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p]),
    "fs_visc3d_solve_packed": (c_int, [c_void_p, c_double, c_double, c_double, c_double, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                       c_double, c_int64, POINTER(CgStats), c_void_p]),
    # viscosity 2-D
    "fs_visc2d_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "fs_visc2d_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_void_p, c_size_t]),
    "fs_visc2d_destroy": (None, [c_void_p]),
    "fs_visc2d_lattice": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int64)]),
    "fs_visc2d_vector_ptr": (c_void_p, [c_void_p, c_int, c_int]),
    "fs_visc2d_pack": (c_int, [c_void_p, c_void_p, c_void_p, c_double, c_void_p]),
    "fs_visc2d_load": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "fs_visc2d_store": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "fs_visc2d_rhs": (c_int, [c_void_p, c_double, c_double, c_int, c_int, c_void_p]),
    "fs_visc2d_apply": (c_int, [c_void_p, c_double, c_double, c_int, c_int, c_void_p]),
    "fs_visc2d_cg": (c_int, [c_void_p, c_double, c_double, c_double, c_int64, POINTER(CgStats), c_void_p]),
    "fs_visc2d_solve": (c_int, [c_void_p, c_double, c_double, c_double, c_double, c_void_p, c_void_p, c_int,
                                c_void_p, c_void_p, c_double, c_int64, POINTER(CgStats), c_void_p]),
    # solid fractions
    "fs_solidfrac3d": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fs_solidfrac2d": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    # pressure
    "fs_press_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "fs_press_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_void_p, c_size_t]),
    "fs_press_destroy": (None, [c_void_p]),
    "fs_press_set_operator": (c_int, [c_void_p, c_int]),
    "fs_press_rhs": (c_int, [c_void_p, POINTER(c_double), c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p]),
    "fs_press_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fs_press_update": (c_int, [c_void_p, POINTER(c_double), c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "fs_press_cg": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_double, c_int64, POINTER(CgStats), c_void_p]),
    "fs_press_cg_enqueue": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int64, c_void_p]),
    # density solver (DensityCGSolver3D)
    "fs_dens3d_scatter": (c_int, [c_int, c_int, c_int, POINTER(c_double), POINTER(c_double), c_void_p, c_int, c_void_p, c_int, c_int64, c_double,
                                  c_void_p, c_void_p, c_void_p]),
    "fs_dens3d_fix_volume": (c_int, [c_int, c_int, c_int, POINTER(c_double), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fs_dens3d_rhs": (c_int, [c_int, c_int, c_int, c_double, c_double, POINTER(c_double), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p]),
    "fs_dens3d_displacement": (c_int, [c_int, c_int, c_int, c_double, POINTER(c_double), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fs_dens3d_gather": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_int, c_int, POINTER(c_double), POINTER(c_double), POINTER(c_double),
                                 c_int, c_void_p]),
    # grid-side steps of the time loop (notebook kernels)
    "fs_grid_p2g": (c_int, [c_int, c_int, c_int, POINTER(c_double), POINTER(c_double), c_int64] + [c_void_p] * 12 + [c_void_p]),
    "fs_grid_g2p": (c_int, [c_int, c_int, c_int, POINTER(c_double), POINTER(c_double), c_int64] + [c_void_p] * 8 + [c_void_p]),
    "fs_grid_levelset": (c_int, [c_int, c_int, c_int, POINTER(c_double), POINTER(c_double), c_int64, c_void_p, c_double, c_double, c_void_p, c_void_p]),
    "fs_grid_fluid_volume": (c_int, [c_int, c_int, c_int, POINTER(c_double), POINTER(c_double), c_int64, c_void_p, c_double, c_double, c_void_p, c_void_p]),
    "fs_grid_extrapolate_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "fs_grid_extrapolate": (c_int, [c_int, c_int, c_int, c_int] + [c_void_p] * 6 + [c_void_p, c_size_t, c_void_p]),
    "fs_grid_boundary": (c_int, [c_int, c_int, c_int, c_double] + [c_void_p] * 11 + [c_void_p]),
    # rigid-body signed distance field (sdf3D)
    "fs_sdf3d_evaluate": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fs_sdf3d_project": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    # UNet surrogate: input features and output gather
    "fs_unet_features": (c_int, [c_int] * 6 + [c_void_p] * 5 + [c_double, ctypes.c_float, c_void_p, c_void_p]),
    "fs_unet_gather": (c_int, [c_int] * 6 + [c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None


def exported_symbols():
    """Names the header declares; used by the CPU-side ABI test."""
    return sorted(_SIGS)


def load():
    """dlopen the library (once) and attach the prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python python-fluid-simulation_b200/build.py` "
            "(or __graft_entry__.build()).  There is no CPU fallback for the solver path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)       # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class NativeError(RuntimeError):
    pass


def check(status, what):
    """Negative status -> RuntimeError carrying the library's message; returns status otherwise."""
    if status < 0:
        msg = load().fs_last_error()
        raise NativeError(f"{what} failed ({status}): {msg.decode() if msg else '?'}")
    return status


def launch_count():
    return int(load().fs_launch_count())


def set_option(name, value):
    """Tuning switch of the kernels (see fs_set_option in include/fluidsolver_b200.h); ``value < 0`` restores the default."""
    check(load().fs_set_option(name.encode(), int(value)), f"fs_set_option({name})")
