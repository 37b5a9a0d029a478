"""Argument adapters: turn whatever the caller passes (torch CUDA tensors, CuPy / Numba device arrays
via ``__cuda_array_interface__``, DLPack exporters, host NumPy arrays, tuples, scalars) into what the
C ABI needs — contiguous device memory with a known dtype — and write results back in place.

The reference passes CuPy ndarrays everywhere (3D_viscous_fluid_sim.ipynb:725-774); ``gres`` is a CuPy
int64 device array and ``bound_size`` a float32 device array or a Python scalar (ipynb:651-656, 778).
"""
import numpy as np
import torch

from . import _native as N

_F = {torch.float32: N.FS_F32, torch.float64: N.FS_F64}


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("fluidsolver_b200 needs a CUDA device: there is no CPU fallback for the solver path")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def to_host_ints(gres):
    """gres in any accepted form -> tuple of Python ints."""
    if isinstance(gres, torch.Tensor):
        gres = gres.detach().cpu().numpy()
    elif hasattr(gres, "get") and not isinstance(gres, np.ndarray):      # CuPy
        gres = gres.get()
    elif hasattr(gres, "__cuda_array_interface__"):
        gres = torch.as_tensor(gres, device="cuda").cpu().numpy()
    arr = np.asarray(gres).reshape(-1)
    if arr.size not in (2, 3):
        raise ValueError(f"gres must have 2 or 3 entries, got {arr.size}")
    if not np.all(arr == np.floor(arr)):
        raise ValueError("gres must be integral")
    out = tuple(int(v) for v in arr)
    if min(out) < 1:
        raise ValueError("gres entries must be >= 1")
    return out


def to_host_f64(v, n=None):
    """bound_size / cell_size -> float64 NumPy vector (scalars broadcast to n entries)."""
    if isinstance(v, torch.Tensor):
        v = v.detach().cpu().numpy()
    elif hasattr(v, "get") and not isinstance(v, np.ndarray):
        v = v.get()
    elif hasattr(v, "__cuda_array_interface__"):
        v = torch.as_tensor(v, device="cuda").cpu().numpy()
    arr = np.asarray(v).astype(np.float64).reshape(-1)
    if n is not None and arr.size == 1:
        arr = np.repeat(arr, n)
    if n is not None and arr.size != n:
        raise ValueError(f"expected {n} entries, got {arr.size}")
    return arr


class Arg:
    """One array argument resolved to a contiguous CUDA tensor `t`; `sync_back()` propagates in-place
    results to the caller's object when a staging copy had to be made."""

    __slots__ = ("orig", "t", "staged", "host")

    def __init__(self, orig, t, staged, host):
        self.orig, self.t, self.staged, self.host = orig, t, staged, host

    @property
    def ptr(self):
        return self.t.data_ptr()

    @property
    def code(self):
        return _F[self.t.dtype]

    def sync_back(self):
        if not self.staged:
            return
        if self.host:
            np.copyto(self.orig, self.t.cpu().numpy().astype(self.orig.dtype, copy=False))
        elif isinstance(self.orig, torch.Tensor):
            self.orig.copy_(self.t)
        else:
            torch.as_tensor(self.orig, device="cuda").copy_(self.t)


def as_arg(a, name, shape=None, dtypes=(torch.float32, torch.float64), want=None):
    """Resolve `a`.  `want`: dtype to convert to when the input dtype differs (read-only inputs);
    for in/out arrays leave `want=None` so the caller's dtype is kept."""
    dev = device()
    host = False
    if isinstance(a, torch.Tensor):
        t = a
    elif isinstance(a, np.ndarray):
        host = True
        t = torch.from_numpy(np.ascontiguousarray(a))
    elif hasattr(a, "__cuda_array_interface__") or hasattr(a, "__dlpack__"):
        t = torch.as_tensor(a, device=dev) if hasattr(a, "__cuda_array_interface__") else torch.from_dlpack(a)
    else:
        raise TypeError(f"{name}: unsupported array type {type(a).__name__}")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    staged = host
    if t.device != dev:
        if t.is_cuda:
            raise ValueError(f"{name}: array lives on {t.device}, solver runs on {dev}")
        t = t.to(dev, non_blocking=True)
        staged = True
    if want is not None and t.dtype != want:
        t = t.to(want)
        staged = True
    if t.dtype not in dtypes:
        raise TypeError(f"{name}: dtype {t.dtype} not supported (expected one of {dtypes})")
    if not t.is_contiguous():
        t = t.contiguous()
        staged = True
    return Arg(a, t, staged, host)
