"""Drop-in for ``solver/SolidFraction2D.py``: ``compute_solid_frac(gres, sphi, wx, wy)`` (:22-26).

Continuous edge open fractions ``1 - edge_in_fraction``; entries ``wx[W,:]``, ``wx[:,H-1]``, ``wy[W-1,:]``,
``wy[:,H]`` are never written, exactly as in the reference."""
import torch

from . import _arrays as A
from . import _native as N
from .SolidFractionCommon import edge_in_fraction  # noqa: F401


def compute_solid_frac(gres, sphi, wx, wy):
    g = A.to_host_ints(gres)
    if len(g) != 2:
        raise ValueError("SolidFraction2D.compute_solid_frac needs a 2-entry gres")
    fine = tuple(2 * n + 1 for n in g)
    s = A.as_arg(sphi, "sphi", shape=fine, want=torch.float64)
    ws = [A.as_arg(w, nm, shape=tuple(n + (1 if i == a else 0) for i, n in enumerate(g)), dtypes=(torch.float64,))
          for a, (w, nm) in enumerate(((wx, "wx"), (wy, "wy")))]
    lib = N.load()
    N.check(lib.fs_solidfrac2d(*g, s.ptr, ws[0].ptr, ws[1].ptr, A.stream_ptr()), "fs_solidfrac2d")
    for w in ws:
        w.sync_back()
