"""Drop-in for ``solver/SolidFraction3D.py``: ``compute_solid_frac(gres, sphi, wx, wy, wz)`` (:28-32).

Face open fractions from the solid SDF nodes, bit-exact with the reference (values in {0, .5, .75, 1});
only the low-side faces are written, the far planes keep their contents."""
import torch

from . import _arrays as A
from . import _native as N
from .SolidFractionCommon import edge_in_fraction  # noqa: F401  (re-export, as in the reference)


def compute_solid_frac(gres, sphi, wx, wy, wz):
    g = A.to_host_ints(gres)
    if len(g) != 3:
        raise ValueError("SolidFraction3D.compute_solid_frac needs a 3-entry gres")
    fine = tuple(2 * n + 1 for n in g)
    s = A.as_arg(sphi, "sphi", shape=fine, want=torch.float64)
    ws = [A.as_arg(w, nm, shape=tuple(n + (1 if i == a else 0) for i, n in enumerate(g)), dtypes=(torch.float64,))
          for a, (w, nm) in enumerate(((wx, "wx"), (wy, "wy"), (wz, "wz")))]
    lib = N.load()
    N.check(lib.fs_solidfrac3d(*g, s.ptr, ws[0].ptr, ws[1].ptr, ws[2].ptr, A.stream_ptr()), "fs_solidfrac3d")
    for w in ws:
        w.sync_back()
