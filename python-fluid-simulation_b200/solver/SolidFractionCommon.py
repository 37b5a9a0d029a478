"""Shared helper kept for import compatibility with the reference's ``from .SolidFractionCommon import *``.

The device functions of the reference (``edge_in_fraction``, ``tri_in_fraction``, ``face_in_fraction``,
SolidFractionCommon.py:4-60) live inside the CUDA kernels here (csrc/fs_press.cu).  ``edge_in_fraction`` is
re-exported by the pressure modules in the reference (``from .SolidFraction3D import compute_solid_frac,
edge_in_fraction``); a host-scalar version is provided for callers that import the name.
"""


def edge_in_fraction(lval, rval):
    l_in = lval < 0
    r_in = rval < 0
    if l_in and r_in:
        return 1
    if not l_in and not r_in:
        return 0
    diff = -abs(lval - rval)
    return lval / diff if l_in else rval / diff
