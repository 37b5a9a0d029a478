"""CPU: the argument behind the sparse set-up of fs_visc3d_solve (csrc/fs_visc3d.cu visc3d_mark_region_kernel),
checked on the NumPy oracle's own extrapolation sweep (oracle/numpy_oracle.py:_extrapolate_sweep, pinned to
ViscosityCGSolver3D.py:8-39 by the reference fixtures).

Claim: if the velocities are only known within R layers (Chebyshev distance, R >= 4) of the faces a solve reads from — the
active rows — and sweep k only examines faces it can reach from there (sweep 1: the region; sweep k > 1: the neighbours of what
sweep k-1 filled), then after the three sweeps every face within one layer of an active row — the stencil neighbours the RHS and
the first apply read — holds exactly the value and validity the dense three sweeps give it.  Everything outside the region is
poisoned with NaN here, so any dependence on it would show."""
import numpy as np
import pytest

from oracle import numpy_oracle as O


def _dilate(mask, layers):
    """Chebyshev dilation of a boolean array by `layers`."""
    out = mask.copy()
    for _ in range(layers):
        grown = out.copy()
        for ax in range(3):
            for sh in (1, -1):
                rolled = np.roll(out, sh, axis=ax)
                idx = [slice(None)] * 3
                idx[ax] = 0 if sh == 1 else -1
                rolled[tuple(idx)] = False             # no wrap-around
                grown |= rolled
        # Chebyshev: dilating axis by axis from `out` three times in sequence covers the diagonals
        out = grown
        for ax in (1, 2):
            g2 = out.copy()
            for sh in (1, -1):
                rolled = np.roll(out, sh, axis=ax)
                idx = [slice(None)] * 3
                idx[ax] = 0 if sh == 1 else -1
                rolled[tuple(idx)] = False
                g2 |= rolled
            out = g2
    return out


def _neighbours(mask):
    """6-neighbourhood of a boolean array (no wrap-around)."""
    out = np.zeros_like(mask)
    for ax in range(3):
        for sh in (1, -1):
            rolled = np.roll(mask, sh, axis=ax)
            idx = [slice(None)] * 3
            idx[ax] = 0 if sh == 1 else -1
            rolled[tuple(idx)] = False
            out |= rolled
    return out


def _restricted_sweeps(v, valid, region, sweeps=3):
    """The sweeps of the sparse set-up on one component: values outside `region` are unknown (NaN); sweep 1 examines the region,
    sweep k > 1 the neighbours of the faces sweep k-1 filled; a face is filled by the oracle's rule from the state before the sweep."""
    v = np.where(region, v, np.nan)
    valid = valid.copy()
    examine = region.copy()
    for _ in range(sweeps):
        nv, nvalid = O._extrapolate_sweep(v, valid)
        filled = nvalid & ~valid & examine
        v = np.where(filled, nv, v)
        valid = valid | filled
        examine = _neighbours(filled)
    return v, valid


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("layers", [4, 5])
def test_restricted_sweeps_equal_dense_sweeps_where_the_solve_reads(seed, layers):
    rng = np.random.default_rng(seed)
    shape = (19, 17, 23)
    v0 = rng.standard_normal(shape)
    # a blobby solid: invalid faces form thick regions with fluid pockets, as in the benchmark scenes
    field = rng.standard_normal(shape)
    for _ in range(2):
        field = sum(np.roll(field, s, a) for a in range(3) for s in (-1, 0, 1)) / 9.0
    valid0 = field > -0.02
    # "active rows": some of the valid faces (near the liquid in a real scene)
    active = valid0 & (rng.random(shape) < 0.03)
    assert active.any() and (~valid0).sum() > 100
    dense_v, dense_valid = v0.copy(), valid0.copy()
    for _ in range(3):
        dense_v, dense_valid = O._extrapolate_sweep(dense_v, dense_valid)
    region = _dilate(active, layers)
    sparse_v, sparse_valid = _restricted_sweeps(v0, valid0, region)
    reads = _dilate(active, 1)                         # the stencil of an active row (incl. its diagonal neighbours)
    assert np.array_equal(sparse_valid[reads], dense_valid[reads])
    seen = reads & dense_valid                         # values of invalid faces are never read (the operator masks them)
    assert np.array_equal(sparse_v[seen], dense_v[seen])
    assert not np.isnan(sparse_v[seen]).any()
