"""CPU: the integer index arithmetic of the work decompositions in csrc/fs_visc3d.cu, restated in Python and checked
exhaustively on small cases — every unit of work is visited exactly once, and the marked region covers what the sparse
set-up needs.  (The kernels themselves are checked on the GPU against the other forms: tests/test_resident_forms_gpu.py.)"""
import itertools

import numpy as np
import pytest


# ---- visc3d_apply_dot_body: warp -> list position of trip q -------------------------------------------------------------
def _k1_positions(nseg, grid, warps_per_cta, blk):
    """positions visited by every warp of the grid, in the order the kernel's loop generates them (kof in the kernel)"""
    seen = []
    nw = grid * warps_per_cta
    for b in range(grid):
        for wl in range(warps_per_cta):
            w0 = b * warps_per_cta + wl
            span = blk * warps_per_cta
            trip = 0
            last = -1
            while True:
                if blk <= 0:
                    k = w0 + trip * nw
                else:
                    m, t = divmod(trip, blk)
                    k = (m * grid + b) * span + t * warps_per_cta + wl
                assert k > last                      # strictly increasing: the kernel's loop ends at the first k >= nseg
                last = k
                if k >= nseg:
                    break
                seen.append(k)
                trip += 1
    return seen


@pytest.mark.parametrize("blk", [0, 1, 3, 4, 8])
@pytest.mark.parametrize("nseg,grid,wc", [(0, 1, 8), (1, 1, 8), (37, 2, 8), (1000, 7, 8), (4253, 296, 8), (5000, 148, 16), (1023, 64, 8)])
def test_k1_mapping_visits_every_segment_once(nseg, grid, wc, blk):
    seen = _k1_positions(nseg, grid, wc, blk)
    assert sorted(seen) == list(range(nseg))


def test_k1_blocked_mapping_keeps_a_cta_on_consecutive_rows():
    """blk consecutive trips of a CTA cover blk * warps consecutive list entries (consecutive lattice rows): what lets L1 serve
    the y-neighbour rows of one trip as the own rows of the next."""
    grid, wc, blk, nseg = 5, 8, 4, 2000
    for b in range(grid):
        per_trip = []
        for trip in range(2 * blk):
            m, t = divmod(trip, blk)
            per_trip.append([(m * grid + b) * blk * wc + t * wc + wl for wl in range(wc)])
        first_block = sorted(k for tr in per_trip[:blk] for k in tr)
        assert first_block == list(range(first_block[0], first_block[0] + blk * wc))
        assert per_trip[blk][0] - per_trip[0][0] == grid * blk * wc          # then the CTA jumps ahead by the grid


# ---- visc3d_cg_sr_resident2_kernel: contiguous runs per CTA, slot = position in the run ----------------------------------
@pytest.mark.parametrize("nseg,grid", [(4253, 148), (29, 148), (0, 1), (4292, 148), (5000, 148), (100, 3)])
def test_resident_runs_partition_the_list(nseg, grid):
    chunk = (nseg + grid - 1) // grid
    seen = []
    for b in range(grid):
        c_lo = b * chunk
        n_own = min(c_lo + chunk, nseg) - c_lo
        for warp in range(16):
            for li in range(warp, max(n_own, 0), 16):
                seen.append(c_lo + li)
    assert sorted(seen) == list(range(nseg))
    if nseg == 4253:
        assert chunk == 29 and 29 * (31 * 32 * 8 + 32 + 4) <= 232448 - 1024       # every segment of the benchmark scene is resident


# ---- visc3d_apply_dot2_tile_kernel: work items (row block, plane chunk) ---------------------------------------------------
@pytest.mark.parametrize("nx,Y,Zp,ty,xl", [(36, 41, 48, 11, 16), (21, 18, 32, 17, 16), (256, 257, 260, 2, 16), (21, 18, 32, 5, 1), (9, 9, 12, 3, 4)])
def test_tile_work_items_cover_every_computable_point_once(nx, Y, Zp, ty, xl):
    sx = Y * Zp
    tile = ty * Zp
    nyb = (Y + ty - 1) // ty
    x_first, x_end = 1, nx
    nxc = (x_end - x_first + xl - 1) // xl
    count = np.zeros((nx + 1) * sx, dtype=np.int32)
    for item in range(nyb * nxc):
        xc, yb = divmod(item, nyb)
        xa = x_first + xc * xl
        xb = min(xa + xl, x_end)
        row0 = yb * tile
        for x in range(xa, xb):
            t = np.arange(tile)
            t = t[row0 + t < sx]
            count[x * sx + row0 + t] += 1
    planes = count.reshape(nx + 1, sx)
    assert (planes[1:nx] == 1).all()                 # planes 1 .. nx-1: the only ones that can hold computed rows
    assert (planes[0] == 0).all() and (planes[nx] == 0).all()


# ---- visc3d_mark_region_kernel: the marked segments cover the neighbourhood the sparse set-up needs -----------------------
@pytest.mark.parametrize("window", [None, (3, 9)])
@pytest.mark.parametrize("seed", [0, 1])
def test_marked_region_covers_four_layers_around_active_segments(seed, window):
    nx, ny, nz = 12, 9, 37
    X, Y, Zp = nx + 1, ny + 1, (nz + 1 + 3) // 4 * 4
    sx, sy, NL = Y * Zp, Zp, X * Y * Zp
    nseg_total = (NL + 31) // 32
    wlo, whi = (0, nx) if window is None else window
    rng = np.random.default_rng(seed)
    lo_seg, hi_seg = wlo * sx // 32, ((whi + 1) * sx + 31) // 32
    active = np.sort(rng.choice(np.arange(lo_seg, min(hi_seg, nseg_total)), size=6, replace=False))
    LAY, PTS = 5, 8                                   # kRegionLayers, kRegionPoints
    flags = np.zeros(nseg_total, dtype=bool)
    w_lo, w_hi = wlo * sx, (whi + 1) * sx - 1
    for sg in active:
        for dx, dy in itertools.product(range(-LAY, LAY + 1), repeat=2):
            base = int(sg) * 32 + dx * sx + dy * sy
            lo, hi = base - PTS, base + 31 + PTS
            if hi < w_lo or lo > w_hi:
                continue
            lo, hi = max(lo, w_lo), min(hi, w_hi)
            flags[lo // 32: min(hi // 32, nseg_total - 1) + 1] = True
    # every lattice point within 4 layers (x, y) and 4 points (z) of a point of an active segment, inside the window, is marked
    for sg in active:
        for p in range(int(sg) * 32, min(int(sg) * 32 + 32, NL)):
            x, rem = divmod(p, sx)
            y, z = divmod(rem, sy)
            for dx, dy, dz in itertools.product(range(-4, 5), range(-4, 5), range(-4, 5)):
                xx, yy, zz = x + dx, y + dy, z + dz
                if not (wlo <= xx <= whi and 0 <= yy < Y and 0 <= zz < Zp):
                    continue
                q = (xx * Y + yy) * Zp + zz
                assert flags[q // 32], (sg, p, dx, dy, dz)
    # and nothing outside the window's planes is marked beyond the segments that straddle its first / last point
    marked = np.nonzero(flags)[0]
    assert marked.min() >= w_lo // 32 and marked.max() <= w_hi // 32
