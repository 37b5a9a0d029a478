"""TEST INFRASTRUCTURE ONLY — NumPy restatement of the notebook's grid-side kernels.

Follows the code cells of ``3D_viscous_fluid_sim.ipynb``: p2g_particle / p2g_grid (:279-344), g2p_particle (:352-393),
compute_fls_kernel (:94-136), compute_fluid_volume_kernel / constrain (:224-268), extrapolate (:501-557),
boundary_condition_{x,y,z} (cell 5).  Vectorised over particles / faces; fp32 exactly where the notebook declares fp32
local arrays.  The reference accumulates with floating-point atomics (order-dependent) into fp32 grids, so parity for the
scattered quantities is a tolerance (1e-5 of the field's maximum), not bit equality; index sets (which nodes receive mass,
which cells the level set touches, which faces get extrapolated / a boundary correction) are exact.
Pinned against fixtures produced by the notebook's own cells under Numba's simulator (tests/test_notebook_kernels_cpu.py).
"""
import itertools

import numpy as np

F32, F64 = np.float32, np.float64


def _stencil(px, bound_min, cell, bias):
    x = px.astype(F32)
    bmin = np.asarray(bound_min, dtype=F32)
    bias = np.asarray(bias, dtype=F32)
    cell = np.asarray(cell, dtype=F64)
    t = (x - bmin).astype(F64) / cell - bias.astype(F64)
    gi = np.floor(t).astype(np.int64)
    gx = ((gi + bias.astype(F64)) * cell + bmin.astype(F64)).astype(F32)
    disp = gx - x
    w = (np.abs(disp).astype(F64) / cell).astype(F32)
    return gi, disp, w


def _corner(gi, w, gres, ix, iy, iz):
    idx = tuple(np.clip(gi[:, d] + o, 0, gres[d] - 1) for d, o in enumerate((ix, iy, iz)))
    ws = [w[:, d] if o else (F32(1) - w[:, d]) for d, o in enumerate((ix, iy, iz))]
    return idx, ws


def p2g(gres, bound_min, cell, px, pm, pv, pc, m_grids, v_grids):
    """pc = (cx, cy, cz); m_grids / v_grids: zeroed fp32 MAC arrays, filled in place (incl. the v/m normalisation)"""
    cell = np.asarray(cell, dtype=F64)
    for axis in range(3):
        bias = np.full(3, 0.5, dtype=F32)
        bias[axis] = 0
        gi, disp, w = _stencil(px, bound_min, cell, bias)
        gm = np.zeros(m_grids[axis].shape, dtype=F64)
        gv = np.zeros(v_grids[axis].shape, dtype=F64)
        va = pv[:, axis].astype(F32).astype(F64)
        for ix, iy, iz in itertools.product((0, 1), repeat=3):
            idx, ws = _corner(gi, w, gres, ix, iy, iz)
            cv = sum((disp[:, d].astype(F64) + o * cell[d]) * pc[axis][:, d] for d, o in enumerate((ix, iy, iz)))
            weight = (ws[0] * ws[1] * ws[2]).astype(F64)
            np.add.at(gm, idx, weight * pm)
            np.add.at(gv, idx, weight * pm * (va + cv))
        m_grids[axis][...] = gm.astype(F32)
        vv = gv.astype(F32)
        pos = m_grids[axis] > 0
        vv[pos] = vv[pos] / m_grids[axis][pos]
        v_grids[axis][...] = vv


def g2p(gres, bound_min, cell, px, v_grids):
    """returns pv (P,3) and (cx, cy, cz) each (P,3), fp64"""
    cell = np.asarray(cell, dtype=F64)
    n = px.shape[0]
    pv = np.zeros((n, 3))
    pc = [np.zeros((n, 3)) for _ in range(3)]
    for axis in range(3):
        bias = np.full(3, 0.5, dtype=F32)
        bias[axis] = 0
        gi, _, w = _stencil(px, bound_min, cell, bias)
        for ix, iy, iz in itertools.product((0, 1), repeat=3):
            idx, ws = _corner(gi, w, gres, ix, iy, iz)
            val = v_grids[axis][idx]
            pv[:, axis] += (ws[0] * ws[1] * ws[2] * val).astype(F64)
            sg = [F32(2 * o - 1) for o in (ix, iy, iz)]
            pc[axis][:, 0] += (sg[0] * ws[1] * ws[2] * val).astype(F64) / cell[0]
            pc[axis][:, 1] += (ws[0] * sg[1] * ws[2] * val).astype(F64) / cell[1]
            pc[axis][:, 2] += (ws[0] * ws[1] * sg[2] * val).astype(F64) / cell[2]
    return pv, pc


def fluid_levelset(gres, bound_min, cell, px, gdx):
    cell = np.asarray(cell, dtype=F64)
    bmin = np.asarray(bound_min, dtype=F32)
    r = gdx * 0.5 * np.sqrt(3.0) * 1.02
    phi = np.full(tuple(gres), gdx * 3, dtype=F64)
    x = px.astype(F32)
    gi = np.floor((x - bmin).astype(F64) / cell).astype(np.int64)
    for off in itertools.product(range(-2, 3), repeat=3):
        gii = [np.clip(gi[:, d] + off[d], 0, gres[d] - 1) for d in range(3)]
        gip = [(((gii[d] + 0.5) * cell[d] + F64(bmin[d])) - x[:, d].astype(F64)).astype(F32) for d in range(3)]
        n2 = sum((g * g).astype(F64) for g in gip)
        np.minimum.at(phi, tuple(gii), np.sqrt(n2) - r)
    return phi


def fluid_volume(res, bound_min, cell, px, pvol):
    """res = 2*gres+1 node grid, cell = bound_size / (2*gres)"""
    cell = np.asarray(cell, dtype=F64)
    gi, _, w = _stencil(px, bound_min, cell, np.zeros(3, dtype=F32))
    vol = np.zeros(tuple(res), dtype=F64)
    for ix, iy, iz in itertools.product((0, 1), repeat=3):
        idx, ws = _corner(gi, w, res, ix, iy, iz)
        np.add.at(vol, idx, (ws[0] * ws[1] * ws[2]).astype(F64) * pvol)
    return np.minimum(vol, float(np.prod(cell)))


def extrapolate(num_iter, vs, ms):
    """in place on the fp32 arrays vs = [vx, vy, vz]; validity = mass > 0; returns the final validity masks"""
    valids = [m > 0 for m in ms]
    for _ in range(num_iter):
        for c in range(3):
            v, valid = vs[c], valids[c]
            if min(v.shape) < 3:
                continue
            inner = tuple(slice(1, -1) for _ in range(3))
            val = np.zeros(tuple(n - 2 for n in v.shape), dtype=F64)
            cnt = np.zeros(val.shape, dtype=np.int64)
            for ax in range(3):
                for sgn in (1, -1):
                    sl = tuple(slice(1 + (sgn if d == ax else 0), v.shape[d] - 1 + (sgn if d == ax else 0)) for d in range(3))
                    mk = valid[sl]
                    val = val + np.where(mk, v[sl].astype(F64), 0.0)
                    cnt = cnt + mk
            upd = (~valid[inner]) & (cnt > 0)
            with np.errstate(invalid="ignore", divide="ignore"):
                filled = (val / cnt).astype(F32)
            nv, nvalid = v.copy(), valid.copy()
            nv[inner] = np.where(upd, filled, v[inner])
            nvalid[inner] = valid[inner] | upd
            vs[c][...] = nv
            valids[c] = nvalid
    return valids


def boundary_condition(gres, dx, vs, ms, sphi, sv):
    """returns [dvx, dvy, dvz] (fp32) computed from the OLD velocities; the caller adds them (apply_boundary_condition)"""
    g = [int(n) for n in gres]
    out = []
    for A in range(3):
        sh = list(g)
        sh[A] += 1
        dv = np.zeros(tuple(sh), dtype=F32)
        if min(sh) < 3:
            out.append(dv)
            continue
        c = np.meshgrid(*[np.arange(1, n - 1) for n in sh], indexing="ij")
        f = [2 * c[d] + (0 if d == A else 1) for d in range(3)]
        ndist = sphi[f[0], f[1], f[2]] / dx
        vel = [None] * 3
        vel[A] = vs[A][c[0], c[1], c[2]].astype(F64)
        for B in range(3):
            if B == A:
                continue
            msum = np.zeros(c[0].shape)
            vsum = np.zeros(c[0].shape)
            for ia, ib in itertools.product((0, 1), repeat=2):
                q = [c[0], c[1], c[2]]
                q[A] = q[A] - ia
                q[B] = q[B] + ib
                mm = ms[B][q[0], q[1], q[2]]
                msum += mm.astype(F64)
                vsum += (vs[B][q[0], q[1], q[2]] * mm).astype(F64)
            with np.errstate(invalid="ignore", divide="ignore"):
                vel[B] = vsum / msum
        for d in range(3):
            vel[d] = vel[d] - sv[f[0], f[1], f[2], d]
        sn = []
        for d in range(3):
            fp = [f[k] + (1 if k == d else 0) for k in range(3)]
            fm = [f[k] - (1 if k == d else 0) for k in range(3)]
            sn.append(sphi[fp[0], fp[1], fp[2]] - sphi[fm[0], fm[1], fm[2]])
        with np.errstate(invalid="ignore", divide="ignore"):
            sn_inv = 1.0 / (sn[0] ** 2 + sn[1] ** 2 + sn[2] ** 2)
            dot = sn[0] * vel[0] + sn[1] * vel[1] + sn[2] * vel[2]
            gsn = np.where(dot < 0, dot, 0.0) * sn[A] * sn_inv        # min(0, NaN) = 0 in Python: NaN-safe like the reference
            val = (-gsn * (1.0 - ndist)).astype(F32)
        dv[1:-1, 1:-1, 1:-1] = np.where(ndist >= 1, F32(0), val)
        out.append(dv)
    return out
