"""TEST INFRASTRUCTURE ONLY — fp64 NumPy restatement of the reference's implicit-solver hot path.

This module is the *checker* for the CUDA product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it; the product package never does (it fails loudly without its CUDA
library instead).

Parity pin: the reference ships no tests or golden vectors, so the pins
are outputs of the reference's *own* kernels executed under Numba's CUDA simulator
(``oracle/refsim/run_reference.py``), committed under ``tests/golden/``.  This
restatement is checked against those fixtures by ``tests/test_oracle_vs_reference.py``
(fp values to <=1e-13 relative, masks / solid fractions bit-exact).

Every function cites the reference lines it follows (paths relative to
``/root/reference/solver/``).  The arithmetic keeps the reference's left-to-right
association (``2*scale*mu*vol*v``, sequential ``val -= ...``) so that results match the
simulator to the last bit where NumPy's elementwise ops are used.

The stencils are written as *term tables* (one row per neighbour term) interpreted by a
single vectorised evaluator, rather than as per-thread straight-line code.
"""
from __future__ import annotations

import numpy as np

F64 = np.float64

# --------------------------------------------------------------------------------------
# helpers: strided views of the (2N+1)^d fine grid and shifted views of MAC arrays
# --------------------------------------------------------------------------------------


def _fine(arr, off, oshape):
    """arr[2*i+off] for i in the interior 1..oshape-2 of every axis (fine-grid gather)."""
    sl = tuple(slice(2 + o, 2 * (n - 2) + o + 1, 2) for o, n in zip(off, oshape))
    return arr[sl]


def _coarse(arr, off, oshape):
    """arr[i+off] for i in the interior 1..oshape-2 of every axis of the OUTPUT array."""
    sl = tuple(slice(1 + o, n - 1 + o) for o, n in zip(off, oshape))
    return arr[sl]


def _interior(oshape):
    return tuple(slice(1, n - 1) for n in oshape)


# --------------------------------------------------------------------------------------
# Viscosity 3-D: term tables (ViscosityCGSolver3D.py:41-246 RHS, :248-456 apply)
#   row spec = (self_sphi_off, vols{name: off}, diag_weights[(mult, volname)...], terms)
#   term = (sphi_off, mult, volname, comp, vel_off, sign)
#   apply:  val -= sign*mult*scale*mu*vol*vel   if sphi >= 0   (:272-314 etc.)
#   rhs:    b   += sign*mult*scale*mu*vol*vel   if sphi <  0   (:64-106 etc.)
# --------------------------------------------------------------------------------------

_V3_ROWS = {
    # u faces, fine index (2x, 2y+1, 2z+1)   — :248-316 / :41-108
    0: dict(
        self_off=(0, 1, 1),
        vols=dict(center=(0, 1, 1), right=(1, 1, 1), left=(-1, 1, 1), top=(0, 2, 1),
                  bottom=(0, 0, 1), front=(0, 1, 2), back=(0, 1, 0)),
        diag=((2, "right"), (2, "left"), (1, "top"), (1, "bottom"), (1, "front"), (1, "back")),
        terms=(
            ((2, 1, 1), 2, "right", 0, (1, 0, 0), +1),
            ((-2, 1, 1), 2, "left", 0, (-1, 0, 0), +1),
            ((0, 3, 1), 1, "top", 0, (0, 1, 0), +1),
            ((0, -1, 1), 1, "bottom", 0, (0, -1, 0), +1),
            ((0, 1, 3), 1, "front", 0, (0, 0, 1), +1),
            ((0, 1, -1), 1, "back", 0, (0, 0, -1), +1),
            ((1, 2, 1), 1, "top", 1, (0, 1, 0), +1),
            ((-1, 2, 1), 1, "top", 1, (-1, 1, 0), -1),
            ((1, 0, 1), 1, "bottom", 1, (0, 0, 0), -1),
            ((-1, 0, 1), 1, "bottom", 1, (-1, 0, 0), +1),
            ((1, 1, 2), 1, "front", 2, (0, 0, 1), +1),
            ((-1, 1, 2), 1, "front", 2, (-1, 0, 1), -1),
            ((1, 1, 0), 1, "back", 2, (0, 0, 0), -1),
            ((-1, 1, 0), 1, "back", 2, (-1, 0, 0), +1),
        ),
    ),
    # v faces, fine index (2x+1, 2y, 2z+1)   — :318-386 / :110-177
    1: dict(
        self_off=(1, 0, 1),
        vols=dict(center=(1, 0, 1), right=(2, 0, 1), left=(0, 0, 1), top=(1, 1, 1),
                  bottom=(1, -1, 1), front=(1, 0, 2), back=(1, 0, 0)),
        diag=((1, "right"), (1, "left"), (2, "top"), (2, "bottom"), (1, "front"), (1, "back")),
        terms=(
            ((3, 0, 1), 1, "right", 1, (1, 0, 0), +1),
            ((-1, 0, 1), 1, "left", 1, (-1, 0, 0), +1),
            ((1, 2, 1), 2, "top", 1, (0, 1, 0), +1),
            ((1, -2, 1), 2, "bottom", 1, (0, -1, 0), +1),
            ((1, 0, 3), 1, "front", 1, (0, 0, 1), +1),
            ((1, 0, -1), 1, "back", 1, (0, 0, -1), +1),
            ((2, 1, 1), 1, "right", 0, (1, 0, 0), +1),
            ((2, -1, 1), 1, "right", 0, (1, -1, 0), -1),
            ((0, 1, 1), 1, "left", 0, (0, 0, 0), -1),
            ((0, -1, 1), 1, "left", 0, (0, -1, 0), +1),
            ((1, 1, 2), 1, "front", 2, (0, 0, 1), +1),
            ((1, -1, 2), 1, "front", 2, (0, -1, 1), -1),
            ((1, 1, 0), 1, "back", 2, (0, 0, 0), -1),
            ((1, -1, 0), 1, "back", 2, (0, -1, 0), +1),
        ),
    ),
    # w faces, fine index (2x+1, 2y+1, 2z)   — :388-456 / :179-246
    2: dict(
        self_off=(1, 1, 0),
        vols=dict(center=(1, 1, 0), right=(2, 1, 0), left=(0, 1, 0), top=(1, 2, 0),
                  bottom=(1, 0, 0), front=(1, 1, 1), back=(1, 1, -1)),
        diag=((1, "right"), (1, "left"), (1, "top"), (1, "bottom"), (2, "front"), (2, "back")),
        terms=(
            ((3, 1, 0), 1, "right", 2, (1, 0, 0), +1),
            ((-1, 1, 0), 1, "left", 2, (-1, 0, 0), +1),
            ((1, 3, 0), 1, "top", 2, (0, 1, 0), +1),
            ((1, -1, 0), 1, "bottom", 2, (0, -1, 0), +1),
            ((1, 1, 2), 2, "front", 2, (0, 0, 1), +1),
            ((1, 1, -2), 2, "back", 2, (0, 0, -1), +1),
            ((2, 1, 1), 1, "right", 0, (1, 0, 0), +1),
            ((2, 1, -1), 1, "right", 0, (1, 0, -1), -1),
            ((0, 1, 1), 1, "left", 0, (0, 0, 0), -1),
            ((0, 1, -1), 1, "left", 0, (0, 0, -1), +1),
            ((1, 2, 1), 1, "top", 1, (0, 1, 0), +1),
            ((1, 2, -1), 1, "top", 1, (0, 1, -1), -1),
            ((1, 0, 1), 1, "bottom", 1, (0, 0, 0), -1),
            ((1, 0, -1), 1, "bottom", 1, (0, 0, -1), +1),
        ),
    ),
}

# Viscosity 2-D (ViscosityCGSolver2D.py:6-102 RHS, :105-206 apply); same table format.
_V2_ROWS = {
    0: dict(  # u faces (2x, 2y+1)  — :105-153 / :6-55
        self_off=(0, 1),
        vols=dict(center=(0, 1), right=(1, 1), left=(-1, 1), top=(0, 2), bottom=(0, 0)),
        diag=((2, "right"), (2, "left"), (1, "top"), (1, "bottom")),
        terms=(
            ((2, 1), 2, "right", 0, (1, 0), +1),
            ((-2, 1), 2, "left", 0, (-1, 0), +1),
            ((0, 3), 1, "top", 0, (0, 1), +1),
            ((0, -1), 1, "bottom", 0, (0, -1), +1),
            ((1, 2), 1, "top", 1, (0, 1), +1),
            ((-1, 2), 1, "top", 1, (-1, 1), -1),
            ((1, 0), 1, "bottom", 1, (0, 0), -1),
            ((-1, 0), 1, "bottom", 1, (-1, 0), +1),
        ),
    ),
    1: dict(  # v faces (2x+1, 2y)  — :156-206 / :57-102
        self_off=(1, 0),
        vols=dict(center=(1, 0), right=(2, 0), left=(0, 0), top=(1, 1), bottom=(1, -1)),
        diag=((1, "right"), (1, "left"), (2, "top"), (2, "bottom")),
        terms=(
            ((3, 0), 1, "right", 1, (1, 0), +1),
            ((-1, 0), 1, "left", 1, (-1, 0), +1),
            ((1, 2), 2, "top", 1, (0, 1), +1),
            ((1, -2), 2, "bottom", 1, (0, -1), +1),
            ((2, 1), 1, "right", 0, (1, 0), +1),
            ((2, -1), 1, "right", 0, (1, -1), -1),
            ((0, 1), 1, "left", 0, (0, 0), -1),
            ((0, -1), 1, "left", 0, (0, -1), +1),
        ),
    ),
}


def _fluid3(s):  # 3-D: fluid iff sphi >= 0, solid iff sphi < 0 (ViscosityCGSolver3D.py:255,272)
    return s >= 0


def _fluid2(s):  # 2-D: fluid iff sphi > 0, solid iff sphi <= 0 (ViscosityCGSolver2D.py:112,129)
    return s > 0


def _visc_row_apply(row, fluid, scale, mu, vel, out, sphi, vol):
    osh = out.shape
    if min(osh) < 3:
        return
    V = {k: _fine(vol, o, osh) for k, o in row["vols"].items()}
    # diag = vol_center + scale*mu*(w0*v0 + w1*v1 + ...)   left-to-right   (:268, :338, :408)
    acc = None
    for mult, name in row["diag"]:
        t = 2 * V[name] if mult == 2 else V[name]
        acc = t if acc is None else acc + t
    diag = V["center"] + scale * mu * acc
    own = row_comp(row)
    val = diag * _coarse(vel[own], (0,) * len(osh), osh)
    for soff, mult, vname, comp, voff, sign in row["terms"]:
        coef = (2 * scale * mu) if mult == 2 else (scale * mu)
        t = coef * V[vname] * _coarse(vel[comp], voff, osh)
        t = np.where(fluid(_fine(sphi, soff, osh)), t, 0.0)
        val = val - t if sign > 0 else val + t
    rowfluid = fluid(_fine(sphi, row["self_off"], osh))
    out[_interior(osh)] = np.where(rowfluid, val, 0.0)


def _visc_row_rhs(row, fluid, scale, mu, vel, b, sphi, vol):
    osh = b.shape
    if min(osh) < 3:
        return
    V = {k: _fine(vol, o, osh) for k, o in row["vols"].items()}
    own = row_comp(row)
    val = _coarse(vel[own], (0,) * len(osh), osh) * V["center"]
    for soff, mult, vname, comp, voff, sign in row["terms"]:
        coef = (2 * scale * mu) if mult == 2 else (scale * mu)
        t = coef * V[vname] * _coarse(vel[comp], voff, osh)
        t = np.where(~fluid(_fine(sphi, soff, osh)), t, 0.0)
        val = val + t if sign > 0 else val - t
    rowfluid = fluid(_fine(sphi, row["self_off"], osh))
    b[_interior(osh)] = np.where(rowfluid, val, 0.0)


def row_comp(row):
    """component index of the row = position of the even entry of its fine-grid parity."""
    return row["self_off"].index(0)


# ------------------------------- 3-D viscosity public API ------------------------------


def visc3d_matvecmul(gres, scale, mu, vx, vy, vz, out_x, out_y, out_z, sphi, vol):
    """q = A d.  ViscosityCGSolver3D.py:515-524 (launcher), :248-456 (kernels)."""
    vel = (vx, vy, vz)
    for c, out in enumerate((out_x, out_y, out_z)):
        _visc_row_apply(_V3_ROWS[c], _fluid3, scale, mu, vel, out, sphi, vol)


def visc3d_initialize_solver(gres, scale, mu, vx, vy, vz, sphi, sv, vol, b_x, b_y, b_z):
    """RHS b.  ViscosityCGSolver3D.py:504-513, :41-246.  ``sv`` is unused (all uses commented out)."""
    vel = (vx, vy, vz)
    for c, b in enumerate((b_x, b_y, b_z)):
        _visc_row_rhs(_V3_ROWS[c], _fluid3, scale, mu, vel, b, sphi, vol)


def _extrapolate_sweep(v, valid):
    """One Jacobi sweep of ViscosityCGSolver3D.py:8-39 -> (new_v, new_valid)."""
    new_v = v.copy()
    new_valid = valid.copy()
    osh = v.shape
    if min(osh) < 3:
        return new_v, new_valid
    val = np.zeros(tuple(n - 2 for n in osh), dtype=F64)
    cnt = np.zeros(val.shape, dtype=np.int64)
    # neighbour order +x,-x,+y,-y,+z,-z (:19-36); sequential accumulation
    for off in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)):
        m = _coarse(valid, off, osh)
        val = val + np.where(m, _coarse(v, off, osh), 0.0)
        cnt = cnt + m
    me_valid = valid[_interior(osh)]
    upd = (~me_valid) & (cnt > 0)
    with np.errstate(invalid="ignore", divide="ignore"):
        filled = val / cnt
    new_v[_interior(osh)] = np.where(upd, filled, v[_interior(osh)])
    new_valid[_interior(osh)] = me_valid | upd
    return new_v, new_valid


def visc3d_validity(sphi):
    """ViscosityCGSolver3D.py:479-481."""
    return (sphi[0::2, 1::2, 1::2] >= 0, sphi[1::2, 0::2, 1::2] >= 0, sphi[1::2, 1::2, 0::2] >= 0)


def visc3d_extrapolate(gres, num_iter, vx, vy, vz, sphi):
    """In-place 3-sweep extrapolation.  ViscosityCGSolver3D.py:472-502."""
    valids = list(visc3d_validity(sphi))
    vs = [vx, vy, vz]
    for _ in range(num_iter):
        for c in range(3):
            nv, nval = _extrapolate_sweep(vs[c], valids[c])
            vs[c][...] = nv
            valids[c] = nval
    return valids


def visc3d_apply_viscosity(gres, vx, vy, vz, out_x, out_y, out_z, sphi, sv):
    """Masked write-back, indices 1..g-1 on all axes for all components.  :458-470, :526-530."""
    g = [int(n) for n in gres]
    sl = tuple(slice(1, n) for n in g)
    for arr, out, off in ((vx, out_x, (0, 1, 1)), (vy, out_y, (1, 0, 1)), (vz, out_z, (1, 1, 0))):
        fs = tuple(slice(2 + o, 2 * (n - 1) + o + 1, 2) for o, n in zip(off, g))
        m = sphi[fs] >= 0
        tgt = arr[sl]
        tgt[m] = out[sl][m].astype(arr.dtype)


class CGTrace:
    """What a solve did: iteration count and the delta history (delta[0] = initial)."""

    def __init__(self):
        self.iterations = 0
        self.deltas = []


def _cg(apply_fn, xs, bs, tol, max_iter, raise_on_fail=True, trace=None):
    """Plain CG exactly as ViscosityCGSolver3D.py:575-612 / PressureCGSolver3D.py:201-223.

    xs, bs: lists of arrays (components).  Returns (delta, alpha, beta, iterations).
    """
    qs = [np.zeros_like(x) for x in xs]
    apply_fn(xs, qs)
    ds = [b - q for b, q in zip(bs, qs)]
    rs = [d.copy() for d in ds]
    delta = _sumsq(rs)
    alpha = beta = 0.0
    iters = 0
    if trace is not None:
        trace.deltas.append(delta)
    if not delta < tol ** 2:
        converged = False
        for _ in range(int(max_iter)):
            apply_fn(ds, qs)
            dq = _dot(ds, qs)
            alpha = delta / dq
            for x, d in zip(xs, ds):
                x += alpha * d
            for r, q in zip(rs, qs):
                r -= alpha * q
            old_delta = delta
            delta = _sumsq(rs)
            iters += 1
            if trace is not None:
                trace.deltas.append(delta)
            if delta < tol ** 2:
                converged = True
                break
            beta = delta / old_delta
            for d, r in zip(ds, rs):
                d[...] = r + beta * d
        if not converged and raise_on_fail:
            if trace is not None:
                trace.iterations = iters
            raise ValueError("Failed to converge!")
    if trace is not None:
        trace.iterations = iters
    return delta, alpha, beta, iters, (ds, rs, qs)


def _sumsq(rs):
    # (cp.sum(r_x**2) + cp.sum(r_y**2) + cp.sum(r_z**2)).item()   (:585)
    tot = None
    for r in rs:
        s = np.sum(r ** 2)
        tot = s if tot is None else tot + s
    return float(tot)


def _dot(ds, qs):
    tot = None
    for d, q in zip(ds, qs):
        s = np.sum(d * q)
        tot = s if tot is None else tot + s
    return float(tot)


class ViscosityCGSolver3D:
    """ViscosityCGSolver3D.py:532-613."""

    def __init__(self, gres, bound_size):
        self.gres = np.asarray(gres, dtype=np.int64)
        self.cell_size = np.asarray(bound_size) / self.gres
        self.cell_vol = float(np.prod(self.cell_size))
        g = self.gres
        self.vol = np.zeros(tuple(2 * g + 1), dtype=F64)
        self.x_x = np.zeros((g[0] + 1, g[1], g[2]), dtype=F64)
        self.x_y = np.zeros((g[0], g[1] + 1, g[2]), dtype=F64)
        self.x_z = np.zeros((g[0], g[1], g[2] + 1), dtype=F64)
        self.b_x, self.b_y, self.b_z = (np.zeros_like(a) for a in (self.x_x, self.x_y, self.x_z))
        self.alpha = self.beta = self.delta = 0.0
        self.max_iter = int(np.prod(g))
        self.trace = CGTrace()

    def solve(self, dt, mu, rho, vx, vy, vz, sphi, sv, lphi, lvol, tol=1e-3):
        scale = dt / self.cell_vol / rho                       # :567
        self.vol[:] = lvol / (self.cell_vol * 0.125)           # :568
        self.x_x[:] = vx
        self.x_y[:] = vy
        self.x_z[:] = vz
        visc3d_extrapolate(self.gres, 3, self.x_x, self.x_y, self.x_z, sphi)        # :573
        visc3d_initialize_solver(self.gres, scale, mu, self.x_x, self.x_y, self.x_z, sphi, sv,
                                 self.vol, self.b_x, self.b_y, self.b_z)            # :574

        def A(vs, qs):
            visc3d_matvecmul(self.gres, scale, mu, vs[0], vs[1], vs[2], qs[0], qs[1], qs[2], sphi, self.vol)

        self.trace = CGTrace()
        try:
            self.delta, self.alpha, self.beta, _, _ = _cg(
                A, [self.x_x, self.x_y, self.x_z], [self.b_x, self.b_y, self.b_z], tol, self.max_iter,
                trace=self.trace)
        finally:
            pass
        visc3d_apply_viscosity(self.gres, vx, vy, vz, self.x_x, self.x_y, self.x_z, sphi, sv)  # :613


# ------------------------------- 2-D viscosity public API ------------------------------


def visc2d_matvecmul(gres, scale, mu, vx, vy, out_x, out_y, sphi, vol):
    """ViscosityCGSolver2D.py:231-238, :105-206."""
    vel = (vx, vy)
    for c, out in enumerate((out_x, out_y)):
        _visc_row_apply(_V2_ROWS[c], _fluid2, scale, mu, vel, out, sphi, vol)


def visc2d_initialize_solver(gres, scale, mu, vx, vy, sphi, sv, vol, b_x, b_y):
    """ViscosityCGSolver2D.py:222-229, :6-102."""
    vel = (vx, vy)
    for c, b in enumerate((b_x, b_y)):
        _visc_row_rhs(_V2_ROWS[c], _fluid2, scale, mu, vel, b, sphi, vol)


def visc2d_apply_viscosity(gres, vx, vy, out_x, out_y, sphi, sv):
    """ViscosityCGSolver2D.py:209-219, :240-244 (fluid test is ``> 0``)."""
    g = [int(n) for n in gres]
    sl = tuple(slice(1, n) for n in g)
    for arr, out, off in ((vx, out_x, (0, 1)), (vy, out_y, (1, 0))):
        fs = tuple(slice(2 + o, 2 * (n - 1) + o + 1, 2) for o, n in zip(off, g))
        m = sphi[fs] > 0
        tgt = arr[sl]
        tgt[m] = out[sl][m].astype(arr.dtype)


class ViscosityCGSolver2D:
    """ViscosityCGSolver2D.py:246-318 (no extrapolation; default tol 1e-4)."""

    def __init__(self, gres, bound_size):
        self.gres = np.asarray(gres, dtype=np.int64)
        self.cell_size = np.asarray(bound_size) / self.gres
        self.cell_vol = float(np.prod(self.cell_size))
        g = self.gres
        self.vol = np.zeros(tuple(2 * g + 1), dtype=F64)
        self.x_x = np.zeros((g[0] + 1, g[1]), dtype=F64)
        self.x_y = np.zeros((g[0], g[1] + 1), dtype=F64)
        self.b_x, self.b_y = np.zeros_like(self.x_x), np.zeros_like(self.x_y)
        self.alpha = self.beta = self.delta = 0.0
        self.max_iter = int(np.prod(g))
        self.trace = CGTrace()

    def solve(self, dt, mu, rho, vx, vy, sphi, sv, lphi, lvol, tol=1e-4, save=False):
        scale = dt / self.cell_vol / rho
        self.vol[:] = lvol / (self.cell_vol * 0.125)
        self.x_x[:] = vx
        self.x_y[:] = vy
        visc2d_initialize_solver(self.gres, scale, mu, self.x_x, self.x_y, sphi, sv, self.vol, self.b_x, self.b_y)

        def A(vs, qs):
            visc2d_matvecmul(self.gres, scale, mu, vs[0], vs[1], qs[0], qs[1], sphi, self.vol)

        self.trace = CGTrace()
        self.delta, self.alpha, self.beta, _, _ = _cg(
            A, [self.x_x, self.x_y], [self.b_x, self.b_y], tol, self.max_iter, trace=self.trace)
        visc2d_apply_viscosity(self.gres, vx, vy, self.x_x, self.x_y, sphi, sv)


# --------------------------------------------------------------------------------------
# Solid fractions (SolidFractionCommon.py:4-60, SolidFraction3D.py:6-32, SolidFraction2D.py:6-26)
# --------------------------------------------------------------------------------------


def edge_in_fraction(lval, rval):
    """SolidFractionCommon.py:4-16, vectorised; returns float64."""
    lval = np.asarray(lval, dtype=F64)
    rval = np.asarray(rval, dtype=F64)
    l_in = lval < 0
    r_in = rval < 0
    diff = -np.abs(lval - rval)
    with np.errstate(invalid="ignore", divide="ignore"):
        lfrac = lval / diff
        rfrac = rval / diff
    out = np.where(l_in & r_in, 1.0, np.where(~l_in & ~r_in, 0.0, np.where(l_in, lfrac, rfrac)))
    return out


def _tri_all_in(a, b, c):
    # tri_in_fraction (:18-50) evaluates edge_in_fraction on the two SAME-sign vertices in the
    # 1-in and 2-in cases, which yields 0 in both; only the 3-in case contributes (1.0).
    return ((a < 0) & (b < 0) & (c < 0)).astype(F64)


def face_in_fraction(bl, br, tl, tr):
    """SolidFractionCommon.py:52-60."""
    ce = 0.25 * (bl + br + tl + tr)
    return 0.25 * (_tri_all_in(bl, br, ce) + _tri_all_in(br, tr, ce) + _tri_all_in(tr, tl, ce) + _tri_all_in(tl, bl, ce))


def solidfrac3d(gres, sphi, wx, wy, wz):
    """SolidFraction3D.py:6-32: low-side faces only; far planes are left untouched."""
    nx, ny, nz = (int(n) for n in gres)
    N = sphi[0::2, 0::2, 0::2]
    blb = N[0:nx, 0:ny, 0:nz]
    brb = N[1:nx + 1, 0:ny, 0:nz]
    tlb = N[0:nx, 1:ny + 1, 0:nz]
    trb = N[1:nx + 1, 1:ny + 1, 0:nz]
    blf = N[0:nx, 0:ny, 1:nz + 1]
    brf = N[1:nx + 1, 0:ny, 1:nz + 1]
    tlf = N[0:nx, 1:ny + 1, 1:nz + 1]
    wx[0:nx, 0:ny, 0:nz] = 1.0 - face_in_fraction(tlb, blb, tlf, blf)   # :22
    wy[0:nx, 0:ny, 0:nz] = 1.0 - face_in_fraction(brb, blb, brf, blf)   # :24
    wz[0:nx, 0:ny, 0:nz] = 1.0 - face_in_fraction(trb, tlb, brb, blb)   # :26


def solidfrac2d(gres, sphi, wx, wy):
    """SolidFraction2D.py:6-26: threads x<=W-2, y<=H-2 write wx[x],wx[x+1],wy[.,y],wy[.,y+1]."""
    W, H = (int(n) for n in gres)
    if W < 2 or H < 2:
        return
    N = sphi[0::2, 0::2]
    # wx[X, y], X in 0..W-1, y in 0..H-2 :  1 - eif(N[X, y+1], N[X, y])     (:17-18)
    wx[0:W, 0:H - 1] = 1.0 - edge_in_fraction(N[0:W, 1:H], N[0:W, 0:H - 1])
    # wy[x, Y], x in 0..W-2, Y in 0..H-1 :  1 - eif(N[x+1, Y], N[x, Y])     (:19-20)
    wy[0:W - 1, 0:H] = 1.0 - edge_in_fraction(N[1:W, 0:H], N[0:W - 1, 0:H])


# --------------------------------------------------------------------------------------
# Pressure (PressureCGSolver3D.py:6-226, PressureCGSolver2D.py:6-179) — dimension-generic
# --------------------------------------------------------------------------------------


def _unit(d, a):
    return tuple(1 if i == a else 0 for i in range(d))


def _face_centre_off(d, a):
    """fine-grid offset of the low face of cell along axis a: even on a, odd elsewhere."""
    return tuple(0 if i == a else 1 for i in range(d))


def press_initialize_solver(cell_size, gres, vel, sphi, sv, lphi, b, ws):
    """Weighted divergence RHS.  PressureCGSolver3D.py:6-50 / PressureCGSolver2D.py:6-44.

    vel, ws: tuples of per-axis face arrays.  Axis order +x,-x,+y,-y,(+z,-z), sequential sums.
    """
    g = tuple(int(n) for n in gres)
    d = len(g)
    if min(g) < 3:
        return
    cs = np.asarray(cell_size, dtype=F64).reshape(-1)
    if cs.size == 1:
        cs = np.repeat(cs, d)
    val = np.zeros(tuple(n - 2 for n in g), dtype=F64)
    zero = (0,) * d
    for a in range(d):
        e = _unit(d, a)
        fo = _face_centre_off(d, a)
        # + face  (:22-24)
        w = _coarse(ws[a], e, g)
        v = _coarse(vel[a], e, g)
        s = _fine(sv[..., a], tuple(f + 2 * ee for f, ee in zip(fo, e)), g)
        val = val + w * v / cs[a]
        val = val - np.where(w < 1, w * s / cs[a], 0.0)
        # - face  (:27-29)
        w = _coarse(ws[a], zero, g)
        v = _coarse(vel[a], zero, g)
        s = _fine(sv[..., a], fo, g)
        val = val - w * v / cs[a]
        val = val + np.where(w < 1, w * s / cs[a], 0.0)
    fluid = lphi[_interior(g)] < 0
    b[_interior(g)] = np.where(fluid, val, 0.0)


def press_matvecmul(gres, v, out, ws, lphi):
    """Ghost-fluid variable-coefficient Poisson apply.  PressureCGSolver3D.py:52-130 / 2D :46-100."""
    g = tuple(int(n) for n in gres)
    d = len(g)
    if min(g) < 3:
        return
    zero = (0,) * d
    phi = lphi[_interior(g)]
    val = np.zeros(phi.shape, dtype=F64)
    diag = np.zeros(phi.shape, dtype=F64)
    for a in range(d):
        e = _unit(d, a)
        for sgn in (+1, -1):
            noff = tuple(sgn * ee for ee in e)
            nphi = _coarse(lphi, noff, g)
            w = _coarse(ws[a], e if sgn > 0 else zero, g)
            nfluid = nphi < 0
            val = val - np.where(nfluid, w * _coarse(v, noff, g), 0.0)
            with np.errstate(invalid="ignore", divide="ignore"):
                frac = np.minimum(1.0, np.maximum(0.01, phi / (phi - nphi)))
                diag = diag + np.where(nfluid, w, w / frac)
    val = val + diag * v[_interior(g)]
    out[_interior(g)] = np.where(phi < 0, val, 0.0)


def press_apply_pressure(gres, cell_size, vel, pv, ws, sv, lphi):
    """Velocity update + solid blend, in place.  PressureCGSolver3D.py:132-153 / 2D :102-120."""
    g = tuple(int(n) for n in gres)
    d = len(g)
    cs = np.asarray(cell_size, dtype=F64).reshape(-1)
    if cs.size == 1:
        cs = np.repeat(cs, d)
    sl = tuple(slice(1, n) for n in g)
    for a in range(d):
        e = _unit(d, a)
        slm = tuple(slice(1 - ee, n - ee) for ee, n in zip(e, g))
        phi_c = lphi[sl]
        phi_m = lphi[slm]
        act = (phi_c < 0) | (phi_m < 0)
        theta = np.minimum(1.0, np.maximum(0.01, edge_in_fraction(phi_c, phi_m)))
        fo = _face_centre_off(d, a)
        fs = tuple(slice(2 + o, 2 * (n - 1) + o + 1, 2) for o, n in zip(fo, g))
        s = sv[..., a][fs]
        w = ws[a][sl]
        old = vel[a][sl]
        new = old + (pv[sl] - pv[slm]) * cs[a] / theta
        new = w * new + (1 - w) * s
        tgt = vel[a][sl]
        tgt[act] = new[act].astype(vel[a].dtype)


class CGSolverBuffer:
    """CGSolverBuffer.py:3-8."""

    def __init__(self, gres):
        g = tuple(int(n) for n in np.asarray(gres))
        self.d = np.zeros(g, dtype=F64)
        self.r = np.zeros(g, dtype=F64)
        self.q = np.zeros(g, dtype=F64)
        self.b = np.zeros(g, dtype=F64)


class _PressureCGSolver:
    raise_on_fail = True

    def __init__(self, buf, gres, bound_size):
        self.gres = np.asarray(gres, dtype=np.int64)
        self.cell_size = np.asarray(bound_size) / self.gres        # :176 (scalar GDX allowed)
        self.buf = buf
        g = tuple(int(n) for n in self.gres)
        d = len(g)
        self.x = np.zeros(g, dtype=F64)
        self.ws = [np.zeros(tuple(n + (1 if i == a else 0) for i, n in enumerate(g)), dtype=F64) for a in range(d)]
        self.alpha = self.beta = self.delta = 0.0
        self.max_iter = int(np.prod(self.gres))
        self.trace = CGTrace()

    def _solve(self, vel, sphi, sv, lphi, ws, tol):
        if ws is None or any(w is None for w in ws):
            if len(vel) == 3:
                solidfrac3d(self.gres, sphi, *self.ws)
            else:
                solidfrac2d(self.gres, sphi, *self.ws)
            ws = self.ws
        self.x *= 0.0                                                # :198
        press_initialize_solver(self.cell_size, self.gres, vel, sphi, sv, lphi, self.buf.b, ws)

        def A(vs, qs):
            press_matvecmul(self.gres, vs[0], qs[0], ws, lphi)

        self.trace = CGTrace()
        self.delta, self.alpha, self.beta, _, (ds, rs, qs) = _cg(
            A, [self.x], [self.buf.b], tol, self.max_iter, raise_on_fail=self.raise_on_fail, trace=self.trace)
        self.buf.d[...] = ds[0]
        self.buf.r[...] = rs[0]
        self.buf.q[...] = qs[0]
        press_apply_pressure(self.gres, self.cell_size, vel, self.x, ws, sv, lphi)   # :226


class PressureCGSolver3D(_PressureCGSolver):
    """PressureCGSolver3D.py:173-226."""

    @property
    def wx(self):
        return self.ws[0]

    @property
    def wy(self):
        return self.ws[1]

    @property
    def wz(self):
        return self.ws[2]

    def solve(self, vx, vy, vz, sphi, sv, lphi, wx=None, wy=None, wz=None, tol=1e-3):
        ws = None if (wx is None or wy is None or wz is None) else (wx, wy, wz)
        self._solve((vx, vy, vz), sphi, sv, lphi, ws, tol)


class PressureCGSolver2D(_PressureCGSolver):
    """PressureCGSolver2D.py:140-179 — the CG loop has no ``else: raise`` (:165-177)."""

    raise_on_fail = False

    @property
    def wx(self):
        return self.ws[0]

    @property
    def wy(self):
        return self.ws[1]

    def solve(self, vx, vy, sphi, sv, lphi, wx=None, wy=None, tol=1e-3):
        ws = None if (wx is None or wy is None) else (wx, wy)
        self._solve((vx, vy), sphi, sv, lphi, ws, tol)


# --------------------------------------------------------------------------------------
# Density (volume-conservation) solver — DensityCGSolver3D.py:8-350   (§8 "next" row f-1)
# --------------------------------------------------------------------------------------


def _trilinear(px, bound_min, cell_size, bias, shape):
    """Indices and weights of the 8 grid points around each particle (DensityCGSolver3D.py:14-34 / :218-238).

    gi = floor((x - bound_min)/cell - bias); w = |gx - x|/cell with gx = (gi + bias)*cell + bound_min;
    corner (ix,iy,iz): index clamped to [0, shape-1], weight prod_d (i_d + (-1)^i_d (1 - w_d)).
    """
    px = np.asarray(px, dtype=F64)
    bm = np.asarray(bound_min, dtype=F64)
    cs = np.asarray(cell_size, dtype=F64)
    bias = np.asarray(bias, dtype=F64)
    gi = np.floor((px - bm) / cs - bias).astype(np.int64)
    gx = (gi + bias) * cs + bm
    w = np.abs(gx - px) / cs
    out = []
    for ix in (0, 1):
        for iy in (0, 1):
            for iz in (0, 1):
                off = (ix, iy, iz)
                idx = tuple(np.clip(gi[:, d] + off[d], 0, shape[d] - 1) for d in range(3))
                ww = [off[d] + ((-1) ** off[d]) * (1 - w[:, d]) for d in range(3)]
                out.append((idx, ww[0] * ww[1] * ww[2]))
    return out


def density_initialize_density(bound_min, cell_size, gres, px, pm, pvol, gm, gvol):
    """Particle -> cell scatter of mass and volume (:8-36).  Accumulation order is particle-major as in a sequential run."""
    g = tuple(int(n) for n in gres)
    for idx, weight in _trilinear(px, bound_min, cell_size, (0.5, 0.5, 0.5), g):
        np.add.at(gm, idx, weight * np.asarray(pm, dtype=F64))
        np.add.at(gvol, idx, weight * float(pvol))


def density_fix_volume(cell_size, gres, lvol, gvol, sphi, lphi, ws):
    """:38-92 — interior cells: full cells deep inside the liquid and away from solids count as cvol; cap by the open fraction."""
    g = tuple(int(n) for n in gres)
    if min(g) < 3:
        return
    cs = np.asarray(cell_size, dtype=F64)
    cvol = float(np.prod(cs))
    dx = float(np.min(cs))
    I = _interior(g)
    fluid_vol = gvol[I].copy()
    near_solid = _fine(sphi, (1, 1, 1), g) < dx
    inside = lphi[I] < 0
    for a in range(3):
        for sgn in (1, -1):
            off = tuple(sgn if k == a else 0 for k in range(3))
            inside = inside & (_coarse(lphi, off, g) < 0)
    fluid_vol = np.where(inside & ~near_solid, cvol, fluid_vol)
    frac = _density_open_fraction(g, ws)
    gvol[I] = np.minimum(fluid_vol, cvol * frac)


def _density_open_fraction(g, ws):
    zero = (0, 0, 0)
    s = _coarse(ws[0], zero, g) + _coarse(ws[0], (1, 0, 0), g)
    s = s + _coarse(ws[1], zero, g)
    s = s + _coarse(ws[1], (0, 1, 0), g)
    s = s + _coarse(ws[2], zero, g)
    s = s + _coarse(ws[2], (0, 0, 1), g)
    return s / 6


def density_initialize_solver(rho0, dt, gres, cell_size, gm, gvol, lphi, ws, b):
    """:94-125 — b = (1 - clamp(density/rho0, 0.5, 1.5)) / dt on interior fluid cells."""
    g = tuple(int(n) for n in gres)
    if min(g) < 3:
        return
    cvol = float(np.prod(np.asarray(cell_size, dtype=F64)))
    I = _interior(g)
    frac = _density_open_fraction(g, ws)
    solid_vol = (1 - frac) * cvol
    solid_mass = rho0 * solid_vol
    cell_mass = gm[I] + solid_mass
    cell_vol = gvol[I] + solid_vol
    dens = cell_mass / np.maximum(cell_vol, 1e-10) / rho0
    dens = np.where(cell_mass < 1e-10, 1.0, dens)
    dens = np.maximum(0.5, np.minimum(1.5, dens))
    b[I] = np.where(lphi[I] < 0, (1 - dens) / dt, 0.0)


def density_matvecmul(gres, v, out, ws, lphi):
    """:127-204 — like the pressure apply but with UNIT diagonal weights, and the -z off-diagonal term reads
    wz[x,y,z+1] (not wz[x,y,z]) exactly as the reference does (:196)."""
    g = tuple(int(n) for n in gres)
    if min(g) < 3:
        return
    zero = (0, 0, 0)
    phi = lphi[_interior(g)]
    val = np.zeros(phi.shape, dtype=F64)
    diag = np.zeros(phi.shape, dtype=F64)
    for a in range(3):
        e = _unit(3, a)
        for sgn in (+1, -1):
            noff = tuple(sgn * ee for ee in e)
            nphi = _coarse(lphi, noff, g)
            woff = e if (sgn > 0 or a == 2) else zero
            w = _coarse(ws[a], woff, g)
            nfluid = nphi < 0
            val = val - np.where(nfluid, w * _coarse(v, noff, g), 0.0)
            with np.errstate(invalid="ignore", divide="ignore"):
                frac = np.minimum(1.0, np.maximum(0.01, phi / (phi - nphi)))
                diag = diag + np.where(nfluid, 1.0, 1.0 / frac)
    val = val + diag * v[_interior(g)]
    out[_interior(g)] = np.where(phi < 0, val, 0.0)


def density_compute_displacement(gres, dt, cell_size, disp, pv, lphi):
    """:206-219 — indices 1..g-1 on all axes; note: NOT restricted to faces next to liquid."""
    g = tuple(int(n) for n in gres)
    cs = np.asarray(cell_size, dtype=F64)
    sl = tuple(slice(1, n) for n in g)
    for a in range(3):
        slm = tuple(slice(1 - (1 if k == a else 0), n - (1 if k == a else 0)) for k, n in enumerate(g))
        theta = np.minimum(1.0, np.maximum(0.01, edge_in_fraction(lphi[sl], lphi[slm])))
        disp[a][sl] = (pv[sl] - pv[slm]) * dt * cs[a] / theta


def density_apply_displacement(px, d_arr, bound_min, cell_size, grid_bias, axis):
    """:221-248 — trilinear gather of one displacement component onto the particles (in place on px[:, axis])."""
    for idx, weight in _trilinear(px.copy(), bound_min, cell_size, grid_bias, d_arr.shape):
        px[:, axis] += weight * d_arr[idx]


class DensityCGSolver3D:
    """DensityCGSolver3D.py:283-350."""

    def __init__(self, buf, gres, bound_min, bound_size):
        self.gres = np.asarray(gres, dtype=np.int64)
        self.bound_min = np.asarray(bound_min, dtype=F64)
        self.cell_size = np.asarray(bound_size, dtype=F64) / self.gres
        g = tuple(int(n) for n in self.gres)
        self.buf = buf
        self.m = np.zeros(g, dtype=F64)
        self.vol = np.zeros(g, dtype=F64)
        self.x = np.zeros(g, dtype=F64)
        self.ws = [np.zeros(tuple(n + (1 if i == a else 0) for i, n in enumerate(g)), dtype=F64) for a in range(3)]
        self.disp = [np.zeros(w.shape, dtype=F64) for w in self.ws]
        self.alpha = self.beta = self.delta = 0.0
        self.max_iter = int(np.prod(self.gres))
        self.trace = CGTrace()

    wx = property(lambda self: self.ws[0])
    wy = property(lambda self: self.ws[1])
    wz = property(lambda self: self.ws[2])
    dx = property(lambda self: self.disp[0])
    dy = property(lambda self: self.disp[1])
    dz = property(lambda self: self.disp[2])

    def solve(self, rho0, dt, px, pm, pvol, vx, vy, vz, sphi, sv, lphi, lvol, wx=None, wy=None, wz=None, tol=1e-3):
        if wx is None or wy is None or wz is None:
            solidfrac3d(self.gres, sphi, *self.ws)
            ws = self.ws
        else:
            ws = (wx, wy, wz)
        self.m *= 0
        self.vol *= 0
        self.x *= 0
        density_initialize_density(self.bound_min, self.cell_size, self.gres, px, pm, pvol, self.m, self.vol)
        density_fix_volume(self.cell_size, self.gres, lvol, self.vol, sphi, lphi, ws)
        density_initialize_solver(rho0, dt, self.gres, self.cell_size, self.m, self.vol, lphi, ws, self.buf.b)

        def A(vs, qs):
            density_matvecmul(self.gres, vs[0], qs[0], ws, lphi)

        self.trace = CGTrace()
        self.delta, self.alpha, self.beta, _, (ds, rs, qs) = _cg(A, [self.x], [self.buf.b], tol, self.max_iter, trace=self.trace)
        self.buf.d[...] = ds[0]
        self.buf.r[...] = rs[0]
        self.buf.q[...] = qs[0]
        density_compute_displacement(self.gres, dt, self.cell_size, self.disp, self.x, lphi)
        bias = ((0, 0.5, 0.5), (0.5, 0, 0.5), (0.5, 0.5, 0))
        for a in range(3):
            density_apply_displacement(px, self.disp[a], self.bound_min, self.cell_size, bias[a], a)
