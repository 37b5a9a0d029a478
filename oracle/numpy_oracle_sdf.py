"""TEST INFRASTRUCTURE ONLY — NumPy restatement of the reference's rigid-body SDF (solver/sdf3D.py): ``evaluate`` (:218-242,
:255-266) and ``project`` (:244-253, :268-273) with the shape functions they call (sphere :52-84, box :86-152, cylinder
:154-226).  Vectorised over positions, same association as the reference's scalar code so that it reproduces the fixtures the
unmodified reference produced under Numba's simulator (tests/golden/sdf3d_bodies.npz) to the last bit for boxes and spheres.
Quirks kept: a flipped box always clamps and re-transforms (``~(in_out)`` is never 0); ``min_sd`` starts at 100.  Repaired:
``cylinder_eval``'s unassigned ``y_clip`` between the end planes (taken as y, like cylinder_project)."""
import numpy as np


def _to_body(rb, p):
    T, R = rb[1:5], rb[5:9]
    q = np.empty_like(p)
    for i in range(3):
        t = 0.0
        for j in range(3):
            t = t - R[j, i] * T[j, 3]
        acc = np.zeros(p.shape[0])
        for j in range(3):
            acc = acc + R[j, i] * p[:, j]
        q[:, i] = acc + t
    return q


def _to_world(rb, q):
    T, R = rb[1:5], rb[5:9]
    p = np.empty_like(q)
    for i in range(3):
        acc = np.zeros(q.shape[0])
        for j in range(3):
            acc = acc + R[i, j] * q[:, j]
        p[:, i] = acc + T[i, 3]
    return p


def _flipped(rb):
    return rb[0, 0] % 2 != 0


def sphere_sd(rb, p):
    d = p - rb[1:4, 3]
    sd = np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2 + d[:, 2] ** 2) - rb[0, 1]
    return -sd if _flipped(rb) else sd


def box_sd(rb, p):
    q = _to_body(rb, p)
    tmp = np.zeros(p.shape[0])
    mx = np.full(p.shape[0], -100.0)
    for i in range(3):
        d = np.abs(q[:, i]) - rb[0, 1 + i] / 2
        tmp = tmp + np.where(d > 0, d * d, 0.0)
        mx = np.where(mx < d, d, mx)
    sd = np.sqrt(tmp)
    sd = np.where(mx < 0, sd + mx, sd)
    return -sd if _flipped(rb) else sd


def cylinder_sd(rb, p):
    q = _to_body(rb, p)
    hh = rb[0, 2] / 2
    y_clip = np.clip(q[:, 1], -hh, hh)
    sd = np.sqrt(q[:, 0] ** 2 + q[:, 2] ** 2) - rb[0, 1]
    capped = (y_clip == hh) | (y_clip == -hh)
    dy = np.abs(y_clip - q[:, 1])
    inside = np.maximum(sd, np.maximum(q[:, 1] - hh, -(q[:, 1] + hh)))
    out = np.where(sd < 0, np.where(capped, dy, inside), np.where(capped, np.sqrt(sd ** 2 + dy ** 2), sd))
    return -out if _flipped(rb) else out


_SD = {0: sphere_sd, 1: box_sd, 2: cylinder_sd}


def evaluate(rb_d, position):
    """returns (sd, vel) for positions of any leading shape"""
    p = np.asarray(position, dtype=np.float64).reshape(-1, 3)
    min_sd = np.full(p.shape[0], 100.0)
    idx = np.zeros(p.shape[0], dtype=np.int64)
    for i in range(rb_d.shape[0]):
        d = _SD[int(rb_d[i, 0, 0] // 2)](rb_d[i], p)
        better = d < min_sd
        min_sd = np.where(better, d, min_sd)
        idx = np.where(better, i, idx)
    vel = np.where((min_sd <= 0)[:, None], rb_d[idx, -1, :3], 0.0)
    lead = np.asarray(position).shape[:-1]
    return min_sd.reshape(lead), vel.reshape(lead + (3,))


def _sphere_project(rb, p):
    d = p - rb[1:4, 3]
    dist = np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2 + d[:, 2] ** 2)
    sd = dist - rb[0, 1]
    if _flipped(rb):
        sd = -sd
    with np.errstate(invalid="ignore", divide="ignore"):
        new = d / dist[:, None] * rb[0, 1] + rb[1:4, 3]
    return np.where((sd < 0)[:, None], new, p)


def _box_project(rb, p):
    q = _to_body(rb, p)
    half = rb[0, 1:4] / 2
    if _flipped(rb):
        return _to_world(rb, np.clip(q, -half, half))
    inside = np.all((q <= half) & (q >= -half), axis=1)
    dist = np.full(p.shape[0], 100.0)
    index = np.zeros(p.shape[0], dtype=np.int64)
    for i in range(3):
        a = half[i] - q[:, i]
        m = a < dist
        dist, index = np.where(m, a, dist), np.where(m, 2 * i, index)
        b = q[:, i] + half[i]
        m = b < dist
        dist, index = np.where(m, b, dist), np.where(m, 2 * i + 1, index)
    q2 = q.copy()
    step = np.where(index % 2 == 1, -dist, dist)
    for i in range(3):
        sel = index // 2 == i
        q2[sel, i] = q[sel, i] + step[sel]
    return np.where(inside[:, None], _to_world(rb, q2), p)


def _cylinder_project(rb, p):
    q = _to_body(rb, p)
    hh, rad = rb[0, 2] / 2, rb[0, 1]
    y_clip = np.clip(q[:, 1], -hh, hh)
    dist = np.sqrt(q[:, 0] ** 2 + q[:, 2] ** 2)
    sd = dist - rad
    q2 = q.copy()
    with np.errstate(invalid="ignore", divide="ignore"):
        side = np.stack([q[:, 0] / dist * rad, q[:, 1], q[:, 2] / dist * rad], axis=1)
    if _flipped(rb):
        outside = (np.abs(y_clip) == hh) | (sd > 0)
        m1 = outside & (sd < 0)
        q2[m1, 1] = y_clip[m1]
        m2 = outside & ~(sd < 0)
        q2[m2, 0], q2[m2, 2], q2[m2, 1] = side[m2, 0], side[m2, 2], y_clip[m2]
        return _to_world(rb, q2)
    inside = (sd < 0) & (np.abs(y_clip) != hh)
    a, b = q[:, 1] - hh, -(q[:, 1] + hh)
    mx = np.maximum(sd, np.maximum(a, b))
    m_side = inside & (mx == sd)
    q2[m_side, 0], q2[m_side, 2] = side[m_side, 0], side[m_side, 2]
    m_top = inside & ~(mx == sd) & (mx == a)
    q2[m_top, 1] = hh
    m_bot = inside & ~(mx == sd) & ~(mx == a)
    q2[m_bot, 1] = -hh
    return np.where(inside[:, None], _to_world(rb, q2), p)


_PROJ = {0: _sphere_project, 1: _box_project, 2: _cylinder_project}


def project(rb_d, position):
    p = np.array(position, dtype=np.float64).reshape(-1, 3)
    for i in range(rb_d.shape[0]):
        p = _PROJ[int(rb_d[i, 0, 0] // 2)](rb_d[i], p)
    return p.reshape(np.asarray(position).shape)
