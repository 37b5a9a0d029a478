"""TEST INFRASTRUCTURE ONLY — NumPy restatement of the UNet surrogate's feature builder and output gather
(``grad_v`` / ``unet_solve``, 3D_viscous_fluid_sim.ipynb:844-913), pinned bit-for-bit against fixtures produced by the
notebook's own cell (tests/golden/unet_features_*.npz)."""
import numpy as np


def pads_of(gres, data_size):
    stg = [2 * int(n) + 1 for n in gres]
    return stg, [int((int(data_size[d]) - stg[d]) / 2) for d in range(3)]


def features(gres, data_size, vx, vy, vz, sphi, lvol, pad_solid=1.0, gdx=0.0125):
    """(1, 11, X, Y, Z) fp32: dxdx dydy dzdz dxdy dxdz dydx dydz dzdx dzdy, solid flag, lvol / gdx^3.
    pad_solid: the solid flag outside the grid — 1 on the first call of a run, 0 afterwards (see run_unet_features.py)."""
    stg, (px, py, pz) = pads_of(gres, data_size)
    X, Y, Z = (int(n) for n in data_size[:3])
    vp = [np.zeros((X, Y, Z)) for _ in range(3)]
    vp[0][px:px + stg[0]:2, py + 1:py + stg[1]:2, pz + 1:pz + stg[2]:2] = vx
    vp[1][px + 1:px + stg[0]:2, py:py + stg[1]:2, pz + 1:pz + stg[2]:2] = vy
    vp[2][px + 1:px + stg[0]:2, py + 1:py + stg[1]:2, pz:pz + stg[2]:2] = vz

    def diff(a, axis):
        out = np.zeros_like(a)
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        mid = [slice(None)] * 3
        lo[axis], hi[axis], mid[axis] = slice(0, -2), slice(2, None), slice(1, -1)
        d = a[tuple(lo)] - a[tuple(hi)]
        d[(a[tuple(lo)] == 0) | (a[tuple(hi)] == 0)] = 0
        out[tuple(mid)] = d
        return out

    g = [[diff(vp[c], ax) for ax in range(3)] for c in range(3)]     # g[c][ax] = d(v_c)/d(ax)
    solid = np.full((X, Y, Z), float(pad_solid))
    solid[px:px + stg[0], py:py + stg[1], pz:pz + stg[2]] = (sphi <= 0).astype(np.float64)
    lv = np.zeros((X, Y, Z))
    lv[px:px + stg[0], py:py + stg[1], pz:pz + stg[2]] = lvol / (gdx ** 3)
    ch = [g[0][0], g[1][1], g[2][2], g[0][1], g[0][2], g[1][0], g[1][2], g[2][0], g[2][1], solid, lv]
    return np.stack(ch, axis=0).astype(np.float32)[None]


def gather(gres, data_size, out, dt):
    """the three MAC velocity increments from the network output (1, 3, X, Y, Z)"""
    stg, (px, py, pz) = pads_of(gres, data_size)
    s = int(1 / dt)
    dvx = out[0, 0, px:px + stg[0]:2, py + 1:py + stg[1]:2, pz + 1:pz + stg[2]:2] / s
    dvy = out[0, 1, px + 1:px + stg[0]:2, py:py + stg[1]:2, pz + 1:pz + stg[2]:2] / s
    dvz = out[0, 2, px + 1:px + stg[0]:2, py + 1:py + stg[1]:2, pz:pz + stg[2]:2] / s
    return dvx, dvy, dvz


def stub_net(x):
    """the stand-in network of the fixtures (oracle/refsim/run_unet_features.py::StubNet)"""
    return np.stack([x[0, 0] + np.float32(0.5) * x[0, 3] + x[0, 9], x[0, 1] - np.float32(0.1) * x[0, 10], x[0, 2] + x[0, 7] - np.float32(0.25) * x[0, 9]], axis=0)[None]
