"""Drop-in for the notebook's UNet viscosity surrogate (``grad_v`` / ``unet_solve``, 3D_viscous_fluid_sim.ipynb:844-913).

    from unet_surrogate import configure, unet_solve
    configure(GRES, DT, ckpt_file=..., data_size=(112, 176, 112))
    delvx, delvy, delvz = unet_solve(grid.x.v, grid.y.v, grid.z.v, solid_levelset.phi, fluid_volume.vol, *scratch_arrays)

``unet_solve`` keeps the notebook's 19-argument signature; the fourteen padded scratch volumes it used to fill are accepted
and left untouched (only the shape of the first one is read when no ``data_size`` was configured): the network input is
written by ONE hand-written kernel straight from the five source arrays (``fs_unet_features``), the network (``model_3d.UNet``,
weights resident on the device, loaded once) runs through cuDNN, and a second kernel gathers the three MAC velocity increments
from its output (``fs_unet_gather``).

Reference behaviour kept on purpose: the solid flag of the padding.  The notebook's in-place masking turns the -1 padding of
``sphi_sympad`` into 1 on the first call and the NEXT call's masking turns that 1 into 0, so the flag outside the grid is 1
for the first call of a run and 0 afterwards; ``UNetSurrogate`` counts its calls and does the same (``pad_solid_first`` /
``pad_solid_later`` to override).
"""
import numpy as np
import torch

from solver import _arrays as A
from solver import _native as N


def default_data_size(gres):
    """the notebook's hard-coded (112, 176, 112) generalised: the (2n+1) fine grid rounded up to a multiple of 16"""
    return tuple(int(-(-(2 * int(n) + 1) // 16) * 16) for n in gres)


class UNetSurrogate:
    def __init__(self, gres, dt, ckpt_file=None, data_size=None, gdx=0.0125, model=None, device=None, autocast_dtype=None,
                 pad_solid_first=1.0, pad_solid_later=0.0):
        self.gres = A.to_host_ints(gres)
        if len(self.gres) != 3:
            raise ValueError("UNetSurrogate needs a 3-entry gres")
        self.dt = float(dt)
        self.divisor = int(1 / self.dt)                       # ipynb:899  `/int(1/DT)`
        self.data_size = tuple(int(n) for n in (data_size[:3] if data_size is not None else default_data_size(self.gres)))
        if any(d < 2 * n + 1 for d, n in zip(self.data_size, self.gres)):
            raise ValueError("data_size is smaller than the (2*gres+1) fine grid")
        self.cell_vol = gdx ** 3                              # ipynb:885  `lvol/(0.0125**3)`
        self.device = device if device is not None else A.device()
        self.autocast_dtype = autocast_dtype
        self.pad_solid = (float(pad_solid_first), float(pad_solid_later))
        self.calls = 0
        if model is None:
            from model_3d import UNet
            model = UNet(in_channels=11)
            if ckpt_file is not None:
                model.load_state_dict(torch.load(ckpt_file, map_location="cpu")["net"])       # ONCE, not per step (ipynb:896-897)
        self.model = model.to(self.device).eval()
        if any(p.dim() == 5 for p in self.model.parameters()):
            self.model = self.model.to(memory_format=torch.channels_last_3d)
        self.input = torch.empty((1, 11) + self.data_size, dtype=torch.float32, device=self.device)

    def features(self, vx, vy, vz, sphi, lvol, pad_solid=None):
        """fill and return the (1, 11, X, Y, Z) fp32 network input"""
        g = self.gres
        sh = [tuple(n + (1 if i == a else 0) for i, n in enumerate(g)) for a in range(3)]
        v = [A.as_arg(a, n, shape=s, dtypes=(torch.float32,), want=torch.float32) for a, n, s in zip((vx, vy, vz), ("vx", "vy", "vz"), sh)]
        fine = tuple(2 * n + 1 for n in g)
        s = A.as_arg(sphi, "sphi", shape=fine, dtypes=(torch.float64,), want=torch.float64)
        lv = A.as_arg(lvol, "lvol", shape=fine, dtypes=(torch.float64,), want=torch.float64)
        if pad_solid is None:
            pad_solid = self.pad_solid[0] if self.calls == 0 else self.pad_solid[1]
        N.check(N.load().fs_unet_features(*g, *self.data_size, v[0].ptr, v[1].ptr, v[2].ptr, s.ptr, lv.ptr, float(self.cell_vol), float(pad_solid),
                                          self.input.data_ptr(), A.stream_ptr()), "fs_unet_features")
        return self.input

    def gather(self, net_out):
        """(delvx, delvy, delvz) fp32 MAC arrays from the (1, 3, X, Y, Z) network output"""
        g = self.gres
        out = net_out.to(torch.float32).contiguous()
        if tuple(out.shape) != (1, 3) + self.data_size:
            raise ValueError(f"network output has shape {tuple(out.shape)}, expected {(1, 3) + self.data_size}")
        sh = [tuple(n + (1 if i == a else 0) for i, n in enumerate(g)) for a in range(3)]
        dv = [torch.empty(s, dtype=torch.float32, device=self.device) for s in sh]
        N.check(N.load().fs_unet_gather(*g, *self.data_size, out.data_ptr(), float(self.divisor), dv[0].data_ptr(), dv[1].data_ptr(), dv[2].data_ptr(),
                                        A.stream_ptr()), "fs_unet_gather")
        return tuple(dv)

    @torch.inference_mode()
    def solve(self, vx, vy, vz, sphi, lvol):
        x = self.features(vx, vy, vz, sphi, lvol)
        self.calls += 1
        if self.autocast_dtype is not None:
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                y = self.model(x.contiguous(memory_format=torch.channels_last_3d))
        else:
            y = self.model(x.contiguous(memory_format=torch.channels_last_3d))
        return self.gather(y)


_default = None


def configure(gres, dt, ckpt_file=None, **kw):
    """set up the surrogate used by the module-level ``unet_solve`` (the notebook keeps these as globals: GRES, DT, ckpt_file)"""
    global _default
    _default = UNetSurrogate(gres, dt, ckpt_file=ckpt_file, **kw)
    return _default


def unet_solve(vx, vy, vz, sphi, lvol, vx_sympad=None, vy_sympad=None, vz_sympad=None, lvol_sympad=None, sphi_sympad=None,
               dxdx=None, dxdy=None, dxdz=None, dydx=None, dydy=None, dydz=None, dzdx=None, dzdy=None, dzdz=None):
    """the notebook's signature (ipynb:878); returns (delvx, delvy, delvz).  Call ``configure`` first."""
    if _default is None:
        raise RuntimeError("unet_surrogate.configure(gres, dt, ckpt_file) must be called before unet_solve")
    return _default.solve(vx, vy, vz, sphi, lvol)


def grad_v(gx_v, gy_v, gz_v, dxdx, dxdy, dxdz, dydx, dydy, dydz, dzdx, dzdy, dzdz):
    """the notebook's stand-alone gradient helper (ipynb:844-876) on already padded torch volumes, in place.
    (``unet_solve`` does not call it: the feature kernel forms the same nine differences without the padded copies.)"""
    outs = ((dxdx, dxdy, dxdz), (dydx, dydy, dydz), (dzdx, dzdy, dzdz))
    for v, row in zip((gx_v, gy_v, gz_v), outs):
        for ax, o in enumerate(row):
            lo = [slice(None)] * v.dim()
            hi = [slice(None)] * v.dim()
            mid = [slice(None)] * v.dim()
            lo[ax], hi[ax], mid[ax] = slice(0, -2), slice(2, None), slice(1, -1)
            a, b = v[tuple(lo)], v[tuple(hi)]
            o[tuple(mid)] = torch.where((a == 0) | (b == 0), torch.zeros_like(a), a - b)
