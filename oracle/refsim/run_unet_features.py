#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — golden fixtures for the UNet surrogate's feature builder and output gather.

Executes code cell 12 of ``/root/reference/3D_viscous_fluid_sim.ipynb`` (``grad_v``, ``unet_solve``; ipynb:844-913) UNMODIFIED
with the NumPy-backed ``cupy`` shim.  The cell instantiates ``model_3d.UNet`` and loads a checkpoint that is not part of the
repository, so the names it resolves at call time are pointed at stand-ins from outside the cell: ``UNet`` -> a stub network
whose output is a fixed linear combination of input channels (so the output gather is exercised with non-trivial values),
``torch.load`` -> an empty state dict, ``torch.as_tensor`` -> ignores the ``device=`` argument (no GPU here).  Everything
the fixture pins — the 11-channel fp32 input tensor the cell builds and the three velocity increments it slices out of the
network output — is computed by the reference's own lines.

Two consecutive calls are recorded: the cell masks ``sphi_sympad`` in place (−1 padding -> 1), and the NEXT call's masking
maps that 1 to 0, so the solid flag of the padding differs between the first and the later calls of a run.

Never run on the GPU box.  Usage: python oracle/refsim/run_unet_features.py
"""
import json
import os
import sys
import types

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("FLUID_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, HERE)
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import cupy as cp  # noqa: E402
import torch  # noqa: E402
from numba import cuda  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
captured = {}


class StubNet:
    def __init__(self, in_channels=5):
        self.in_channels = in_channels

    def load_state_dict(self, sd):
        pass

    def to(self, dev):
        return self

    def __call__(self, x):
        captured["input"] = x.detach().clone()
        return torch.stack([x[0, 0] + 0.5 * x[0, 3] + x[0, 9], x[0, 1] - 0.1 * x[0, 10], x[0, 2] + x[0, 7] - 0.25 * x[0, 9]], dim=0).unsqueeze(0)


def main():
    with open(os.path.join(REF, "3D_viscous_fluid_sim.ipynb")) as f:
        nb = json.load(f)
    src = "".join(nb["cells"][12]["source"])
    g = np.array([6, 7, 8], dtype=np.int64)
    if not hasattr(cp, "concatenate"):
        cp.concatenate = lambda arrs, axis=0: cp.asarray(np.concatenate([np.asarray(a) for a in arrs], axis=axis))
    if not hasattr(cp, "expand_dims"):
        cp.expand_dims = lambda a, axis: cp.asarray(np.expand_dims(np.asarray(a), axis))
    cp.from_dlpack = lambda t: cp.asarray(t.detach().cpu().numpy()) if isinstance(t, torch.Tensor) else cp.asarray(np.from_dlpack(t))
    ns = {"cp": cp, "cuda": cuda, "GRES": cp.asarray(g), "DT": 1 / 300, "ckpt_file": "unused.ckpt", "device_num": 0, "__name__": "cell12"}
    import torch.utils.dlpack as tud
    real_load, real_as_tensor, real_to_dlpack = torch.load, torch.as_tensor, tud.to_dlpack
    torch.load = lambda *a, **k: {"net": {}}
    torch.as_tensor = lambda data, *a, **k: real_as_tensor(np.asarray(data))
    tud.to_dlpack = lambda t: t                       # the stubbed cp.from_dlpack takes the tensor itself
    try:
        exec(compile(src, "<ipynb cell 12: unet>", "exec"), ns)
        ns["UNet"] = StubNet                          # resolved when unet_solve runs
        rng = np.random.default_rng(31)
        out = dict(gres=g, dt=1 / 300, data_size=np.array(ns["data_size"]))
        pads = [ns["x_pad_l"], ns["y_pad_l"], ns["z_pad_l"]]
        out["pads"] = np.array(pads)
        bufs = [ns[k] for k in ("vx_sympad", "vy_sympad", "vz_sympad", "lvol_sympad", "sphi_sympad", "dxdx", "dxdy", "dxdz", "dydx", "dydy", "dydz", "dzdx", "dzdy", "dzdz")]
        for call in (1, 2):
            sh = [tuple(int(n) + (1 if i == a else 0) for i, n in enumerate(g)) for a in range(3)]
            v = [rng.normal(0, 1, s).astype(np.float32) for s in sh]
            for a in v:                               # exact zeros: faces without mass (the cell tests == 0)
                a[rng.random(a.shape) < 0.3] = 0.0
            fine = tuple(2 * int(n) + 1 for n in g)
            sphi = rng.normal(0.02, 0.05, fine)
            sphi[rng.random(fine) < 0.05] = 0.0
            lvol = np.clip(rng.normal(0.3, 0.5, fine), 0, 1) * (0.0125 / 2) ** 3
            d = ns["unet_solve"](cp.asarray(v[0]), cp.asarray(v[1]), cp.asarray(v[2]), cp.asarray(sphi), cp.asarray(lvol), *bufs)
            out.update({f"vx{call}": v[0], f"vy{call}": v[1], f"vz{call}": v[2], f"sphi{call}": sphi, f"lvol{call}": lvol,
                        f"input{call}": captured["input"].numpy(), f"delvx{call}": np.asarray(d[0]), f"delvy{call}": np.asarray(d[1]),
                        f"delvz{call}": np.asarray(d[2])})
    finally:
        torch.load, torch.as_tensor, tud.to_dlpack = real_load, real_as_tensor, real_to_dlpack
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, "unet_features_6x7x8.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB); input {out['input1'].shape} {out['input1'].dtype}, delvx {out['delvx1'].shape} {out['delvx1'].dtype}")


if __name__ == "__main__":
    main()
