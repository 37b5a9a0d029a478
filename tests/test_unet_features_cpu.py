"""CPU: NumPy restatement of the UNet surrogate's feature builder / output gather against the notebook cell's own output."""
import numpy as np

from conftest import load_golden
from oracle import numpy_oracle_unet as U


def test_features_and_gather_bit_exact():
    f = load_golden("unet_features_6x7x8")
    for call, pad_solid in ((1, 1.0), (2, 0.0)):        # the padding's solid flag flips after the first call (reference quirk)
        x = U.features(f["gres"], f["data_size"], f[f"vx{call}"], f[f"vy{call}"], f[f"vz{call}"], f[f"sphi{call}"], f[f"lvol{call}"], pad_solid)
        assert x.dtype == np.float32 and x.shape == f[f"input{call}"].shape
        assert np.array_equal(x, f[f"input{call}"])
        d = U.gather(f["gres"], f["data_size"], U.stub_net(x), float(f["dt"]))
        for a, k in zip(d, ("delvx", "delvy", "delvz")):
            assert np.array_equal(a, f[f"{k}{call}"])


def test_unet_mirror_matches_reference_model():
    """state_dict layout and one forward pass of the mirror against the UNMODIFIED reference network (seed-0 default init)"""
    import json
    import os
    import sys

    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "python-fluid-simulation_b200"))
    from model_3d import UNet
    f = load_golden("unet_model_ref")
    layout = json.loads(str(f["layout"]))
    torch.manual_seed(0)
    net = UNet(in_channels=11).eval()
    sd = net.state_dict()
    assert list(sd.keys()) == list(layout.keys())                       # a reference checkpoint loads unchanged
    assert all(list(sd[k].shape) == layout[k] for k in layout)
    if str(f["torch_version"]) == torch.__version__:                      # same init stream -> same weights -> same output
        for k, v in json.loads(str(f["probes"])).items():
            assert abs(float(sd[k].double().sum()) - v) <= 1e-9 * max(1.0, abs(v))
        with torch.no_grad():
            y = net(torch.from_numpy(f["x"]))
        assert float((y - torch.from_numpy(f["y"])).abs().max()) <= 1e-5 * float(np.abs(f["y"]).max())
