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
