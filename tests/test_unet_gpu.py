"""GPU: the UNet surrogate path (unet_surrogate.py -> fs_unet_* kernels, model_3d.UNet through cuDNN).
The feature builder and the output gather are bit-exact against the fixtures the notebook's own cell produced; the network
mirror has the reference's state_dict layout and (small check) agrees with a plain fp32 PyTorch evaluation of itself."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import numpy_oracle_unet as U

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


class Stub(torch.nn.Module):
    """the stand-in network of the fixtures (oracle/refsim/run_unet_features.py::StubNet)"""

    def forward(self, x):
        return torch.stack([x[0, 0] + 0.5 * x[0, 3] + x[0, 9], x[0, 1] - 0.1 * x[0, 10], x[0, 2] + x[0, 7] - 0.25 * x[0, 9]], dim=0).unsqueeze(0)


def test_features_and_gather_vs_reference():
    from unet_surrogate import UNetSurrogate
    f = load_golden("unet_features_6x7x8")
    s = UNetSurrogate(f["gres"], float(f["dt"]), data_size=tuple(int(n) for n in f["data_size"][:3]), model=Stub())
    assert s.divisor == 300
    for call in (1, 2):                                  # the padding's solid flag is 1 on the first call, 0 afterwards
        args = [dev(f[f"{k}{call}"]) for k in ("vx", "vy", "vz", "sphi", "lvol")]
        d = s.solve(*args)
        assert np.array_equal(s.input.cpu().numpy(), f[f"input{call}"])
        for a, k in zip(d, ("delvx", "delvy", "delvz")):
            assert a.dtype == torch.float32
            assert np.array_equal(a.cpu().numpy(), f[f"{k}{call}"])


def test_features_on_the_notebook_grid_vs_oracle():
    """the notebook's own sizes: gres (48, 80, 48) in a (112, 176, 112) volume"""
    from unet_surrogate import UNetSurrogate, default_data_size
    g = (48, 80, 48)
    assert default_data_size(g) == (112, 176, 112)
    rng = np.random.default_rng(9)
    sh = [tuple(n + (1 if i == a else 0) for i, n in enumerate(g)) for a in range(3)]
    v = [rng.normal(0, 1, s).astype(np.float32) for s in sh]
    for a in v:
        a[rng.random(a.shape) < 0.4] = 0.0
    fine = tuple(2 * n + 1 for n in g)
    sphi = rng.normal(0.02, 0.05, fine)
    lvol = np.clip(rng.normal(0.3, 0.5, fine), 0, 1) * (0.0125 / 2) ** 3
    s = UNetSurrogate(g, 1 / 300, model=Stub())
    x = s.features(*[dev(a) for a in v], dev(sphi), dev(lvol))
    ref = U.features(g, s.data_size, *v, sphi, lvol, 1.0)
    assert np.array_equal(x.cpu().numpy(), ref)
    d = s.gather(Stub()(x))
    for a, b in zip(d, U.gather(g, s.data_size, U.stub_net(ref), 1 / 300)):
        assert np.array_equal(a.cpu().numpy(), b)


def test_unet_mirror_layout_and_inference():
    from model_3d import UNet
    torch.manual_seed(0)
    net = UNet(in_channels=11)
    keys = list(net.state_dict().keys())
    assert keys[0] == "enc1_1.0.weight" and keys[-1] == "fc.bias" and len(keys) == 46
    assert net.state_dict()["enc5_1.0.weight"].shape == (1024, 512, 3, 3, 3)
    assert net.state_dict()["unpool4.weight"].shape == (512, 512, 2, 2, 2)
    assert net.state_dict()["dec4_2.0.weight"].shape == (512, 1024, 3, 3, 3)
    x = torch.randn(1, 11, 16, 32, 16)
    with torch.no_grad():
        ref = net(x)                                     # plain fp32 PyTorch on the CPU
    from unet_surrogate import UNetSurrogate
    s = UNetSurrogate((7, 15, 7), 1 / 300, data_size=(16, 32, 16), model=net)
    with torch.inference_mode():
        y = s.model(x.cuda().contiguous(memory_format=torch.channels_last_3d)).float().cpu()
    assert y.shape == (1, 3, 16, 32, 16)
    assert float((y - ref).abs().max()) < 1e-3 * float(ref.abs().max())
