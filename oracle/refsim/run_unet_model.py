#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — pins for the UNet mirror (python-fluid-simulation_b200/model_3d.py): the reference network's
state_dict layout (keys + shapes) and one forward pass of the UNMODIFIED ``/root/reference/model_3d.py`` with seed-0 default
initialisation on a small seeded input.  Writes tests/golden/unet_model_ref.npz.  Never run on the GPU box."""
import importlib.util
import json
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("FLUID_REFERENCE_ROOT", "/root/reference")


def main():
    spec = importlib.util.spec_from_file_location("ref_model_3d", os.path.join(REF, "model_3d.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(0)
    net = ref.UNet(in_channels=11).eval()
    layout = {k: list(v.shape) for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 11, 16, 16, 16, generator=g)
    with torch.no_grad():
        y = net(x)
    sd = net.state_dict()
    probes = {k: float(sd[k].double().sum()) for k in ("enc1_1.0.weight", "enc5_1.0.bias", "unpool3.weight", "fc.weight")}
    path = os.path.join(REPO, "tests", "golden", "unet_model_ref.npz")
    np.savez_compressed(path, layout=json.dumps(layout), x=x.numpy(), y=y.numpy(), probes=json.dumps(probes), torch_version=torch.__version__)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB), {len(layout)} tensors, {sum(int(np.prod(s)) for s in layout.values()) / 1e6:.1f} M parameters")


if __name__ == "__main__":
    main()
