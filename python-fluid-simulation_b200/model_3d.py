"""Drop-in for the reference's ``model_3d.py``: the 3-D UNet surrogate of the implicit viscosity step.

Same class name, constructor and ``state_dict`` layout as the reference (model_3d.py:9-136: five resolution levels, two
3x3x3 convolutions + tanh per level, average pooling down, 2x2x2 transposed convolutions up, skip concatenation, a 1x1x1
head with 3 outputs), so a checkpoint trained with the reference loads unchanged (``load_state_dict(ckpt['net'])``).

What differs is how it is run.  The notebook builds a new network and re-reads the checkpoint on EVERY time step
(ipynb:896-898); here the weights stay resident on the device in ``channels_last_3d`` layout, inference runs under
``torch.inference_mode`` and, optionally, in bf16 autocast so that cuDNN's implicit-GEMM kernels use the Blackwell tensor
cores.  The convolutions themselves are library calls (cuDNN through PyTorch) — this file does not contain hand-written
tensor-core kernels; the hand-written parts of the surrogate path are the input feature builder and the output gather
(``csrc/fs_unet.cu``), see ``unet_surrogate.py``.
"""
import torch
import torch.nn as nn


class UNet(nn.Module):
    LEVELS = (64, 128, 256, 512, 1024)

    def __init__(self, in_channels=5):
        super().__init__()
        self.in_channels = in_channels

        def cbr(cin, cout):
            return nn.Sequential(nn.Conv3d(cin, cout, kernel_size=3, stride=1, padding=1, bias=True), nn.Tanh())

        c1, c2, c3, c4, c5 = self.LEVELS
        # registration order = the reference's (it fixes both the state_dict keys and the consumption of the init RNG)
        self.enc1_1 = cbr(in_channels, c1)
        self.enc1_2 = cbr(c1, c1)
        self.pool1 = nn.AvgPool3d(kernel_size=2)
        self.enc2_1 = cbr(c1, c2)
        self.enc2_2 = cbr(c2, c2)
        self.pool2 = nn.AvgPool3d(kernel_size=2)
        self.enc3_1 = cbr(c2, c3)
        self.enc3_2 = cbr(c3, c3)
        self.pool3 = nn.AvgPool3d(kernel_size=2)
        self.enc4_1 = cbr(c3, c4)
        self.enc4_2 = cbr(c4, c4)
        self.pool4 = nn.AvgPool3d(kernel_size=2)
        self.enc5_1 = cbr(c4, c5)
        self.dec5_1 = cbr(c5, c4)
        self.unpool4 = nn.ConvTranspose3d(c4, c4, kernel_size=2, stride=2, padding=0, bias=True)
        self.dec4_2 = cbr(2 * c4, c4)
        self.dec4_1 = cbr(c4, c3)
        self.unpool3 = nn.ConvTranspose3d(c3, c3, kernel_size=2, stride=2, padding=0, bias=True)
        self.dec3_2 = cbr(2 * c3, c3)
        self.dec3_1 = cbr(c3, c2)
        self.unpool2 = nn.ConvTranspose3d(c2, c2, kernel_size=2, stride=2, padding=0, bias=True)
        self.dec2_2 = cbr(2 * c2, c2)
        self.dec2_1 = cbr(c2, c1)
        self.unpool1 = nn.ConvTranspose3d(c1, c1, kernel_size=2, stride=2, padding=0, bias=True)
        self.dec1_2 = cbr(2 * c1, c1)
        self.dec1_1 = cbr(c1, c1)
        self.fc = nn.Conv3d(c1, 3, kernel_size=1, stride=1, padding=0, bias=True)

    def forward(self, x):
        e1 = self.enc1_2(self.enc1_1(x))
        e2 = self.enc2_2(self.enc2_1(self.pool1(e1)))
        e3 = self.enc3_2(self.enc3_1(self.pool2(e2)))
        e4 = self.enc4_2(self.enc4_1(self.pool3(e3)))
        b = self.dec5_1(self.enc5_1(self.pool4(e4)))
        d4 = self.dec4_1(self.dec4_2(torch.cat((self.unpool4(b), e4), dim=1)))
        d3 = self.dec3_1(self.dec3_2(torch.cat((self.unpool3(d4), e3), dim=1)))
        d2 = self.dec2_1(self.dec2_2(torch.cat((self.unpool2(d3), e2), dim=1)))
        d1 = self.dec1_1(self.dec1_2(torch.cat((self.unpool1(d2), e1), dim=1)))
        return self.fc(d1)
