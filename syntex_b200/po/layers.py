"""Generator building blocks of the PO models, NHWC at the boundary.

The reference builds these from tensorpack 0.9.8 layers (Conv2D, InstanceNorm) and tf.pad / resize ops
(po/progressive_model.py:24-143). tensorpack is absent from /root/reference, so its defaults are restated from
its published source and are UNVERIFIED here: Conv2D kernels He-normal on fan-in (variance_scaling 2.0), zero biases;
InstanceNorm per (image, channel) over H x W with eps 1e-5 and an affine gamma = 1 / beta = 0.
These layers run on stock ATen/cuDNN kernels: the generator is the "next" row, outside the hand-written hot path.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def symmetric_pad(x, before, after=None):
    """tf.pad(mode="SYMMETRIC") on the two spatial axes of an NCHW tensor: reflects INCLUDING the edge pixel."""
    after = before if after is None else after
    if before == 0 and after == 0:
        return x
    H, W = x.shape[-2], x.shape[-1]
    if before > H or after > H or before > W or after > W:
        raise ValueError("symmetric_pad: padding larger than the tensor")
    parts = []
    if before:
        parts.append(x[..., :before, :].flip(-2))
    parts.append(x)
    if after:
        parts.append(x[..., H - after:, :].flip(-2))
    x = torch.cat(parts, dim=-2)
    parts = []
    if before:
        parts.append(x[..., :before].flip(-1))
    parts.append(x)
    if after:
        parts.append(x[..., W - after:].flip(-1))
    return torch.cat(parts, dim=-1)


class SymConv(nn.Module):
    """tf.pad(SYMMETRIC) + Conv2D(padding="VALID") (+ InstanceNorm + ReLU when `act`), progressive_model.py:171-175."""

    def __init__(self, cin, cout, ksize=3, act=True, use_bias=False):
        super().__init__()
        self.ksize = ksize
        self.conv = nn.Conv2d(cin, cout, ksize, padding=0, bias=use_bias)
        nn.init.kaiming_normal_(self.conv.weight, mode="fan_in", nonlinearity="relu")
        if use_bias:
            nn.init.zeros_(self.conv.bias)
        self.norm = nn.InstanceNorm2d(cout, eps=1e-5, affine=True) if act else None

    def forward(self, x):
        p0 = self.ksize // 2
        x = self.conv(symmetric_pad(x, p0, self.ksize - 1 - p0))
        if self.norm is not None:
            x = F.relu(self.norm(x))
        return x


class ResBlock(nn.Module):
    """build_res_block (progressive_model.py:77-99): relu -> conv0 (IN+ReLU) -> conv1 -> IN, plus the shortcut."""

    def __init__(self, chan):
        super().__init__()
        self.conv0 = SymConv(chan, chan, 3, act=True)
        self.conv1 = SymConv(chan, chan, 3, act=False)
        self.inorm = nn.InstanceNorm2d(chan, eps=1e-5, affine=True)

    def forward(self, x):
        return self.inorm(self.conv1(self.conv0(F.relu(x)))) + x


class PreResBlock(nn.Module):
    """build_pre_res_block (progressive_model.py:101-121): IN+ReLU -> conv0 (IN+ReLU) -> conv1, plus the shortcut."""

    def __init__(self, chan):
        super().__init__()
        self.inorm = nn.InstanceNorm2d(chan, eps=1e-5, affine=True)
        self.conv0 = SymConv(chan, chan, 3, act=True)
        self.conv1 = SymConv(chan, chan, 3, act=False)

    def forward(self, x):
        return self.conv1(self.conv0(F.relu(self.inorm(x)))) + x


class UpsampleNN(nn.Module):
    """build_upsampling_nn (progressive_model.py:123-134): nearest x2, symmetric pad, (2*scale+1)^2 conv, no bias."""

    def __init__(self, cin, cout, scale=2):
        super().__init__()
        self.scale = scale
        self.conv = SymConv(cin, cout, 2 * scale + 1, act=False)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=self.scale, mode="nearest"))


class UpsampleDeconv(nn.Module):
    """build_upsampling_deconv (progressive_model.py:136-143): Conv2DTranspose(ksize = 2*scale, stride = scale, SAME)."""

    def __init__(self, cin, cout, scale=2):
        super().__init__()
        self.deconv = nn.ConvTranspose2d(cin, cout, 2 * scale, stride=scale, padding=scale // 2, bias=False)
        nn.init.kaiming_normal_(self.deconv.weight, mode="fan_in", nonlinearity="relu")

    def forward(self, x):
        return self.deconv(x)
