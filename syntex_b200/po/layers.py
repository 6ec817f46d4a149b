"""Generator building blocks of the PO models, NHWC at the boundary.

The reference builds these from tensorpack 0.9.8 layers (Conv2D, InstanceNorm) and tf.pad / resize ops
(po/progressive_model.py:24-143). tensorpack is absent from /root/reference, so its defaults are restated from
its published source and are UNVERIFIED here: Conv2D kernels He-normal on fan-in (variance_scaling 2.0), zero biases;
InstanceNorm per (image, channel) over H x W with eps 1e-5 and an affine gamma = 1 / beta = 0.
Two routes share these modules and their parameters: `forward` on stock ATen/cuDNN kernels (NCHW fp32; also what
the CPU tests run over the oracle door), and `forward_native` on the hand-written kernels of po/gen_ops.py (NHWC bf16
activations: tcgen05 convolutions incl. the filter gradient, fused InstanceNorm / ReLU / padding passes).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


_PAD_INDEX = {}


def _pad_index(n, before, after, device):
    key = (n, before, after, str(device))
    idx = _PAD_INDEX.get(key)
    if idx is None:
        idx = torch.tensor(list(range(before - 1, -1, -1)) + list(range(n)) + list(range(n - 1, n - 1 - after, -1)),
                           dtype=torch.long, device=device)
        _PAD_INDEX[key] = idx
    return idx


def symmetric_pad(x, before, after=None):
    """tf.pad(mode="SYMMETRIC") on the two spatial axes of an NCHW tensor: reflects INCLUDING the edge pixel.
    Two gathers on the NHWC view (one per axis; their backward is one index_add each), so a channels_last input
    stays channels_last - slicing + flipping + concatenating costs a dozen kernels and a dozen full-size
    zero-fills in backward."""
    after = before if after is None else after
    if before == 0 and after == 0:
        return x
    H, W = x.shape[-2], x.shape[-1]
    if before > H or after > H or before > W or after > W:
        raise ValueError("symmetric_pad: padding larger than the tensor")
    if x.dim() != 4:
        raise ValueError("symmetric_pad expects [B,C,H,W]")
    if before <= 1 and after <= 1:
        # one pixel of symmetric padding is edge replication: a single native kernel, forward and backward
        return F.pad(x, (before, after, before, after), mode="replicate")
    y = x.permute(0, 2, 3, 1)                       # a contiguous view when x is channels_last
    y = y.index_select(1, _pad_index(H, before, after, x.device))
    y = y.index_select(2, _pad_index(W, before, after, x.device))
    return y.permute(0, 3, 1, 2)


class InstanceNormAffine(nn.Module):
    """tensorpack InstanceNorm: per image and channel over H x W, eps 1e-5, affine. Two stock routes to the same
    arithmetic, chosen by measurement: group_norm with one channel per group when a backward pass will follow (its
    backward is ~2.5x faster than instance_norm's batch-norm route), instance_norm for inference (faster forward)."""

    def __init__(self, chan, eps=1e-5):
        super().__init__()
        self.chan, self.eps = chan, eps
        self.weight = nn.Parameter(torch.ones(chan))
        self.bias = nn.Parameter(torch.zeros(chan))

    def forward(self, x):
        if torch.is_grad_enabled():
            return F.group_norm(x, self.chan, self.weight, self.bias, self.eps)
        return F.instance_norm(x, None, None, self.weight, self.bias, True, 0.0, self.eps)

    def forward_native(self, x, relu=False, shortcut=None):
        """relu: False = none, True = ReLU, a float = LeakyReLU slope (fused into the normalisation pass)."""
        from . import gen_ops
        return gen_ops.instance_norm(x, self.weight, self.bias, self.eps, relu, shortcut)


def instance_norm(chan):
    return InstanceNormAffine(chan)


class SymConv(nn.Module):
    """tf.pad(SYMMETRIC) + Conv2D(padding="VALID") (+ InstanceNorm + ReLU when `act`), progressive_model.py:171-175."""

    def __init__(self, cin, cout, ksize=3, act=True, use_bias=False, zero_pad=False):
        super().__init__()
        self.ksize = ksize
        self.zero_pad = zero_pad          # Conv2D(padding="SAME") instead of tf.pad(SYMMETRIC) + VALID
        self.conv = nn.Conv2d(cin, cout, ksize, padding=ksize // 2 if zero_pad else 0, bias=use_bias)
        nn.init.kaiming_normal_(self.conv.weight, mode="fan_in", nonlinearity="relu")
        if use_bias:
            nn.init.zeros_(self.conv.bias)
        self.norm = instance_norm(cout) if act else None

    def forward(self, x):
        p0 = self.ksize // 2
        x = self.conv(x if self.zero_pad else symmetric_pad(x, p0, self.ksize - 1 - p0))
        if self.norm is not None:
            x = F.relu(self.norm(x))
        return x

    def forward_native(self, x, relu_in=False):
        """x NHWC bf16; relu_in: the ReLU the caller applies to x first, fused into the padding pass."""
        from . import gen_ops
        if self.ksize != 3:
            raise NotImplementedError("native generator route: 3x3 layers only (5x5 runs as 4 phase filters)")
        if self.zero_pad and relu_in:
            x, relu_in = torch.relu(x), False
        x = gen_ops.conv3x3(x, self.conv.weight, self.conv.bias, "zero" if self.zero_pad else "rep", relu_in)
        if self.norm is not None:
            x = self.norm.forward_native(x, relu=True)
        return x


class ResBlock(nn.Module):
    """build_res_block (progressive_model.py:77-99): relu -> conv0 (IN+ReLU) -> conv1 -> IN, plus the shortcut."""

    def __init__(self, chan):
        super().__init__()
        self.conv0 = SymConv(chan, chan, 3, act=True)
        self.conv1 = SymConv(chan, chan, 3, act=False)
        self.inorm = instance_norm(chan)

    def forward(self, x):
        return self.inorm(self.conv1(self.conv0(F.relu(x)))) + x

    def forward_native(self, x):
        return self.inorm.forward_native(self.conv1.forward_native(self.conv0.forward_native(x, relu_in=True)),
                                         relu=False, shortcut=x)


class PreResBlock(nn.Module):
    """build_pre_res_block (progressive_model.py:101-121): IN+ReLU -> conv0 (IN+ReLU) -> conv1, plus the shortcut."""

    def __init__(self, chan):
        super().__init__()
        self.inorm = instance_norm(chan)
        self.conv0 = SymConv(chan, chan, 3, act=True)
        self.conv1 = SymConv(chan, chan, 3, act=False)

    def forward(self, x):
        return self.conv1(self.conv0(F.relu(self.inorm(x)))) + x

    def forward_native(self, x):
        return self.conv1.forward_native(self.conv0.forward_native(self.inorm.forward_native(x, relu=True))) + x


class UpsampleNN(nn.Module):
    """build_upsampling_nn (progressive_model.py:123-134): nearest x2, symmetric pad, (2*scale+1)^2 conv, no bias."""

    def __init__(self, cin, cout, scale=2):
        super().__init__()
        self.scale = scale
        self.conv = SymConv(cin, cout, 2 * scale + 1, act=False)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=self.scale, mode="nearest"))

    def forward_native(self, x, relu_in=False):
        """Nearest x2 + symmetric pad 2 + 5x5 conv == edge-replicated 3x3 conv of the low-resolution map with four
        phase filters (gen_ops.upsample_weights), then a pixel shuffle: 36 tap-GEMMs on h x w pixels instead of 25 on
        4 h x w."""
        from . import gen_ops
        if self.scale != 2:
            raise NotImplementedError("native generator route: x2 up-sampling only")
        w5 = self.conv.conv.weight
        y = gen_ops.conv3x3(x, gen_ops.upsample_weights(w5), None, "rep", relu_in)
        return gen_ops.pixel_shuffle_nhwc(y, w5.shape[0])


class UpsampleDeconv(nn.Module):
    """build_upsampling_deconv (progressive_model.py:136-143): Conv2DTranspose(ksize = 2*scale, stride = scale, SAME)."""

    def __init__(self, cin, cout, scale=2):
        super().__init__()
        self.deconv = nn.ConvTranspose2d(cin, cout, 2 * scale, stride=scale, padding=scale // 2, bias=False)
        nn.init.kaiming_normal_(self.deconv.weight, mode="fan_in", nonlinearity="relu")

    def forward(self, x):
        return self.deconv(x)
