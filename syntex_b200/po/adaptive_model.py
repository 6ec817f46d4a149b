"""AdaptiveSynTex (po/adaptive_model.py:58-366) over the CUDA hot path.

`n_stage` unfoldings; every stage looks at the current image through VGG19, back-propagates the whole style loss to
the five Gram layers (tf.gradients(loss, [feat_k]): a full VGG backward) and lets the generator turn those gradients
into a pre-image update. Training differentiates through the gradients unless --stop-grad: the double backward
through VGG19 (`po.texture_op.vgg_stage_preparation`: tangent sweep + closed-form injection). The generator is stock
torch; tensorpack layer defaults restated from its published source, unverified.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..aparse import ArgParser
from .layers import instance_norm, symmetric_pad
from .model import SynTexModelDesc
from .progressive_model import LAYER_CHANNELS
from .texture_op import CudaTexture

MAX_EPOCH = 200
STEPS_PER_EPOCH = 4000
IMAGESIZE = 224
LR = 5e-5
BETA1 = 0.5
BETA2 = 0.999
BATCH = 1
TEST_BATCH = 1


def _pad(x, before, after, pad_type):
    """tf.pad(mode=pad_type) on the spatial axes of NCHW: REFLECT / SYMMETRIC / CONSTANT (adaptive_model.py:50-56)."""
    if before == 0 and after == 0:
        return x
    if pad_type == "SYMMETRIC":
        return symmetric_pad(x, before, after)
    if pad_type == "REFLECT":
        return F.pad(x, (before, after, before, after), mode="reflect")
    if pad_type == "CONSTANT":
        return F.pad(x, (before, after, before, after))
    raise ValueError("pad type %r" % (pad_type,))


def _norm(norm_type, chan):
    if norm_type == "instance":
        return instance_norm(chan)
    if norm_type == "batch":
        return nn.BatchNorm2d(chan, eps=1e-5, momentum=0.1)
    return nn.Identity()


def _act(act_type, x):
    return F.relu(x) if act_type == "relu" else F.leaky_relu(x, 0.2)


# ---- the same modules on the hand-written kernels (po/gen_ops.py): NHWC bf16 activations ------------------------
_NATIVE_PAD = {"SYMMETRIC": "symmetric", "REFLECT": "reflect", "CONSTANT": "zero"}


def _act_arg(act_type):
    """Activation argument of the fused kernels: True = ReLU, a float = LeakyReLU slope."""
    return True if act_type == "relu" else 0.2


def _norm_is_native(norm, batch):
    """InstanceNorm always; BatchNorm with batch statistics (po/improved_model.py:353-360) at any batch: in NHWC the
    batch folds into the pixel axis, so it is InstanceNorm of ONE image of B*H*W pixels (see _norm_native); nothing to
    do for Identity."""
    from .layers import InstanceNormAffine
    if norm is None or isinstance(norm, (nn.Identity, InstanceNormAffine)):
        return True
    return type(norm).__name__ == "_BatchStatsNorm"


def _norm_native(norm, x, act=None, shortcut=None):
    from . import gen_ops
    if norm is None or isinstance(norm, nn.Identity):
        if act is not None:
            x = torch.relu(x) if act is True else F.leaky_relu(x, float(act))
        return x if shortcut is None else x + shortcut
    if type(norm).__name__ == "_BatchStatsNorm" and x.shape[0] > 1:
        # statistics over (N, H, W) per channel, biased variance: exactly the per-image statistics of the batch viewed
        # as one [1, N*H, W, C] image (NHWC: a reshape, no copy) - same kernels, forward and backward
        B, H, W, C = x.shape
        sc = None if shortcut is None else shortcut.reshape(1, B * H, W, C)
        y = gen_ops.instance_norm(x.reshape(1, B * H, W, C), norm.weight, norm.bias, norm.eps,
                                  act if act is not None else False, sc)
        return y.reshape(B, H, W, C)
    return gen_ops.instance_norm(x, norm.weight, norm.bias, norm.eps, act if act is not None else False, shortcut)


class PadConv2D(nn.Module):
    """adaptive_model.py:50-56: tf.pad + Conv2D(VALID); norm_act=True appends norm + activation (the argscope
    default), False leaves it linear."""

    def __init__(self, cin, cout, ksize, pad_type, norm_act, use_bias, norm_type="instance", act_type="relu"):
        super().__init__()
        self.ksize, self.pad_type, self.act_type = ksize, pad_type, act_type
        self.conv = nn.Conv2d(cin, cout, ksize, bias=use_bias)
        nn.init.kaiming_normal_(self.conv.weight, mode="fan_in", nonlinearity="relu")
        if use_bias:
            nn.init.zeros_(self.conv.bias)
        self.norm = _norm(norm_type, cout) if norm_act else None

    def forward(self, x):
        psize = self.ksize - 1
        x = self.conv(_pad(x, psize // 2, psize - psize // 2, self.pad_type))
        if self.norm is not None:
            x = _act(self.act_type, self.norm(x))
        return x

    def forward_native(self, x, act_in=None):
        """x NHWC bf16; act_in: the activation the caller applies to x first (True = ReLU, float = LeakyReLU slope),
        fused into the padding pass."""
        from . import gen_ops
        if self.ksize != 3:
            raise NotImplementedError("native generator route: 3x3 layers only")
        mode = _NATIVE_PAD[self.pad_type]
        if mode == "zero" and act_in is not None:
            x, act_in = _norm_native(None, x, act=act_in), None
        x = gen_ops.conv3x3(x, self.conv.weight, self.conv.bias, mode, act_in if act_in is not None else False)
        if self.norm is not None:
            x = _norm_native(self.norm, x, act=_act_arg(self.act_type))
        return x


class _ResBlock(nn.Module):
    def __init__(self, chan, pad_type, norm_type, act_type, pre_act):
        super().__init__()
        self.pre_act, self.act_type = pre_act, act_type
        self.conv0 = PadConv2D(chan, chan, 3, pad_type, True, False, norm_type, act_type)
        self.conv1 = PadConv2D(chan, chan, 3, pad_type, False, False, norm_type, act_type)
        self.norm = _norm(norm_type, chan)

    def forward(self, x):
        if self.pre_act:       # build_pre_res_block (adaptive_model.py:149-176)
            return self.conv1(self.conv0(_act(self.act_type, self.norm(x)))) + x
        return self.norm(self.conv1(self.conv0(_act(self.act_type, x)))) + x      # build_res_block (:121-147)

    def forward_native(self, x):
        a = _act_arg(self.act_type)
        if self.pre_act:
            return self.conv1.forward_native(self.conv0.forward_native(_norm_native(self.norm, x, act=a))) + x
        return _norm_native(self.norm, self.conv1.forward_native(self.conv0.forward_native(x, act_in=a)), shortcut=x)


class _Upsample(nn.Module):
    def __init__(self, cin, cout, pad_type, nn_upsample, scale=2):
        super().__init__()
        self.scale, self.nn_upsample = scale, nn_upsample
        if nn_upsample:     # build_upsampling_nnconv (:178-193): tile x scale, then a (2*scale-1)^2 conv, no bias
            self.conv = PadConv2D(cin, cout, 2 * scale - 1, pad_type, False, False)
        else:               # build_upsampling_deconv (:195-202)
            self.conv = nn.ConvTranspose2d(cin, cout, 2 * scale, stride=scale, padding=scale // 2, bias=False)
            nn.init.kaiming_normal_(self.conv.weight, mode="fan_in", nonlinearity="relu")

    def forward(self, x):
        if self.nn_upsample:
            x = F.interpolate(x, scale_factor=self.scale, mode="nearest")
        return self.conv(x)

    def forward_native(self, x, act_in=None):
        """Tile x2 + pad 1 + 3x3 conv == one 3x3 conv of the LOW-resolution map with four phase filters
        (gen_ops.upsample_weights) + a pixel shuffle. One pixel of SYMMETRIC or REFLECT padding of the tiled map both
        read low-resolution edge pixel 0 (tiled rows 0 and 1 are copies of it): edge replication; CONSTANT stays zero."""
        from . import gen_ops
        if not self.nn_upsample or self.scale != 2:
            raise NotImplementedError("native generator route: nearest-neighbour x2 up-sampling only")
        mode = "zero" if self.conv.pad_type == "CONSTANT" else "symmetric"
        if mode == "zero" and act_in is not None:
            x, act_in = _norm_native(None, x, act=act_in), None
        w = self.conv.conv.weight
        y = gen_ops.conv3x3(x, gen_ops.upsample_weights(w), None, mode, act_in if act_in is not None else False)
        return gen_ops.pixel_shuffle_nhwc(y, w.shape[0])


class _Stage(nn.Module):
    def __init__(self, n_block, pre_act, nn_upsample, pad_type, norm_type, act_type, grad_ksize):
        super().__init__()
        self.layers = list(LAYER_CHANNELS.keys())
        self.pre_act, self.act_type = pre_act, act_type
        self.grad_conv1, self.grad_conv2 = nn.ModuleDict(), nn.ModuleDict()
        self.up, self.pre_norm, self.post_norm, self.res = nn.ModuleDict(), nn.ModuleDict(), nn.ModuleDict(), nn.ModuleDict()
        deeper = None
        for layer in reversed(self.layers):
            chan = LAYER_CHANNELS[layer]
            self.grad_conv1[layer] = PadConv2D(chan, chan, grad_ksize, pad_type, True, False, norm_type, act_type)
            self.grad_conv2[layer] = PadConv2D(chan, chan, grad_ksize, pad_type, False, False, norm_type, act_type)
            if deeper is not None:
                self.up[layer] = _Upsample(deeper, chan, pad_type, nn_upsample)
                if pre_act:
                    self.pre_norm[layer] = _norm(norm_type, deeper)
            if not pre_act:
                self.post_norm[layer] = _norm(norm_type, chan)
            self.res[layer] = nn.ModuleList([_ResBlock(chan, pad_type, norm_type, act_type, pre_act)
                                             for _ in range(n_block)])
            deeper = chan
        self.last_norm = _norm(norm_type, deeper) if pre_act else None
        self.convlast = PadConv2D(deeper, 3, 3, pad_type, False, True)

    def forward(self, grad_per_layer):
        delta = None
        for layer in reversed(self.layers):
            g = grad_per_layer[layer].permute(0, 3, 1, 2)
            g = self.grad_conv2[layer](self.grad_conv1[layer](g))
            if delta is None:
                delta = g
            else:
                delta = _act(self.act_type, self.pre_norm[layer](delta) if self.pre_act else delta)
                delta = g + self.up[layer](delta)
            if not self.pre_act:
                delta = self.post_norm[layer](delta)
            for blk in self.res[layer]:
                delta = blk(delta)
        delta = _act(self.act_type, self.last_norm(delta) if self.pre_act else delta)
        return self.convlast(delta).permute(0, 2, 3, 1)

    def native_supported(self, batch):
        return stage_native_supported(self, batch)

    def forward_native(self, grad_per_layer):
        return stage_forward_native(self, grad_per_layer)


def stage_native_supported(stage, batch):
    """Does the hand-written route cover this generator stage (shared by the adaptive and improved mirrors)?"""
    for m in stage.modules():
        if isinstance(m, PadConv2D) and m.ksize != 3:
            return False
        if isinstance(m, _Upsample) and (not m.nn_upsample or m.scale != 2):
            return False
        if isinstance(m, nn.BatchNorm2d):
            return False
    return True


def stage_forward_native(stage, grad_per_layer):
    """The graph of _Stage.forward on the kernels of po/gen_ops.py: activations fused into the padding pass of the
    convolution that follows or into the normalisation that precedes, shortcut adds fused into the normalisation."""
    a = _act_arg(stage.act_type)
    delta = None
    for layer in reversed(stage.layers):
        g = grad_per_layer[layer].to(torch.bfloat16)
        g = stage.grad_conv2[layer].forward_native(stage.grad_conv1[layer].forward_native(g))
        if delta is None:
            delta = g
        elif stage.pre_act:
            delta = g + stage.up[layer].forward_native(_norm_native(stage.pre_norm[layer], delta, act=a))
        else:
            delta = g + stage.up[layer].forward_native(delta, act_in=a)
        if not stage.pre_act:
            delta = _norm_native(stage.post_norm[layer], delta)
        for blk in stage.res[layer]:
            delta = blk.forward_native(delta)
    if stage.pre_act:
        out = stage.convlast.forward_native(_norm_native(stage.last_norm, delta, act=a))
    else:
        out = stage.convlast.forward_native(delta, act_in=a)
    return out.to(torch.float32)


def run_generator(model, stage, inputs):
    """Route a stage through the hand-written kernels or stock torch (`generator` = auto / native / torch, as in
    po/progressive_model.py)."""
    first = next(iter(inputs.values()))
    ok = first.is_cuda and stage.native_supported(first.shape[0])
    if model._generator == "native":
        if not ok:
            raise NotImplementedError("generator='native': this configuration is not covered by the hand-written "
                                      "kernels (needs CUDA, 3x3 gradient convolutions, nearest-neighbour up-sampling, "
                                      "instance / no normalisation or batch statistics at batch 1)")
        return stage.forward_native(inputs)
    if model._generator == "auto" and ok and model._door_is_cuda:
        return stage.forward_native(inputs)
    return stage(inputs)


class AdaptiveSynTex(SynTexModelDesc, nn.Module):
    def __init__(self, args, texture_fn=None):
        nn.Module.__init__(self)
        SynTexModelDesc.__init__(self, args)
        self._image_size = args.get("image_size") or IMAGESIZE
        # The loss is averaged among gpus. So we need to scale the lr accordingly. (adaptive_model.py:61-66)
        n_gpu = args.get("n_gpu") or 1
        batch_size = (args.get("batch_size") or BATCH)
        equi_batch_size = max(n_gpu, 1) * batch_size
        self._lr = (args.get("lr") or LR) * equi_batch_size
        self._beta1 = args.get("beta1") or BETA1
        self._beta2 = args.get("beta2") or BETA2
        self._n_stage = args.get("n_stage") or 1
        self._n_block = args.get("n_block") or 2
        act = args.get("act", "sigmoid")
        if act == "sigmoid":
            self._act = torch.sigmoid
        elif act == "tanh":
            self._act = torch.tanh
        elif act == "identity":
            self._act = lambda x: x
        else:
            raise ValueError("Invalid activation: " + str(type(act)))
        self._act_name = act
        self._loss_scale = args.get("loss_scale", 1.)
        self._pre_act = args.get("pre_act") or False
        self._nn_upsample = not (args.get("deconv_upsample") or False)
        pad_type = args.get("pad_type") or "reflect"
        self._pad_type = "CONSTANT" if pad_type == "zero" else pad_type.upper()
        self._norm_type = args.get("norm_type") or "instance"
        self._act_type = args.get("act_type") or "relu"
        self._grad_ksize = args.get("grad_ksize") or 3
        self._stop_grad = args.get("stop_grad") or False
        self._generator = args.get("generator", "auto")       # extension: auto / native / torch (po/gen_ops.py)
        if self._generator not in ("auto", "native", "torch"):
            raise ValueError("generator must be auto, native or torch")
        self._door_is_cuda = texture_fn is None
        self._texture_fn = texture_fn or CudaTexture(self._vgg19, SynTexModelDesc.DEFAULT_COEFS.keys())
        self.syn = nn.ModuleList([_Stage(self._n_block, self._pre_act, self._nn_upsample, self._pad_type,
                                         self._norm_type, self._act_type, self._grad_ksize)
                                  for _ in range(self._n_stage)])
        self.image_outputs, self.losses, self.loss_layer_output = [], [], OrderedDict()

    def inputs(self):
        s = self._image_size
        return [((None, s, s, 3), np.float32, "pre_image_input"), ((None, s, s, 3), np.float32, "image_target")]

    @property
    def vars(self):
        return list(self.syn.parameters())

    @staticmethod
    def get_parser(ps=None):
        ps = SynTexModelDesc.get_parser(ps)
        ps = ArgParser(ps, name="progressive")
        ps.add("--image-size", type=int, default=IMAGESIZE)
        ps.add("--lr", type=float, default=LR)
        ps.add("--beta1", type=float, default=BETA1)
        ps.add("--beta2", type=float, default=BETA2)
        ps.add("--n-stage", type=int, default=1, help="number of unfolding")
        ps.add("--n-block", type=int, default=2, help="number of res blocks in each scale.")
        ps.add("--act", type=str, default="sigmoid", choices=["sigmoid", "identity"])
        ps.add("--loss-scale", type=float, default=1.)
        ps.add("--pad-type", type=str, default="reflect", choices=["reflect", "zero", "symmetric"])
        ps.add("--grad-ksize", type=int, default=3)
        ps.add("--norm-type", type=str, default="instance", choices=["instance", "batch", "none"])
        ps.add("--act-type", type=str, default="relu", choices=["relu", "lrelu"])
        ps.add_flag("--pre-act")
        ps.add_flag("--deconv-upsample")
        ps.add_flag("--stop-grad")
        ps.add("--n-gpu", type=int, default=1)
        return ps

    def build_stage_preparation(self, image_input, gram_target, coefs, name="prep", slot=0):
        """adaptive_model.py:204-222: (feat_input, loss_input [B], loss_per_layer, grad_per_layer); here loss_input
        sums ALL the layers of `coefs` (improved_model.py excludes the last one)."""
        feat_input, _, loss_per_layer, grad_per_layer = self._texture_fn.prep(image_input, gram_target, coefs,
                                                                              self._stop_grad, slot)
        layers = list(coefs.keys())
        loss_input = sum(coefs[k] * 1. / 4 * loss_per_layer[k] for k in layers)
        return feat_input, loss_input, loss_per_layer, grad_per_layer

    def build_stage(self, pre_image_input, gram_target, n_loss_layer, name="stage"):
        """adaptive_model.py:224-311."""
        s = n_loss_layer - 1
        coefs = SynTexModelDesc.DEFAULT_COEFS
        image_input = self._act(pre_image_input)
        slot = s if torch.is_grad_enabled() else 0
        feat_input, loss_input, loss_per_layer, grad_per_layer = \
            self.build_stage_preparation(image_input, gram_target, coefs, slot=slot)
        delta_input = run_generator(self, self.syn[s], grad_per_layer)
        pre_image_output = pre_image_input + delta_input
        return image_input, loss_input, loss_per_layer, pre_image_output

    def build_graph(self, pre_image_input, image_target):
        """adaptive_model.py:313-361."""
        dev = next(self.syn.parameters()).device
        dtype = next(self.syn.parameters()).dtype
        pre_image_input = torch.as_tensor(pre_image_input).to(device=dev, dtype=dtype)
        image_target = torch.as_tensor(image_target).to(device=dev, dtype=dtype) / 255.
        with torch.no_grad():
            gram_target = self._texture_fn.grams(image_target)
        self.image_outputs, self.losses = [], []
        pre_image_output = pre_image_input
        for s in range(self._n_stage):
            image_input, loss_overall_input, _, pre_image_output = \
                self.build_stage(pre_image_output, gram_target, s + 1, name="stage%d" % s)
            self.image_outputs.append(image_input)
            self.losses.append(loss_overall_input.mean())
        image_output = self._act(pre_image_output)
        layers = list(SynTexModelDesc.DEFAULT_COEFS.keys())
        loss_layer_output, _ = self._texture_fn(image_output, layers, gram_target,
                                                self._n_stage if torch.is_grad_enabled() else 0, False)
        loss_overall_output = sum(SynTexModelDesc.DEFAULT_COEFS[k] * 1. / 4 * loss_layer_output[k] for k in layers)
        self.image_outputs.append(image_output)
        self.losses.append(loss_overall_output.mean())
        self.loss_layer_output = OrderedDict((k, v.mean()) for k, v in loss_layer_output.items())
        weights = [1.]
        for _ in range(len(self.losses) - 1):
            weights.append(weights[-1] * self._loss_scale)
        # skip the first loss as it is computed from noise
        self.loss = sum(weights[i] * loss for i, loss in enumerate(reversed(self.losses[1:])))
        return self.loss

    def optimizer(self):
        return torch.optim.Adam(self.vars, lr=self._lr, betas=(self._beta1, self._beta2), eps=1e-3)
