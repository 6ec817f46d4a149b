"""ImprovedSynTex - the `ProgressiveSynTex` class of po/improved_model.py:57-414 - over the CUDA hot path.

Five progressive stages like po/progressive_model.py, but stage s is fed `tf.gradients(loss, [feat_k])` of its first
s+1 Gram layers - a full VGG19 back-propagation to every Gram layer (`build_stage_preparation`, :209-232) - and is
trained THROUGH those gradients unless --stop-grad: the double backward through VGG19 of
`po.texture_op.vgg_stage_preparation` (tangent sweep + closed-form injection). The generator
differs from progressive_model.py in its options (padding reflect / symmetric / zero, instance / batch / no
normalisation - BatchNorm always with batch statistics, :353-360 -, ReLU / LeakyReLU(0.2), gradient-conv kernel size,
3x3 convolution after the nearest-neighbour up-sampling, :185-199) and shares its layer modules with the
AdaptiveSynTex mirror (po/adaptive_model.py); it runs on the hand-written kernels of po/gen_ops.py (reflect /
symmetric / zero padding, InstanceNorm or batch-1 BatchNorm, ReLU / LeakyReLU fused) or on stock torch. tensorpack layer defaults restated from
its published source, unverified.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..aparse import ArgParser
from .adaptive_model import (PadConv2D, _ResBlock, _Upsample, _act, _norm, run_generator, stage_forward_native,
                             stage_native_supported)
from .model import SynTexModelDesc
from .progressive_model import LAYER_CHANNELS
from .texture_op import CudaTexture

MAX_EPOCH = 200
STEPS_PER_EPOCH = 4000
IMAGESIZE = 224
LR = 2e-4
BETA1 = 0.5
BETA2 = 0.999
BATCH = 1
TEST_BATCH = 1


class _BatchStatsNorm(nn.Module):
    """tensorpack BatchNorm under argscope(BatchNorm, training=True) (improved_model.py:353-360): batch statistics in
    every phase, biased variance, eps 1e-5, affine (with the per-GPU batch of 1 the reference trains with, this is
    InstanceNorm). Plain tensor arithmetic rather than nn.BatchNorm2d with its running statistics stripped: that
    combination failed inside cuDNN on the B200 (CUDNN_STATUS_EXECUTION_FAILED_CUDART on this generator's permuted
    views), while nn.BatchNorm2d as shipped - which the adaptive mirror uses - works."""

    def __init__(self, chan, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(chan))
        self.bias = nn.Parameter(torch.zeros(chan))

    def forward(self, x):
        mean = x.mean(dim=(0, 2, 3), keepdim=True)
        var = x.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
        return (x - mean) * torch.rsqrt(var + self.eps) * self.weight.view(1, -1, 1, 1) + self.bias.view(1, -1, 1, 1)


def _always_batch_stats(module):
    """Swap every nn.BatchNorm2d the shared layer modules (po/adaptive_model.py) created for _BatchStatsNorm."""
    for name, child in list(module.named_children()):
        if isinstance(child, nn.BatchNorm2d):
            setattr(module, name, _BatchStatsNorm(child.num_features, child.eps))
        else:
            _always_batch_stats(child)
    return module


class _Stage(nn.Module):
    """The generator of one stage (improved_model.py:234-318) over the first `n_loss_layer` Gram layers."""

    def __init__(self, n_loss_layer, n_block, pre_act, nn_upsample, pad_type, norm_type, act_type, grad_ksize):
        super().__init__()
        self.layers = list(LAYER_CHANNELS.keys())[:n_loss_layer]
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
        _always_batch_stats(self)

    def forward(self, grad_per_layer):
        """grad_per_layer: {layer: [B,h,w,C]} (NHWC) -> delta [B,H,W,3]."""
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


class ImprovedSynTex(SynTexModelDesc, nn.Module):
    def __init__(self, args, texture_fn=None):
        nn.Module.__init__(self)
        SynTexModelDesc.__init__(self, args)
        self._image_size = args.get("image_size") or IMAGESIZE
        # The loss is averaged among gpus. So we need to scale the lr accordingly. (improved_model.py:60-66)
        n_gpu = args.get("n_gpu") or 1
        batch_size = (args.get("batch_size") or BATCH)
        equi_batch_size = max(n_gpu, 1) * batch_size
        self._lr = (args.get("lr") or LR) * equi_batch_size
        self._beta1 = args.get("beta1") or BETA1
        self._beta2 = args.get("beta2") or BETA2
        self._n_stage = args.get("n_stage") or 5
        assert self._n_stage == 5, "Num of stages is 5 for progressive model"
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
        self._sync_stats = args.get("sync_stats")          # accepted; the reference has the sync variant commented out
        self._generator = args.get("generator", "auto")       # extension: auto / native / torch (po/gen_ops.py)
        if self._generator not in ("auto", "native", "torch"):
            raise ValueError("generator must be auto, native or torch")
        self._door_is_cuda = texture_fn is None
        self._texture_fn = texture_fn or CudaTexture(self._vgg19, SynTexModelDesc.DEFAULT_COEFS.keys())
        self.syn = nn.ModuleList([_Stage(s + 1, self._n_block, self._pre_act, self._nn_upsample, self._pad_type,
                                         self._norm_type, self._act_type, self._grad_ksize)
                                  for s in range(self._n_stage)])
        self.image_outputs, self.losses, self.loss_layer_output = [], [], OrderedDict()

    def inputs(self):
        s = self._image_size
        return [((None, s, s, 3), np.float32, "pre_image_input"), ((None, s, s, 3), np.float32, "image_target")]

    @property
    def vars(self):
        return list(self.syn.parameters())

    @staticmethod
    def get_parser(ps=None):
        """improved_model.py:95-122 (note the parser defaults differ from the constructor's fall-backs, as there)."""
        ps = SynTexModelDesc.get_parser(ps)
        ps = ArgParser(ps, name="progressive")
        ps.add("--image-size", type=int, default=IMAGESIZE)
        ps.add("--lr", type=float, default=LR)
        ps.add("--beta1", type=float, default=BETA1)
        ps.add("--beta2", type=float, default=BETA2)
        ps.add("--n-stage", type=int, default=5, help="This argument is fixed to 5.")
        ps.add("--n-block", type=int, default=2, help="number of res blocks in each scale.")
        ps.add("--act", type=str, default="sigmoid", choices=["sigmoid", "identity"])
        ps.add("--loss-scale", type=float, default=1.)
        ps.add("--pad-type", type=str, default="zero", choices=["reflect", "zero", "symmetric"])
        ps.add("--grad-ksize", type=int, default=3)
        ps.add("--norm-type", type=str, default="batch", choices=["instance", "batch", "none"])
        ps.add("--act-type", type=str, default="lrelu", choices=["relu", "lrelu"])
        ps.add_flag("--pre-act")
        ps.add_flag("--deconv-upsample")
        ps.add_flag("--stop-grad")
        ps.add("--n-gpu", type=int, default=1)
        ps.add("--sync-stats", type=str, choices=["nccl", "horovod"])
        return ps

    def build_stage_preparation(self, image_input, gram_target, coefs, name="prep", slot=0):
        """improved_model.py:209-232: (feat_input, loss_input [B], loss_per_layer, grad_per_layer). loss_input leaves
        out the LAST layer of `coefs` ("the input is not computed from the gradient of the last layer"), while the
        gradients handed to the generator come from the loss over all of them."""
        return self._texture_fn.prep(image_input, gram_target, coefs, self._stop_grad, slot)

    def build_stage(self, pre_image_input, gram_target, n_loss_layer, name="stage"):
        """improved_model.py:234-326."""
        s = n_loss_layer - 1
        coefs = OrderedDict()
        for k in list(SynTexModelDesc.DEFAULT_COEFS.keys())[:n_loss_layer]:
            coefs[k] = SynTexModelDesc.DEFAULT_COEFS[k]
        image_input = self._act(pre_image_input)
        slot = s if torch.is_grad_enabled() else 0
        _, loss_input, loss_per_layer, grad_per_layer = \
            self.build_stage_preparation(image_input, gram_target, coefs, slot=slot)
        delta_input = run_generator(self, self.syn[s], grad_per_layer)
        if self._stop_grad:
            pre_image_output = pre_image_input.detach() + delta_input
        else:
            pre_image_output = pre_image_input + delta_input
        return image_input, loss_input, loss_per_layer, pre_image_output

    def build_graph(self, pre_image_input, image_target):
        """improved_model.py:328-393. pre_image_input: values before the activation; image_target in [0, 255]."""
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
            self.losses.append(torch.as_tensor(loss_overall_input, device=dev, dtype=dtype).mean())
        image_output = self._act(pre_image_output)
        layers = list(SynTexModelDesc.DEFAULT_COEFS.keys())
        loss_layer_output, _ = self._texture_fn(image_output, layers, gram_target,
                                                self._n_stage if torch.is_grad_enabled() else 0, False)
        loss_overall_output = sum(SynTexModelDesc.DEFAULT_COEFS[k] * 1. / 4 * loss_layer_output[k] for k in layers)
        self.image_outputs.append(image_output)
        self.losses.append(loss_overall_output.mean())
        self.loss_layer_output = OrderedDict((k, v.mean()) for k, v in loss_layer_output.items())
        # average losses from all stages; skip the first loss as it is computed from noise
        weights = [1.]
        for _ in range(len(self.losses) - 1):
            weights.append(weights[-1] * self._loss_scale)
        self.loss = sum(weights[i] * loss for i, loss in enumerate(reversed(self.losses[1:])))
        return self.loss

    def optimizer(self):
        """improved_model.py:395-398: Adam(lr, beta1, beta2, epsilon=1e-3)."""
        return torch.optim.Adam(self.vars, lr=self._lr, betas=(self._beta1, self._beta2), eps=1e-3)

    def viz(self, target=None):
        """The 'stages-target/viz' image (improved_model.py:333-345): uint8 [B,H,(n+1)*W,3], rounded."""
        ims = list(self.image_outputs) + ([target] if target is not None else [])
        im = torch.cat([i.detach() for i in ims], dim=2)
        im = (im + 1.0) * 127.5 if self._act_name == "tanh" else im * 255
        return im.clamp(0, 255).round().to(torch.uint8)


def test_only(model, data, test_folder=None):
    """improved_model.py:453-484 (--test-only): run the feed-forward synthesis over `data` (an iterable of
    (pre_image_input, image_target) batches), return ([viz uint8 arrays], [loss_output]) and, when `test_folder` is
    given, save test-<idx>.jpg there (PIL; imageio is not in this image)."""
    import os
    outputs, losses = [], []
    idx = 0
    if test_folder is not None and not os.path.exists(test_folder):
        os.makedirs(test_folder)
    for pii, it in data:
        with torch.no_grad():
            model.build_graph(pii, it)
        target = torch.as_tensor(it).to(model.image_outputs[-1]) / 255.
        arr = model.viz(target).cpu().numpy()
        outputs.append(arr)
        losses.append(float(model.losses[-1]))
        if test_folder is not None:
            from PIL import Image
            for i in range(arr.shape[0]):
                Image.fromarray(arr[i]).save(os.path.join(test_folder, "test-%d.jpg" % idx))
                idx += 1
    return outputs, losses
