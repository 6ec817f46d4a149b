"""ProgressiveSynTex (po/progressive_model.py:34-272) over the CUDA hot path.

Five stages; stage s looks at the current image through VGG19, takes the closed-form style gradients of the first
s+1 Gram layers (build_texture_loss(calc_grad=True), utils.py:50-55) and lets a small generator turn them into a
pre-image update. Everything VGG / Gram - forward, the gradients handed to the generator, and the back-propagation
of the training loss through them (first- and second-order) - runs on the hand-written kernels through
`po.texture_op.vgg_texture`; the generator (row f-1 of SURVEY.md section 8) runs on the hand-written kernels of
po/gen_ops.py (`generator="native"`, the default on the CUDA door) or on stock torch (`generator="torch"`).
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..aparse import ArgParser
from .layers import PreResBlock, ResBlock, SymConv, UpsampleDeconv, UpsampleNN, instance_norm
from .model import SynTexModelDesc
from .texture_op import CudaTexture

IMAGESIZE = 224
LR = 5e-5
BETA1 = 0.9
BETA2 = 0.999
BATCH = 1
TEST_BATCH = 1

LAYER_CHANNELS = OrderedDict([("conv1_1", 64), ("pool1", 64), ("pool2", 128), ("pool3", 256), ("pool4", 512)])


class _Stage(nn.Module):
    """The generator of one stage (progressive_model.py:161-206): per Gram layer two gradient convs, merge with the
    up-sampled deeper branch, InstanceNorm, n_block residual blocks; then ReLU + 3x3 conv to an RGB update."""

    def __init__(self, n_loss_layer, n_block, pre_act, nn_upsample, zero_pad_inputs=False):
        super().__init__()
        self.layers = list(LAYER_CHANNELS.keys())[:n_loss_layer]
        self.pre_act = pre_act
        block = PreResBlock if pre_act else ResBlock
        up = UpsampleNN if nn_upsample else UpsampleDeconv
        self.grad_conv1 = nn.ModuleDict()
        self.grad_conv2 = nn.ModuleDict()
        self.up = nn.ModuleDict()
        self.pre_inorm = nn.ModuleDict()
        self.post_inorm = nn.ModuleDict()
        self.res = nn.ModuleDict()
        deeper = None
        for layer in reversed(self.layers):
            chan = LAYER_CHANNELS[layer]
            # single_model.py:170-171 feeds the features through plain Conv2D (SAME zero padding); the progressive
            # model pads its gradient inputs symmetrically (progressive_model.py:171-175)
            self.grad_conv1[layer] = SymConv(chan, chan, 3, act=True, zero_pad=zero_pad_inputs)
            self.grad_conv2[layer] = SymConv(chan, chan, 3, act=False, zero_pad=zero_pad_inputs)
            if deeper is not None:
                self.up[layer] = up(deeper, chan)
                if pre_act:
                    self.pre_inorm[layer] = instance_norm(deeper)
            if not pre_act:
                self.post_inorm[layer] = instance_norm(chan)
            self.res[layer] = nn.ModuleList([block(chan) for _ in range(n_block)])
            deeper = chan
        self.actlast_inorm = instance_norm(deeper) if pre_act else None
        self.convlast = SymConv(deeper, 3, 3, act=False, use_bias=True)

    def forward(self, grad_layer):
        """grad_layer: {layer: [B,h,w,C]} (NHWC, as the hot path produces them) -> delta [B,H,W,3]."""
        delta = None
        for layer in reversed(self.layers):
            g = grad_layer[layer].permute(0, 3, 1, 2)
            g = self.grad_conv2[layer](self.grad_conv1[layer](g))
            if delta is None:
                delta = g
            else:
                delta = F.relu(self.pre_inorm[layer](delta)) if self.pre_act else F.relu(delta)
                delta = g + self.up[layer](delta)
            if not self.pre_act:
                delta = self.post_inorm[layer](delta)
            for blk in self.res[layer]:
                delta = blk(delta)
        delta = F.relu(self.actlast_inorm(delta)) if self.pre_act else F.relu(delta)
        return self.convlast(delta).permute(0, 2, 3, 1)

    def native_supported(self):
        return all(hasattr(m, "forward_native") for m in self.up.values())

    def forward_native(self, grad_layer):
        """The same graph on the hand-written kernels (po/gen_ops.py): NHWC bf16 activations end to end, ReLUs fused
        into the padding pass of the convolution that follows, InstanceNorm + ReLU / + shortcut in one pass."""
        delta = None
        for layer in reversed(self.layers):
            g = grad_layer[layer].to(torch.bfloat16)
            g = self.grad_conv2[layer].forward_native(self.grad_conv1[layer].forward_native(g))
            if delta is None:
                delta = g
            elif self.pre_act:
                delta = g + self.up[layer].forward_native(self.pre_inorm[layer].forward_native(delta, relu=True))
            else:
                delta = g + self.up[layer].forward_native(delta, relu_in=True)
            if not self.pre_act:
                delta = self.post_inorm[layer].forward_native(delta)
            for blk in self.res[layer]:
                delta = blk.forward_native(delta)
        if self.pre_act:
            out = self.convlast.forward_native(self.actlast_inorm.forward_native(delta, relu=True))
        else:
            out = self.convlast.forward_native(delta, relu_in=True)
        return out.to(torch.float32)


class ProgressiveSynTex(SynTexModelDesc, nn.Module):
    def __init__(self, args, texture_fn=None):
        nn.Module.__init__(self)
        SynTexModelDesc.__init__(self, args)
        self._image_size = args.get("image_size", IMAGESIZE)
        # The loss is averaged among gpus. So we need to scale the lr accordingly. (progressive_model.py:36-37)
        self._lr = args.get("lr", LR) * max(args.get("n_gpu", 1), 1)
        self._beta1 = args.get("beta1", BETA1)
        self._beta2 = args.get("beta2", BETA2)
        self._n_stage = args.get("n_stage", 5)
        assert self._n_stage == 5, "Num of stages is 5 for progressive model"
        self._n_block = args.get("n_block", 2)
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
        self._pre_act = args.get("pre_act", False)
        self._nn_upsample = not args.get("deconv_upsample", False)
        # extension (not in the reference parser): run the generator's stock convolutions under bf16 autocast
        self._generator_bf16 = bool(args.get("generator_bf16", False))
        # extension: "native" = the generator on the hand-written kernels (po/gen_ops.py), "torch" = stock ATen/cuDNN,
        # "auto" = native when the door is the CUDA plan and the configuration is covered (nearest-neighbour up-sampling)
        self._generator = args.get("generator", "auto")
        if self._generator not in ("auto", "native", "torch"):
            raise ValueError("generator must be auto, native or torch")
        self._door_is_cuda = texture_fn is None
        # The hot path: an object with grams(image) and __call__(image, layers, gram_target, slot, calc_grad) ->
        # (loss_layer, grad_layer). Default: the CUDA plan; tests swap in the CPU oracle to check the training gradient.
        self._texture_fn = texture_fn or CudaTexture(self._vgg19, SynTexModelDesc.DEFAULT_COEFS.keys())
        self.syn = nn.ModuleList([_Stage(s + 1, self._n_block, self._pre_act, self._nn_upsample)
                                  for s in range(self._n_stage)])
        self.image_outputs = []
        self.losses = []
        self.loss_layer_output = OrderedDict()

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
        ps.add("--n-stage", type=int, default=5, help="This argument is fixed to 5.")
        ps.add("--n-block", type=int, default=2, help="number of res blocks in each scale.")
        ps.add("--act", type=str, default="sigmoid", choices=["sigmoid", "tanh", "identity"])
        ps.add("--loss-scale", type=float, default=1.)
        ps.add_flag("--pre-act")
        ps.add_flag("--deconv-upsample")
        ps.add("--n-gpu", type=int, default=1)
        return ps

    def build_stage(self, pre_image_input, gram_target, n_loss_layer, name="stage"):
        """progressive_model.py:145-206. Returns (image_input, loss_overall_input [B], loss_layer_input, pre_image_output)."""
        s = n_loss_layer - 1
        coefs = OrderedDict()
        for k in list(SynTexModelDesc.DEFAULT_COEFS.keys())[:n_loss_layer]:
            coefs[k] = SynTexModelDesc.DEFAULT_COEFS[k]
        image_input = self._act(pre_image_input)
        # every stage keeps its own plan alive until backward; inference can reuse one
        slot = s if torch.is_grad_enabled() else 0
        # the hand-written generator consumes the gradients in the kernels' own bf16: skip the fp32 round trip
        native = image_input.is_cuda and self._door_is_cuda and self.use_native_generator(self.syn[s])
        loss_layer_input, grad_layer = self._texture_fn(image_input, list(coefs.keys()), gram_target, slot,
                                                        "bf16" if native else True)
        loss_overall_input = sum(coefs[k] * 1. / 4 * loss_layer_input[k] for k in coefs)
        delta_input = self.run_generator(self.syn[s], grad_layer, image_input.is_cuda)
        pre_image_output = pre_image_input + delta_input
        return image_input, loss_overall_input, loss_layer_input, pre_image_output

    def use_native_generator(self, stage):
        if self._generator == "torch":
            return False
        ok = self._nn_upsample and stage.native_supported()
        if self._generator == "native":
            if not ok:
                raise NotImplementedError("generator='native' covers nearest-neighbour up-sampling only")
            return True
        return ok and self._door_is_cuda

    def run_generator(self, stage, inputs, on_cuda):
        if on_cuda and self.use_native_generator(stage):
            return stage.forward_native(inputs)
        if self._generator == "native":
            raise RuntimeError("generator='native' needs CUDA tensors; there is no CPU route")
        if self._generator_bf16 and on_cuda:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return stage(inputs).float()
        return stage(inputs)

    def build_graph(self, pre_image_input, image_target):
        """progressive_model.py:208-266. pre_image_input: values before the activation; image_target in [0, 255].
        Sets self.loss (training objective), self.losses, self.image_outputs, self.loss_layer_output."""
        pre_image_input = torch.as_tensor(pre_image_input).to(dtype=torch.float32)
        image_target = torch.as_tensor(image_target).to(dtype=torch.float32) / 255.
        dev = next(self.syn.parameters()).device
        pre_image_input, image_target = pre_image_input.to(dev), image_target.to(dev)
        with torch.no_grad():
            gram_target = self._texture_fn.grams(image_target)
        self.image_outputs = list()
        self.losses = list()
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
        # average losses from all stages; skip the first loss as it is computed from noise
        weights = [1.]
        for _ in range(len(self.losses) - 1):
            weights.append(weights[-1] * self._loss_scale)
        self.loss = sum(weights[i] * loss for i, loss in enumerate(reversed(self.losses[1:])))
        return self.loss

    def optimizer(self):
        """progressive_model.py:268-271: Adam(lr, beta1, beta2, epsilon=1e-3)."""
        return torch.optim.Adam(self.vars, lr=self._lr, betas=(self._beta1, self._beta2), eps=1e-3)

    def viz(self, target=None):
        """The 'stages-target/viz' summary image (progressive_model.py:222-233): uint8 [B,H,(n+1)*W,3]."""
        ims = list(self.image_outputs) + ([target] if target is not None else [])
        im = torch.cat([i.detach() for i in ims], dim=2)
        im = (im + 1.0) * 128 if self._act_name == "tanh" else im * 256
        return im.clamp(0, 255).to(torch.uint8)
