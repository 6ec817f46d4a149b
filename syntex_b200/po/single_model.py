"""SingleSynTex (po/single_model.py:33-262) over the CUDA hot path - the model BASELINE.json's config 3 names.

The generator is fed the VGG19 FEATURES of the current image (no explicit gradients: "that information is implicitly
provided by final loss", single_model.py:157-158). As shipped the reference loops over every recorded layer of the
current extractor (`for layer in reversed(feat_input)`, :165) and cannot build (its up-sampling chain assumes the
five Gram layers; po/README.md:58 marks the file stale): this mirror follows the diagram at single_model.py:156-162
- the five DEFAULT_COEFS layers. Training back-propagates through the features, i.e. through VGG19
(`texture_op.vgg_stage_preparation`, direct feature gradients); the generator runs on the hand-written kernels of
po/gen_ops.py (or stock torch: `generator="torch"`).
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from ..aparse import ArgParser
from .model import SynTexModelDesc
from .progressive_model import ProgressiveSynTex, _Stage
from .texture_op import CudaTexture

IMAGESIZE = 224
LR = 5e-5
BETA1 = 0.9
BETA2 = 0.999
BATCH = 1
TEST_BATCH = 32


class SingleSynTex(SynTexModelDesc, nn.Module):
    def __init__(self, args, texture_fn=None):
        nn.Module.__init__(self)
        SynTexModelDesc.__init__(self, args)
        self._image_size = args.get("image_size", IMAGESIZE)
        self._lr = args.get("lr", LR)
        self._beta1 = args.get("beta1", BETA1)
        self._beta2 = args.get("beta2", BETA2)
        self._n_stage = args.get("n_stage", 1)
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
        self._loss_scale = args.get("loss_scale", 0.)
        self._pre_act = args.get("pre_act", False)
        self._nn_upsample = not args.get("deconv_upsample", False)
        # generator route, as in ProgressiveSynTex: hand-written kernels ("native") or stock torch
        self._generator_bf16 = bool(args.get("generator_bf16", False))
        self._generator = args.get("generator", "auto")
        if self._generator not in ("auto", "native", "torch"):
            raise ValueError("generator must be auto, native or torch")
        self._door_is_cuda = texture_fn is None
        self._texture_fn = texture_fn or CudaTexture(self._vgg19, SynTexModelDesc.DEFAULT_COEFS.keys())
        self.syn = nn.ModuleList([_Stage(5, self._n_block, self._pre_act, self._nn_upsample, zero_pad_inputs=True)
                                  for _ in range(self._n_stage)])
        self.image_outputs, self.losses, self.loss_layer_output = [], [], OrderedDict()

    use_native_generator = ProgressiveSynTex.use_native_generator
    run_generator = ProgressiveSynTex.run_generator

    def inputs(self):
        s = self._image_size
        return [((None, s, s, 3), np.float32, "pre_image_input"), ((None, s, s, 3), np.float32, "image_target")]

    @property
    def vars(self):
        return list(self.syn.parameters())

    @staticmethod
    def get_parser(ps=None):
        ps = SynTexModelDesc.get_parser(ps)
        ps = ArgParser(ps, name="single")
        ps.add("--image-size", type=int, default=IMAGESIZE)
        ps.add("--lr", type=float, default=LR)
        ps.add("--beta1", type=float, default=BETA1)
        ps.add("--beta2", type=float, default=BETA2)
        ps.add("--n-stage", type=int, default=1)
        ps.add("--n-block", type=int, default=2, help="number of res blocks in each scale.")
        ps.add("--act", type=str, default="sigmoid", choices=["sigmoid", "tanh", "identity"])
        ps.add("--loss-scale", type=float, default=0.)
        ps.add_flag("--pre-act")
        ps.add_flag("--deconv-upsample")
        return ps

    def build_stage(self, pre_image_input, gram_target, name="stage", stage_index=0):
        """single_model.py:143-197."""
        coefs = SynTexModelDesc.DEFAULT_COEFS
        image_input = self._act(pre_image_input)
        slot = stage_index if torch.is_grad_enabled() else 0
        feat_input, _, loss_layer_input, _ = self._texture_fn.prep(image_input, gram_target, coefs, True, slot,
                                                                   need_grads=False)
        loss_overall_input = sum(coefs[k] * 1. / 4 * loss_layer_input[k] for k in coefs)
        delta_input = self.run_generator(self.syn[stage_index], feat_input, image_input.is_cuda)
        pre_image_output = pre_image_input + delta_input
        return image_input, loss_overall_input, loss_layer_input, pre_image_output

    def build_graph(self, pre_image_input, image_target):
        """single_model.py:199-255."""
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
                self.build_stage(pre_image_output, gram_target, name="stage%d" % s, stage_index=s)
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

    @torch.no_grad()
    def synthesize(self, pre_image_input, image_target):
        """Feed-forward synthesis only: the output image [B,H,W,3] in [0,1], without the loss of the result that
        build_graph evaluates for the summaries (one VGG19 forward less per call; the reference's predictor fetches
        'stages-target/viz' and 'loss_output' from the same graph, po/improved_model.py:453-484)."""
        dev = next(self.syn.parameters()).device
        dtype = next(self.syn.parameters()).dtype
        pre = torch.as_tensor(pre_image_input).to(device=dev, dtype=dtype)
        target = torch.as_tensor(image_target).to(device=dev, dtype=dtype) / 255.
        gram_target = self._texture_fn.grams(target)
        for s in range(self._n_stage):
            _, _, _, pre = self.build_stage(pre, gram_target, name="stage%d" % s, stage_index=s)
        return self._act(pre)
