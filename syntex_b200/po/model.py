"""Host mirror of the hot-path doors of po/model.py: SynTexModelDesc._build_extractor / _build_loss.

The reference class is a tensorpack ModelDescBase; only its two entry points into the VGG/Gram path
(po/model.py:51-69) and DEFAULT_COEFS (po/model.py:27-33) are on the hot path this package replaces. The
generator networks, trainer and dataflow (tensorpack glue) are outside it.
"""
from collections import OrderedDict

import numpy as np

from ..aparse import ArgParser
from ..texture_utils import Vgg19Extractor, build_gram, build_texture_loss


class SynTexModelDesc(object):
    """Texture Synthesis

    Attributes
    ----------
    loss : tensor
        Need to define the "loss" in the "build_graph" method.
    """
    DEFAULT_COEFS = OrderedDict([   # po/model.py:27-33
        ("conv1_1", 1e5),
        ("pool1", 1e5),
        ("pool2", 1e5),
        ("pool3", 1e5),
        ("pool4", 1e5)
    ])

    def __init__(self, args):
        self._vgg19 = Vgg19Extractor(args.get("vgg19_npy_path"))
        self.loss = None

    @staticmethod
    def get_parser(ps=None):
        ps = ArgParser(ps, name="syntex-model")
        ps.add("--vgg19-npy-path")
        return ps

    def _build_loss(self, image_output, gram_target, calc_grad=False, name="texture_loss"):
        """po/model.py:51-58: VGG features of image_output + texture loss against gram_target."""
        feat_output = self._vgg19.build(image_output)
        loss_overall, loss_layer, calc_grad = build_texture_loss(
            feat_output, gram_target, SynTexModelDesc.DEFAULT_COEFS,
            calc_grad=calc_grad, name="loss"
        )
        return loss_overall, loss_layer, calc_grad

    def _build_extractor(self, image_input, calc_gram=True, name="ext"):
        """po/model.py:60-69: VGG features (+ the five Gram matrices)."""
        feat_input = self._vgg19.build(image_input)
        if calc_gram:
            gram_input = OrderedDict()
            for k in SynTexModelDesc.DEFAULT_COEFS:
                gram_input[k] = build_gram(feat_input[k])
        else:
            gram_input = None
        return feat_input, gram_input

    def build_graph(self, *inputs):
        pass


class RandomZData(object):
    """po/model.py:138-147: endless uniform noise source."""

    def __init__(self, shape, minv=0., maxv=1., seed=None):
        self.shape = shape
        self.minv = minv
        self.maxv = maxv
        # the reference draws from the global numpy state; `seed` (an extension) gives a private, reproducible stream
        self.rng = np.random if seed is None else np.random.RandomState(seed)

    def __iter__(self):
        while True:
            yield [self.rng.uniform(self.minv, self.maxv, size=self.shape)]
