"""CPU oracle for the SynTex hot path — TEST INFRASTRUCTURE ONLY.

This is a restatement, in torch-CPU float64/float32, of the arithmetic the reference delegates to
TensorFlow 1.12 (absent from /root/reference and not installable here: Python 3.12, no network) and
scipy's L-BFGS-B. It exists to check the CUDA path and to serve as the timed CPU baseline. Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the
product package `syntex_b200` never does.

PARITY UNPINNED: the reference ships no tests, golden vectors, weights or fixtures for this path
 and cannot be executed in this container, so this oracle is anchored on the reference's
call sites and on TF's documented op semantics, not on reference outputs:

  texture_utils/vgg19.py:9-10     VGG_MEAN (BGR) / VGG_MEAN_RGB
  texture_utils/vgg19.py:11-30    layer shapes up to pool4
  texture_utils/vgg19.py:90-91    avg_pool 2x2 stride 2 SAME
  texture_utils/vgg19.py:124-127  conv1_1 input-channel flip for RGB input
  texture_utils/vgg19.py:142-146  x*255 - mean
  texture_utils/vgg19.py:157-167  layer loop, every layer recorded until `topmost`
  texture_utils/vgg19.py:171-179  relu(conv2d(x, W, stride 1, SAME) + b)   (cross-correlation, HWIO)
  texture_utils/utils.py:8-21     build_gram:  G = F^T F / (HW + eps)  (optional mask)
  texture_utils/utils.py:23-57    build_texture_loss: mean((Ga-Gb)^2) per layer, sum coef/4, optional dloss_k/dF_k
  gatys/synthesizer.py:17-23      DEFAULT_COEFS
  gatys/synthesizer.py:59         sigmoid pre-image
  gatys/synthesizer.py:79         grad of the (batch-summed) overall loss w.r.t. the placeholder
  gatys/synthesizer.py:127-151    scipy.optimize.minimize(L-BFGS-B, bounds [0,1], maxcor, ftol=gtol=0)
  gatys/synthesizer.py:161-176    resize_crop
  gatys/synthesizer.py:179-187    initialisation distributions
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

VGG_MEAN = [103.939, 116.779, 123.68]  # BGR, [0,255]   (vgg19.py:9)
VGG_MEAN_RGB = list(reversed(VGG_MEAN))  # (vgg19.py:10)

# vgg19.py:11-30, truncated at pool4 (the Vgg19Extractor default topmost, vgg19_extractor.py:5-6)
PARAM_SHAPE = OrderedDict([
    ("conv1_1", [(3, 3, 3, 64), (64,)]),
    ("conv1_2", [(3, 3, 64, 64), (64,)]),
    ("pool1", []),
    ("conv2_1", [(3, 3, 64, 128), (128,)]),
    ("conv2_2", [(3, 3, 128, 128), (128,)]),
    ("pool2", []),
    ("conv3_1", [(3, 3, 128, 256), (256,)]),
    ("conv3_2", [(3, 3, 256, 256), (256,)]),
    ("conv3_3", [(3, 3, 256, 256), (256,)]),
    ("conv3_4", [(3, 3, 256, 256), (256,)]),
    ("pool3", []),
    ("conv4_1", [(3, 3, 256, 512), (512,)]),
    ("conv4_2", [(3, 3, 512, 512), (512,)]),
    ("conv4_3", [(3, 3, 512, 512), (512,)]),
    ("conv4_4", [(3, 3, 512, 512), (512,)]),
    ("pool4", []),
])
LAYER_NAMES = list(PARAM_SHAPE.keys())

DEFAULT_COEFS = OrderedDict([  # synthesizer.py:17-23 == po/model.py:27-33
    ("conv1_1", 1e5), ("pool1", 1e5), ("pool2", 1e5), ("pool3", 1e5), ("pool4", 1e5)])


# --------------------------------------------------------------------------------------------- weights
# Data generators (seeded synthetic weights / textures / initial images) are shared with the product package:
# they are inputs, not arithmetic under test.
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from syntex_b200.synthetic import init_input, synthetic_texture, synthetic_vgg_weights  # noqa: E402,F401


def load_vgg_weights(path_or_dict, is_rgb_input=True, topmost="pool4"):
    """Vgg19.__init__ (vgg19.py:97-130): npz/npy loading, truncation at topmost, conv1_1 channel flip."""
    if isinstance(path_or_dict, str):
        if path_or_dict.endswith(".npy"):
            data = np.load(path_or_dict, encoding="latin1", allow_pickle=True).item()
        elif path_or_dict.endswith(".npz"):
            data = np.load(path_or_dict, encoding="latin1", allow_pickle=True)
        else:
            raise ValueError("vgg19_npy_path should be a 'npy' or 'npz' file")
    else:
        data = path_or_dict
    out = OrderedDict()
    for layer in PARAM_SHAPE:
        if layer in data:
            w, b = data[layer][0], data[layer][1]
            out[layer] = [np.asarray(w, dtype=np.float32), np.asarray(b, dtype=np.float32)]
        if layer == topmost:
            break
    if is_rgb_input:
        w = out["conv1_1"][0]
        assert w.shape == PARAM_SHAPE["conv1_1"][0]
        out["conv1_1"][0] = np.ascontiguousarray(w[:, :, ::-1, :])
    return out


# --------------------------------------------------------------------------------------------- network
class TorchWeights(dict):
    """Weights pre-converted to torch OIHW tensors of one dtype (avoids re-converting 10.6 M params per call)."""

    def __init__(self, weights, dtype):
        super().__init__()
        self.dtype = dtype
        for k, (w, b) in weights.items():
            self[k] = (torch.as_tensor(np.asarray(w), dtype=dtype).permute(3, 2, 0, 1).contiguous(),
                       torch.as_tensor(np.asarray(b), dtype=dtype))


def _conv_layer(x, w, b):
    """vgg19.py:171-179. x: [B,C,H,W] torch; w HWIO numpy (or OIHW torch from TorchWeights); SAME 3x3 stride 1 =
    zero pad 1; cross-correlation (both TF conv2d and torch conv2d)."""
    if not isinstance(w, torch.Tensor):
        w = torch.as_tensor(w, dtype=x.dtype).permute(3, 2, 0, 1).contiguous()
        b = torch.as_tensor(b, dtype=x.dtype)
    y = F.conv2d(x, w, b, stride=1, padding=1)
    return F.relu(y)


def vgg_build(image_nhwc: torch.Tensor, weights, topmost="pool4", is_rgb_input=True, use_avg_pool=True):
    """Vgg19.build (vgg19.py:132-169). image [B,H,W,3] in [0,1]; returns OrderedDict layer -> [B,H,W,C]."""
    mean = VGG_MEAN_RGB if is_rgb_input else VGG_MEAN
    x = image_nhwc * 255.0 - torch.tensor(mean, dtype=image_nhwc.dtype)
    x = x.permute(0, 3, 1, 2)
    layers = OrderedDict()
    for layer in PARAM_SHAPE:
        if layer.startswith("conv"):
            x = _conv_layer(x, weights[layer][0], weights[layer][1])
        else:
            # 2x2 stride 2 SAME; identical to VALID for even sizes, the only ones any config uses
            if use_avg_pool:
                x = F.avg_pool2d(x, 2, 2, ceil_mode=True, count_include_pad=False)
            else:
                x = F.max_pool2d(x, 2, 2, ceil_mode=True)
        # the recorded NHWC tensor is the one that feeds the next layer, so gradients taken w.r.t. a recorded layer
        # (tf.gradients(loss, feat[k]), po/improved_model.py:227-231) include every path through the layers above
        feat = x.permute(0, 2, 3, 1)
        layers[layer] = feat
        x = feat.permute(0, 3, 1, 2)
        if layer == topmost:
            break
    return layers


def build_gram(feat: torch.Tensor, mask=None, eps=1e-6):
    """utils.py:8-21. feat [B,H,W,C] -> [B,C,C]."""
    shape = list(feat.shape)
    if mask is not None:
        mask = mask.unsqueeze(-1)
        feat = feat * mask
        normalizer = mask.sum(dim=(1, 2, 3), keepdim=True).reshape(-1, 1, 1)
    else:
        normalizer = float(np.float32(np.prod(shape[1:-1])))
    fr = feat.reshape(shape[0], -1, shape[-1])
    # reference adds float32(eps) to a float32 normaliser (utils.py:20): inert for HW >= 16 in fp32
    if feat.dtype == torch.float32 and mask is None:
        normalizer = float(np.float32(normalizer) + np.float32(eps))
    else:
        normalizer = normalizer + eps
    return torch.matmul(fr.transpose(1, 2), fr) / normalizer


def build_texture_loss(a, b, coefs, is_gram_a=False, is_gram_b=True, calc_grad=False, create_graph=False):
    """utils.py:23-57. Returns (loss_overall [B], loss_layer {k: [B]}, grad_layer {k: like a[k]} or None).
    create_graph=True keeps grad_layer differentiable (what TF does when a PO model trains through it)."""
    gram_a = a if is_gram_a else OrderedDict((k, build_gram(a[k])) for k in coefs)
    gram_b = b if is_gram_b else OrderedDict((k, build_gram(b[k])) for k in coefs)
    loss_layer = OrderedDict()
    for k in coefs:
        diff = gram_a[k] - gram_b[k]
        loss_layer[k] = (diff * diff).mean(dim=(1, 2))
    loss_overall = sum(coefs[k] * 1.0 / 4 * loss_layer[k] for k in loss_layer)
    if calc_grad:
        grad_layer = OrderedDict()
        for k in coefs:
            # tf.gradients(loss_layer[k], a[k]) = gradient of the batch sum
            grad_layer[k] = torch.autograd.grad(loss_layer[k].sum(), a[k], retain_graph=True,
                                                create_graph=create_graph)[0]
        return loss_overall, loss_layer, grad_layer
    return loss_overall, loss_layer, None


class OracleTexture:
    """CPU restatement of the PO models' door into the hot path (po/progressive_model.py:157-160 +
    po/model.py:51-69): same protocol as syntex_b200.po.texture_op.CudaTexture, plain torch autograd with
    create_graph=True, so that back-propagating a training loss through grad_layer exercises the second-order
    term exactly as tf.gradients does in the reference graph. Test infrastructure only."""

    def __init__(self, weights, default_layers, dtype=torch.float64):
        self.weights = weights if isinstance(weights, TorchWeights) else TorchWeights(weights, dtype)
        self.dtype = dtype
        self.default_layers = list(default_layers)

    def grams(self, image, layers=None):
        layers = self.default_layers if layers is None else list(layers)
        feat = vgg_build(image.detach().to("cpu", self.dtype), self.weights, topmost=layers[-1])
        return OrderedDict((k, build_gram(feat[k]).detach()) for k in layers)

    def __call__(self, image, layers, gram_target, slot, calc_grad):
        with torch.enable_grad():        # grad_layer is itself an autograd.grad result, also in inference
            return self._call(image, layers, gram_target, slot, calc_grad)

    def _call(self, image, layers, gram_target, slot, calc_grad):
        layers = list(layers)
        image = image.to("cpu", self.dtype)
        if not image.requires_grad:      # first stage: the input is data; autograd.grad still needs a graph
            image = image.detach().requires_grad_(True)
        feat = vgg_build(image, self.weights, topmost=layers[-1])
        coefs = OrderedDict((k, 1.0) for k in layers)
        target = OrderedDict((k, gram_target[k].to("cpu", self.dtype)) for k in layers)
        _, loss_layer, grad_layer = build_texture_loss(feat, target, coefs, calc_grad=calc_grad, create_graph=True)
        return loss_layer, grad_layer


    def prep(self, image, gram_target, coefs, stop_grad, slot, need_grads=True):
        with torch.enable_grad():
            return self._prep(image, gram_target, coefs, stop_grad, slot)

    def _prep(self, image, gram_target, coefs, stop_grad, slot):
        """ImprovedSynTex / AdaptiveSynTex.build_stage_preparation (po/improved_model.py:209-232,
        po/adaptive_model.py:204-222): grads = tf.gradients(sum_k coef_k/4 loss_k, [feat_k]), kept differentiable."""
        layers = list(coefs.keys())
        image = image.to("cpu", self.dtype)
        if not image.requires_grad:
            image = image.detach().requires_grad_(True)
        feat = vgg_build(image, self.weights, topmost=layers[-1])
        target = OrderedDict((k, gram_target[k].to("cpu", self.dtype)) for k in layers)
        _, loss_per_layer, _ = build_texture_loss(feat, target, coefs)
        loss_grad = sum(coefs[k] * 1.0 / 4 * loss_per_layer[k] for k in layers)
        grads = torch.autograd.grad(loss_grad.sum(), [feat[k] for k in layers], create_graph=not stop_grad,
                                    retain_graph=True)
        if stop_grad:
            grads = [g.detach() for g in grads]
        loss_input = sum(coefs[k] * 1.0 / 4 * loss_per_layer[k] for k in layers[:-1]) if len(layers) > 1 \
            else loss_per_layer[layers[0]] * 0.0
        return (OrderedDict((k, feat[k]) for k in layers), loss_input, loss_per_layer,
                OrderedDict(zip(layers, grads)))


class SynthesizerOracle:
    """gatys/synthesizer.py:16-159 with the TF session replaced by torch-CPU autograd."""

    def __init__(self, weights, image_size, coefs=None, dtype=torch.float64, maxcor=20):
        self.weights = weights if isinstance(weights, TorchWeights) else TorchWeights(weights, dtype)
        self.image_size = image_size
        self.coefs = OrderedDict(coefs) if coefs is not None else DEFAULT_COEFS
        self.dtype = dtype
        self.maxcor = maxcor
        self.gram_target = None
        self.shape_input = [1, image_size, image_size, 3]
        self.nfev = 0

    def _image(self, x, is_preimage):
        return torch.sigmoid(x) if is_preimage else x

    def compute_gram(self, image, update=False, is_preimage=False):
        image = np.asarray(image)
        if image.ndim == 3:
            image = image[np.newaxis]
        # float64 fed to a float32 placeholder is rounded to fp32 first (synthesizer.py:57, :115)
        x = torch.as_tensor(image.astype(np.float32)).to(self.dtype)
        with torch.no_grad():
            feat = vgg_build(self._image(x, is_preimage), self.weights)
            grams = OrderedDict((k, build_gram(feat[k])) for k in self.coefs)
        if update:
            self.gram_target = grams
        return OrderedDict((k, v.numpy()) for k, v in grams.items())

    def step(self, inp, is_preimage=False, calc_grad=False):
        inp = np.asarray(inp)
        if inp.ndim == 3:
            inp = inp[np.newaxis]
        x = torch.as_tensor(inp.astype(np.float32)).to(self.dtype).requires_grad_(True)
        feat = vgg_build(self._image(x, is_preimage), self.weights)
        loss_overall, loss_layer, grad_layer = build_texture_loss(feat, self.gram_target, self.coefs,
                                                                  calc_grad=calc_grad)
        (grad,) = torch.autograd.grad(loss_overall.sum(), x)
        self.nfev += 1
        ret = {"func": loss_overall.detach().numpy(), "grad": grad.numpy(),
               "loss_layer": OrderedDict((k, v.detach().numpy()) for k, v in loss_layer.items())}
        if calc_grad:
            ret["grad_layer"] = OrderedDict((k, v.numpy()) for k, v in grad_layer.items())
        return ret

    def synthesize(self, input_init, image_target, maxiter=100, bounds=None, is_preimage=False, callback=None):
        from scipy.optimize import minimize

        shape = list(np.asarray(input_init).shape)
        if len(shape) == 3:
            shape = [1] + shape

        def f(x):
            ret = self.step(x.reshape(shape), is_preimage=is_preimage)
            return [np.sum(ret["func"]), np.array(ret["grad"].ravel(), dtype=np.float64)]

        if bounds is None and not is_preimage:
            n = int(np.prod(shape))
            bounds = np.tile(np.array([[0.0, 1.0]]), (n, 1))   # get_bounds, synthesizer.py:10-14
        self.compute_gram(image_target, update=True)
        self.nfev = 0
        # scipy's BLAS threads spin between calls and starve torch's OpenMP pool (5x slowdown measured here):
        # L-BFGS-B's own linear algebra is tiny, so pin it to one thread.
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1, user_api="blas"):
            res = minimize(f, np.asarray(input_init, dtype=np.float64).ravel(), method="L-BFGS-B", jac=True,
                           bounds=bounds, callback=callback,
                           options={"maxiter": maxiter, "maxcor": self.maxcor, "ftol": 0.0, "gtol": 0.0})
        return res["x"].reshape(shape[1:]), res["fun"], res


def resize_crop(img, size):
    """synthesizer.py:161-176 (cv2 bilinear resize of the shortest edge + centre crop)."""
    import cv2
    assert img.ndim == 3
    h, w = img.shape[:2]
    if h < w:
        newh, neww = size, int(w * 1.0 * size / h + 0.5)
        h_off, w_off = 0, (neww - size) // 2
    else:
        newh, neww = int(h * 1.0 * size / w + 0.5), size
        h_off, w_off = (newh - size) // 2, 0
    ret = cv2.resize(img, (neww, newh), interpolation=cv2.INTER_LINEAR)
    if ret.ndim == 2:
        ret = ret[:, :, np.newaxis]
    return ret[h_off:h_off + size, w_off:w_off + size, :]
