"""Host mirror of the PO models' stage preparation on the hot path.

po/improved_model.py:209-232 (`build_stage_preparation`) and po/adaptive_model.py:204-222 feed their generators the
VGG features of the current image and the gradients of the texture loss with respect to those features,
`tf.gradients(loss_input + coef/4 * loss_last, [feat[k] for k in layers])`, i.e. a full back-propagation through
VGG19 from every Gram layer above. Here that is one CUDA plan step whose backward keeps dL/dF_k for every loss
layer (STX_PLAN_EXPORT_LAYER_GRADS). The generator networks themselves are outside this package (row f-1).
"""
from collections import OrderedDict

import numpy as np

from .. import _lib
from ..texture_utils.vgg19 import LAYER_INDEX, PARAM_SHAPE


def build_stage_preparation(vgg19, image_input, gram_target, coefs, stop_grad=True, name="prep"):
    """Mirror of ImprovedSynTex.build_stage_preparation (po/improved_model.py:209-232).

    vgg19        Vgg19Extractor
    image_input  torch CUDA tensor [B,H,W,3] in [0,1]
    gram_target  {layer: tensor [B or 1, C, C]}
    coefs        OrderedDict layer -> weight; VGG is evaluated up to the last key only (improved_model.py:213)
    Returns (feat_input, loss_input [B], loss_per_layer {k: [B]}, grad_per_layer {k: [B,h,w,c]}).
    The gradients are constants w.r.t. autograd, as with the reference's default `--stop-grad` (improved_model.py:229);
    differentiating through them (double backward) is not built."""
    import torch
    if not stop_grad:
        raise NotImplementedError("syntex_b200: second-order terms through the VGG gradients are not built "
                                  "(use stop_grad=True, the reference's --stop-grad)")
    layers = list(coefs.keys())
    x = image_input.to(device="cuda", dtype=torch.float32).contiguous()
    B, H, W, _ = x.shape
    plan = vgg19.get_plan(B, H, W, topmost=layers[-1], coefs=coefs, export_layer_grads=True, full_features=True)
    L = _lib.lib()
    st = _lib.stream_ptr()
    for k, name_k in enumerate(layers):
        g = gram_target[name_k].to(device="cuda", dtype=torch.float32).contiguous()
        per_image = 1 if g.shape[0] == B and B > 1 else 0
        if g.shape[0] not in (1, B):
            raise ValueError("gram_target[%s] has batch %d, expected 1 or %d" % (name_k, g.shape[0], B))
        _lib.check(L.stx_plan_set_target_gram(plan.handle, k, g.data_ptr(), per_image, st))
    func = torch.empty(B, dtype=torch.float32, device="cuda")
    grad = torch.empty_like(x)
    layer_loss = torch.empty((len(layers), B), dtype=torch.float32, device="cuda")
    _lib.check(L.stx_plan_step(plan.handle, x.data_ptr(), 0, func.data_ptr(), grad.data_ptr(), layer_loss.data_ptr(),
                               st))
    feat_input = OrderedDict()
    for layer in PARAM_SHAPE:
        idx = LAYER_INDEX[layer]
        h, w, c = plan.layer_shape(idx)
        out = torch.empty((B, h, w, c), dtype=torch.float32, device="cuda")
        _lib.check(L.stx_plan_copy_feature_f32(plan.handle, idx, out.data_ptr(), st))
        feat_input[layer] = out
        if layer == layers[-1]:
            break
    loss_per_layer = OrderedDict((name_k, layer_loss[k]) for k, name_k in enumerate(layers))
    if len(layers) > 1:
        loss_input = sum(coefs[k] * 1. / 4 * loss_per_layer[k] for k in layers[:-1])
    else:
        loss_input = torch.zeros(B, dtype=torch.float32, device="cuda")
    grad_per_layer = OrderedDict()
    for k, name_k in enumerate(layers):
        h, w, c = plan.layer_shape(LAYER_INDEX[name_k])
        out = torch.empty((B, h, w, c), dtype=torch.float32, device="cuda")
        _lib.check(L.stx_plan_copy_layer_total_grad(plan.handle, k, out.data_ptr(), st))
        grad_per_layer[name_k] = out
    return feat_input, loss_input, loss_per_layer, grad_per_layer
