"""The PO models' differentiable door into the hot path.

In the reference the PO graphs call `_build_extractor` + `build_texture_loss(..., calc_grad=True)` on the current
image (po/progressive_model.py:157-160), feed the per-layer gradients `grad_layer[k] = d loss_layer[k] / d feat[k]`
to the generator, and TensorFlow differentiates through all of it when training: the loss terms need VGG back-prop
to the image, and the `grad_layer` inputs need the *second-order* term (a double backward through utils.py:50-55).

Here that is one autograd node over the CUDA plan. With n = HW + eps, G = F^T F / n, D = G - T, alpha = 4 / (C^2 n):

    loss_layer[k] = mean(D^2)                          grad_layer[k] = alpha * F D
    upstream: w = d l / d loss_layer[k] (per image),   U = d l / d grad_layer[k]
    d l / d F  =  F * [ w * alpha * D  +  alpha / n * (F^T U + U^T F) ]  +  alpha * U D
               =  F * M + E

M (C x C per image) goes into the plan's fused Gram-gradient slot, E is added by the dgrad epilogue that produces
dL/dF_k, and the sum is back-propagated through VGG19 by the plan's backward program (stx_plan_backward_custom).
F^T U is the cross-Gram tensor-core kernel (stx_gram_cross_bf16); U D is a 1x1 tensor-core GEMM per image.
"""
import ctypes
from collections import OrderedDict

import numpy as np

from .. import _lib
from ..texture_utils.vgg19 import LAYER_INDEX

EPS = 1e-6


def _norm(hw):
    # fp32 arithmetic as in the reference: float32(HW) + float32(eps)   (utils.py:15, :20)
    return float(np.float32(hw) + np.float32(EPS))


def _feature_ptrs(plan, layer_idx):
    hi, lo = ctypes.c_void_p(), ctypes.c_void_p()
    _lib.check(_lib.lib().stx_plan_feature_bf16(plan.handle, layer_idx, ctypes.byref(hi), ctypes.byref(lo)))
    return hi.value, lo.value


def _make_function():
    import torch

    class VggTextureFunction(torch.autograd.Function):
        """(image, *gram_target) -> (loss_layer [K,B], grad_layer_0, ..., grad_layer_{K-1})."""

        @staticmethod
        def forward(ctx, image, vgg19, layers, slot, calc_grad, *gram_target):
            _lib.require_cuda()
            L = _lib.lib()
            st = _lib.stream_ptr()
            x = image.detach().to(device="cuda", dtype=torch.float32).contiguous()
            B, H, W, _ = x.shape
            coefs = OrderedDict((k, 1.0) for k in layers)
            plan = vgg19.get_plan(B, H, W, topmost=layers[-1], coefs=coefs, slot=slot)
            targets = []
            for k, name in enumerate(layers):
                t = gram_target[k].detach().to(device="cuda", dtype=torch.float32).contiguous()
                if t.shape[0] not in (1, B):
                    raise ValueError("gram_target[%s] has batch %d, expected 1 or %d" % (name, t.shape[0], B))
                _lib.check(L.stx_plan_set_target_gram(plan.handle, k, t.data_ptr(), 1 if (t.shape[0] == B and B > 1) else 0,
                                                      st))
                targets.append(t)
            _lib.check(L.stx_plan_forward(plan.handle, x.data_ptr(), 0, 1, st))
            diffs, losses, grads = [], [], []
            for k, name in enumerate(layers):
                h, w, c = plan.layer_shape(LAYER_INDEX[name])
                g = torch.empty((B, c, c), dtype=torch.float32, device="cuda")
                _lib.check(L.stx_plan_copy_gram(plan.handle, k, g.data_ptr(), st))
                d = g - targets[k]
                diffs.append(d)
                losses.append(torch.mean(d * d, dim=(1, 2)))
                if calc_grad:
                    out = torch.empty((B, h, w, c), dtype=torch.float32, device="cuda")
                    _lib.check(L.stx_plan_layer_grad(plan.handle, k, out.data_ptr(), st))
                    grads.append(out)
            ctx.plan = plan
            ctx.layers = list(layers)
            ctx.calc_grad = calc_grad
            ctx.save_for_backward(x, *diffs)
            ctx.set_materialize_grads(False)
            return (torch.stack(losses, 0),) + tuple(grads)

        @staticmethod
        def backward(ctx, d_loss, *d_grads):
            L = _lib.lib()
            st = _lib.stream_ptr()
            plan = ctx.plan
            saved = ctx.saved_tensors
            x, diffs = saved[0], saved[1:]
            B = x.shape[0]
            K = len(ctx.layers)
            mats, extras, keep = [], [], []
            for k, name in enumerate(ctx.layers):
                idx = LAYER_INDEX[name]
                h, w, c = plan.layer_shape(idx)
                n = _norm(h * w)
                alpha = 4.0 / (c * c * n)
                D = diffs[k]
                if d_loss is not None:
                    M = (d_loss[k].to(torch.float32).reshape(B, 1, 1) * alpha) * D
                else:
                    M = torch.zeros_like(D)
                U = d_grads[k] if (ctx.calc_grad and k < len(d_grads)) else None
                E_ptr = None
                if U is not None:
                    U = U.to(device="cuda", dtype=torch.float32).contiguous()
                    u16 = torch.empty((B, h, w, c), dtype=torch.bfloat16, device="cuda")
                    _lib.check(L.stx_f32_to_bf16(U.data_ptr(), U.numel(), u16.data_ptr(), st))
                    f_hi, _ = _feature_ptrs(plan, idx)
                    # X = F^T U / n  (tensor cores, split-K over the pixels)
                    ws = torch.empty(max(L.stx_gram_cross_workspace_bytes(B, h * w, c) // 4, 1), dtype=torch.float32,
                                     device="cuda")
                    X = torch.empty((B, c, c), dtype=torch.float32, device="cuda")
                    _lib.check(L.stx_gram_cross_bf16(f_hi, u16.data_ptr(), B, h * w, c, ctypes.c_float(EPS),
                                                     X.data_ptr(), ws.data_ptr(), st))
                    M = M + alpha * (X + X.transpose(1, 2))
                    # E = alpha * U D: per-image 1x1 GEMM, weights [cout][cin] = alpha * D (symmetric)
                    d16 = torch.empty((B, c, c), dtype=torch.bfloat16, device="cuda")
                    aD = (alpha * D).contiguous()
                    _lib.check(L.stx_f32_to_bf16(aD.data_ptr(), aD.numel(), d16.data_ptr(), st))
                    e16 = torch.empty((B, h, w, c), dtype=torch.bfloat16, device="cuda")
                    for i in range(B):
                        _lib.check(L.stx_conv_bf16(u16[i].data_ptr(), 1, h, w, c, d16[i].data_ptr(), c, 1, None, None,
                                                   None, 0, e16[i].data_ptr(), 0, st))
                    E_ptr = e16.data_ptr()
                    keep += [u16, ws, X, d16, aD, e16]
                M = M.contiguous()
                keep.append(M)
                mats.append(M.data_ptr())
                extras.append(E_ptr)
            Mp = (ctypes.c_void_p * K)(*mats)
            Ep = (ctypes.c_void_p * K)(*extras)
            grad = torch.empty_like(x)
            _lib.check(L.stx_plan_backward_custom(plan.handle, x.data_ptr(), 0, Mp, Ep, grad.data_ptr(), st))
            return (grad, None, None, None, None) + (None,) * K

    return VggTextureFunction


_FUNCTION = None


def vgg_texture(image, vgg19, layers, gram_target, slot=0, calc_grad=True):
    """Differentiable VGG19 + Gram statistics of `image` against `gram_target`.

    image        torch tensor [B,H,W,3] in [0,1] (may require grad)
    vgg19        Vgg19Extractor
    layers       Gram layer names, shallow to deep; VGG is evaluated up to layers[-1]
    gram_target  {layer: [B or 1, C, C]}
    slot         plan slot: every VGG evaluation that must stay alive until backward needs its own
    Returns (loss_layer OrderedDict {k: [B]}, grad_layer OrderedDict {k: [B,h,w,C]} or None); both differentiable
    with respect to `image` (utils.py:44-55 semantics: grad_layer[k] = d loss_layer[k] / d feat[k], unweighted)."""
    global _FUNCTION
    if _FUNCTION is None:
        _FUNCTION = _make_function()
    layers = list(layers)
    outs = _FUNCTION.apply(image, vgg19, layers, slot, bool(calc_grad), *[gram_target[k] for k in layers])
    loss_layer = OrderedDict((k, outs[0][i]) for i, k in enumerate(layers))
    grad_layer = OrderedDict((k, outs[1 + i]) for i, k in enumerate(layers)) if calc_grad else None
    return loss_layer, grad_layer


class CudaTexture(object):
    """The hot path as the PO models see it: `grams(image)` for the targets (po/model.py:60-69) and
    `__call__(image, layers, gram_target, slot, calc_grad)` for a differentiable stage evaluation."""

    def __init__(self, vgg19, default_layers):
        self.vgg19 = vgg19
        self.default_layers = list(default_layers)

    def grams(self, image, layers=None):
        import torch
        layers = self.default_layers if layers is None else list(layers)
        L = _lib.lib()
        st = _lib.stream_ptr()
        x = image.detach().to(device="cuda", dtype=torch.float32).contiguous()
        B, H, W, _ = x.shape
        plan = self.vgg19.get_plan(B, H, W, topmost=layers[-1], coefs=OrderedDict((k, 1.0) for k in layers), slot=-1)
        _lib.check(L.stx_plan_forward(plan.handle, x.data_ptr(), 0, 1, st))
        out = OrderedDict()
        for k, name in enumerate(layers):
            _, _, c = plan.layer_shape(LAYER_INDEX[name])
            g = torch.empty((B, c, c), dtype=torch.float32, device="cuda")
            _lib.check(L.stx_plan_copy_gram(plan.handle, k, g.data_ptr(), st))
            out[name] = g
        return out

    def __call__(self, image, layers, gram_target, slot, calc_grad):
        return vgg_texture(image, self.vgg19, layers, gram_target, slot=slot, calc_grad=calc_grad)
