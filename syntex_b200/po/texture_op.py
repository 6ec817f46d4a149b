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


def _check_generation(L, plan, generation):
    """The backward of these nodes reads the plan's activations, Gram partials and scaled differences. Plans are
    cached per (shape, layers, slot) by Vgg19.get_plan, so a second forward with the same key - two graphs alive at
    once, or an evaluation sharing the slot - overwrites them: refuse instead of returning wrong gradients."""
    now = L.stx_plan_generation(plan.handle)
    if now != generation:
        raise RuntimeError("VGG plan was re-run since this graph's forward (generation %d -> %d): a plan keeps the "
                           "activations of its LAST forward only; give concurrent graphs distinct `slot`s"
                           % (generation, now))


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
                if calc_grad == "bf16":
                    # the kernels' own precision, for the hand-written generator: no fp32 round trip either way
                    out = torch.empty((B, h, w, c), dtype=torch.bfloat16, device="cuda")
                    _lib.check(L.stx_plan_layer_grad_bf16(plan.handle, k, out.data_ptr(), st))
                    grads.append(out)
                elif calc_grad:
                    out = torch.empty((B, h, w, c), dtype=torch.float32, device="cuda")
                    _lib.check(L.stx_plan_layer_grad(plan.handle, k, out.data_ptr(), st))
                    grads.append(out)
            ctx.plan = plan
            ctx.generation = L.stx_plan_generation(plan.handle)
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
            _check_generation(L, plan, ctx.generation)
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
                    if U.dtype == torch.bfloat16 and U.is_cuda:
                        u16 = U.contiguous()
                    else:
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


def _make_prep_function():
    """Autograd node for the stage preparation of po/improved_model.py:209-232 / po/adaptive_model.py:204-222:
    (image) -> (feat_k ..., loss_per_layer [K,B], grad_per_layer_k ...) with
        grad_per_layer = tf.gradients(sum_k coef_k/4 loss_k, [feat_k ...])      (full VGG back-propagation)
    and a backward for ALL outputs - including the second-order term of training through
    grad_per_layer when --stop-grad is off (the reference's default).

    Second order. With delta_l = dL/da_l, the back-propagation is delta_l = inj_l + Back_{l+1}(delta_{l+1}),
    inj_l = beta_l a_l D_l at the loss layers (beta = coef / (C^2 n)) and Back_l(d) = W_l^T * (d . mask_l) for a
    convolution (mask_l = [z_l > 0], piecewise constant) or AvgPoolGrad. An upstream U_k = dl/d delta_k therefore
    reaches the loss layers through the TRANSPOSED operators, a forward-like "tangent sweep"
        V_l = U_l + mask_l . (W_l * V_{l-1})          (bias-free masked convolutions; average pooling at the pools)
    after which l depends on the activations only through the closed forms <V_l, beta_l a_l D_l(a_l)>: the same
    F M + E injection as in VggTextureFunction with U := V_l, handed to the plan's ordinary backward program."""
    import torch
    from ..texture_utils.vgg19 import LAYER_NAMES, PARAM_SHAPE

    def tangent_sweep(plan, vgg_handle, B, loss_idx, U):
        """U: {layer index: bf16 [B,h,w,C]} -> V: same keys. Masked forward convolutions of the running tangent."""
        L = _lib.lib()
        st = _lib.stream_ptr()
        lowest, top = min(loss_idx), max(loss_idx)
        V, cur = {}, None
        conv_index = -1
        for l, name in enumerate(LAYER_NAMES):
            is_conv = bool(PARAM_SHAPE[name])
            if is_conv:
                conv_index += 1
            if l > top:
                break
            h, w, c = plan.layer_shape(l)
            if cur is not None:
                out = torch.empty((B, h, w, c), dtype=torch.bfloat16, device="cuda")
                if is_conv:
                    wf = ctypes.c_void_p()
                    _lib.check(L.stx_vgg_weights(vgg_handle, conv_index, ctypes.byref(wf), None, None))
                    mask, _ = _feature_ptrs(plan, l)
                    cin = cur.shape[-1]
                    _lib.check(L.stx_conv_bf16(cur.data_ptr(), B, h, w, cin, wf, c, 9, None, None, mask, 0,
                                               out.data_ptr(), 0, st))
                else:
                    _lib.check(L.stx_avgpool2_fwd(cur.data_ptr(), None, B, cur.shape[1], cur.shape[2], c, out.data_ptr(),
                                                  None, st))
                cur = out
            if l in U:
                cur = U[l] if cur is None else (cur.float() + U[l].float()).to(torch.bfloat16)
            if l in loss_idx and l >= lowest and cur is not None:
                V[l] = cur
        return V

    class VggStagePreparation(torch.autograd.Function):
        @staticmethod
        def forward(ctx, image, vgg19, layers, coef_values, slot, *gram_target):
            _lib.require_cuda()
            L = _lib.lib()
            st = _lib.stream_ptr()
            x = image.detach().to(device="cuda", dtype=torch.float32).contiguous()
            B, H, W, _ = x.shape
            K = len(layers)
            coefs = OrderedDict(zip(layers, coef_values))
            plan = vgg19.get_plan(B, H, W, topmost=layers[-1], coefs=coefs, export_layer_grads=True,
                                  full_features=True, slot=slot)
            targets = []
            for k, name in enumerate(layers):
                t = gram_target[k].detach().to(device="cuda", dtype=torch.float32).contiguous()
                if t.shape[0] not in (1, B):
                    raise ValueError("gram_target[%s] has batch %d, expected 1 or %d" % (name, t.shape[0], B))
                _lib.check(L.stx_plan_set_target_gram(plan.handle, k, t.data_ptr(),
                                                      1 if (t.shape[0] == B and B > 1) else 0, st))
                targets.append(t)
            func = torch.empty(B, dtype=torch.float32, device="cuda")
            grad = torch.empty_like(x)
            layer_loss = torch.empty((K, B), dtype=torch.float32, device="cuda")
            _lib.check(L.stx_plan_step(plan.handle, x.data_ptr(), 0, func.data_ptr(), grad.data_ptr(),
                                       layer_loss.data_ptr(), st))
            feats, grads, diffs = [], [], []
            for k, name in enumerate(layers):
                idx = LAYER_INDEX[name]
                h, w, c = plan.layer_shape(idx)
                f = torch.empty((B, h, w, c), dtype=torch.float32, device="cuda")
                _lib.check(L.stx_plan_copy_feature_f32(plan.handle, idx, f.data_ptr(), st))
                feats.append(f)
                g = torch.empty((B, h, w, c), dtype=torch.float32, device="cuda")
                _lib.check(L.stx_plan_copy_layer_total_grad(plan.handle, k, g.data_ptr(), st))
                grads.append(g)
                G = torch.empty((B, c, c), dtype=torch.float32, device="cuda")
                _lib.check(L.stx_plan_copy_gram(plan.handle, k, G.data_ptr(), st))
                diffs.append(G - targets[k])
            ctx.plan, ctx.vgg19 = plan, vgg19
            ctx.generation = L.stx_plan_generation(plan.handle)
            ctx.layers, ctx.coef_values = list(layers), [float(v) for v in coef_values]
            ctx.save_for_backward(x, *diffs)
            ctx.set_materialize_grads(False)
            return tuple(feats) + (layer_loss,) + tuple(grads)

        @staticmethod
        def backward(ctx, *ups):
            L = _lib.lib()
            st = _lib.stream_ptr()
            plan = ctx.plan
            _check_generation(L, plan, ctx.generation)
            K = len(ctx.layers)
            d_feat, d_loss, d_grad = ups[:K], ups[K], ups[K + 1:]
            saved = ctx.saved_tensors
            x, diffs = saved[0], saved[1:]
            B = x.shape[0]
            loss_idx = [LAYER_INDEX[n] for n in ctx.layers]
            U = {}
            for k, u in enumerate(d_grad):
                if u is not None:
                    u = u.to(device="cuda", dtype=torch.float32).contiguous()
                    u16 = torch.empty(u.shape, dtype=torch.bfloat16, device="cuda")
                    _lib.check(L.stx_f32_to_bf16(u.data_ptr(), u.numel(), u16.data_ptr(), st))
                    U[loss_idx[k]] = u16
            V = tangent_sweep(plan, ctx.vgg19._vgg_handle(), B, loss_idx, U) if U else {}
            mats, extras, keep = [], [], []
            for k, name in enumerate(ctx.layers):
                idx = loss_idx[k]
                h, w, c = plan.layer_shape(idx)
                n = _norm(h * w)
                alpha = 4.0 / (c * c * n)
                beta = ctx.coef_values[k] / (c * c * n)
                D = diffs[k]
                M = torch.zeros_like(D)
                if d_loss is not None:
                    M = M + (d_loss[k].to(torch.float32).reshape(B, 1, 1) * alpha) * D
                E = None
                if d_feat[k] is not None:
                    E = d_feat[k].to(device="cuda", dtype=torch.float32)
                if idx in V:
                    v16 = V[idx].contiguous()
                    f_hi, _ = _feature_ptrs(plan, idx)
                    ws = torch.empty(max(L.stx_gram_cross_workspace_bytes(B, h * w, c) // 4, 1), dtype=torch.float32,
                                     device="cuda")
                    X = torch.empty((B, c, c), dtype=torch.float32, device="cuda")
                    _lib.check(L.stx_gram_cross_bf16(f_hi, v16.data_ptr(), B, h * w, c, ctypes.c_float(EPS),
                                                     X.data_ptr(), ws.data_ptr(), st))
                    M = M + beta * (X + X.transpose(1, 2))
                    d16 = torch.empty((B, c, c), dtype=torch.bfloat16, device="cuda")
                    bD = (beta * D).contiguous()
                    _lib.check(L.stx_f32_to_bf16(bD.data_ptr(), bD.numel(), d16.data_ptr(), st))
                    e16 = torch.empty((B, h, w, c), dtype=torch.bfloat16, device="cuda")
                    for i in range(B):
                        _lib.check(L.stx_conv_bf16(v16[i].data_ptr(), 1, h, w, c, d16[i].data_ptr(), c, 1, None, None,
                                                   None, 0, e16[i].data_ptr(), 0, st))
                    E = e16.float() if E is None else E + e16.float()
                    keep += [v16, ws, X, d16, bD, e16]
                E_ptr = None
                if E is not None:
                    Ec = E.contiguous()
                    e_b = torch.empty(Ec.shape, dtype=torch.bfloat16, device="cuda")
                    _lib.check(L.stx_f32_to_bf16(Ec.data_ptr(), Ec.numel(), e_b.data_ptr(), st))
                    keep += [Ec, e_b]
                    E_ptr = e_b.data_ptr()
                M = M.contiguous()
                keep.append(M)
                mats.append(M.data_ptr())
                extras.append(E_ptr)
            Mp = (ctypes.c_void_p * K)(*mats)
            Ep = (ctypes.c_void_p * K)(*extras)
            grad = torch.empty_like(x)
            _lib.check(L.stx_plan_backward_custom(plan.handle, x.data_ptr(), 0, Mp, Ep, grad.data_ptr(), st))
            return (grad, None, None, None, None) + (None,) * K

    return VggStagePreparation


_FUNCTION = None
_PREP_FUNCTION = None


def vgg_stage_preparation(image, vgg19, gram_target, coefs, stop_grad=False, slot=0):
    """Differentiable mirror of ImprovedSynTex.build_stage_preparation (po/improved_model.py:209-232).

    Returns (feat {k: [B,h,w,C]} for the Gram layers, loss_input [B], loss_per_layer {k: [B]},
    grad_per_layer {k: [B,h,w,C]}); everything is differentiable with respect to `image`, including
    grad_per_layer unless stop_grad (the reference's --stop-grad) detaches it."""
    global _PREP_FUNCTION
    if _PREP_FUNCTION is None:
        _PREP_FUNCTION = _make_prep_function()
    layers = list(coefs.keys())
    K = len(layers)
    outs = _PREP_FUNCTION.apply(image, vgg19, layers, [float(coefs[k]) for k in layers], slot,
                                *[gram_target[k] for k in layers])
    feat = OrderedDict((k, outs[i]) for i, k in enumerate(layers))
    loss_per_layer = OrderedDict((k, outs[K][i]) for i, k in enumerate(layers))
    grads = outs[K + 1:]
    if stop_grad:
        grads = [g.detach() for g in grads]
    grad_per_layer = OrderedDict(zip(layers, grads))
    if K > 1:
        loss_input = sum(coefs[k] * 1. / 4 * loss_per_layer[k] for k in layers[:-1])
    else:
        loss_input = outs[K][0] * 0.
    return feat, loss_input, loss_per_layer, grad_per_layer


def build_stage_preparation(vgg19, image_input, gram_target, coefs, stop_grad=False, name="prep", slot=0):
    """ImprovedSynTex.build_stage_preparation (po/improved_model.py:209-232) with the reference's argument order
    (the VGG extractor is `self.vgg19` there) and its default: `--stop-grad` is OFF (po/improved_model.py:117), so
    the per-layer gradients stay differentiable. Thin wrapper over vgg_stage_preparation."""
    return vgg_stage_preparation(image_input, vgg19, gram_target, coefs, stop_grad=stop_grad, slot=slot)


def vgg_texture(image, vgg19, layers, gram_target, slot=0, calc_grad=True):
    """Differentiable VGG19 + Gram statistics of `image` against `gram_target`.

    image        torch tensor [B,H,W,3] in [0,1] (may require grad)
    vgg19        Vgg19Extractor
    layers       Gram layer names, shallow to deep; VGG is evaluated up to layers[-1]
    gram_target  {layer: [B or 1, C, C]}
    calc_grad    False / True (fp32 gradients, as the reference) / "bf16" (the kernels' own precision)
    slot         plan slot: every VGG evaluation that must stay alive until backward needs its own
    Returns (loss_layer OrderedDict {k: [B]}, grad_layer OrderedDict {k: [B,h,w,C]} or None); both differentiable
    with respect to `image` (utils.py:44-55 semantics: grad_layer[k] = d loss_layer[k] / d feat[k], unweighted)."""
    global _FUNCTION
    if _FUNCTION is None:
        _FUNCTION = _make_function()
    layers = list(layers)
    outs = _FUNCTION.apply(image, vgg19, layers, slot, calc_grad if calc_grad == "bf16" else bool(calc_grad),
                           *[gram_target[k] for k in layers])
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

    def prep(self, image, gram_target, coefs, stop_grad, slot, need_grads=True):
        """Stage preparation of the improved / adaptive models: features, losses and full-VGG-backprop gradients.
        need_grads=False (SingleSynTex reads the features only) under torch.no_grad() skips the VGG back-propagation:
        feed-forward synthesis then costs one VGG forward per stage instead of a forward and a backward."""
        import torch
        if not need_grads and not torch.is_grad_enabled():
            return self._features_and_losses(image, gram_target, coefs, slot)
        return vgg_stage_preparation(image, self.vgg19, gram_target, coefs, stop_grad=stop_grad, slot=slot)

    def _features_and_losses(self, image, gram_target, coefs, slot):
        import torch
        layers = list(coefs.keys())
        L = _lib.lib()
        st = _lib.stream_ptr()
        x = image.detach().to(device="cuda", dtype=torch.float32).contiguous()
        B, H, W, _ = x.shape
        plan = self.vgg19.get_plan(B, H, W, topmost=layers[-1], coefs=OrderedDict((k, float(coefs[k])) for k in layers),
                                   export_layer_grads=True, full_features=True, slot=slot)
        _lib.check(L.stx_plan_forward(plan.handle, x.data_ptr(), 0, 1, st))
        feat, loss_per_layer = OrderedDict(), OrderedDict()
        for k, name in enumerate(layers):
            idx = LAYER_INDEX[name]
            h, w, c = plan.layer_shape(idx)
            f = torch.empty((B, h, w, c), dtype=torch.float32, device="cuda")
            _lib.check(L.stx_plan_copy_feature_f32(plan.handle, idx, f.data_ptr(), st))
            feat[name] = f
            g = torch.empty((B, c, c), dtype=torch.float32, device="cuda")
            _lib.check(L.stx_plan_copy_gram(plan.handle, k, g.data_ptr(), st))
            d = g - gram_target[name].to(device="cuda", dtype=torch.float32)
            loss_per_layer[name] = torch.mean(d * d, dim=(1, 2))
        loss_input = sum(coefs[k] * 1. / 4 * loss_per_layer[k] for k in layers[:-1]) if len(layers) > 1 \
            else loss_per_layer[layers[0]] * 0.
        return feat, loss_input, loss_per_layer, None
