"""Host mirror of texture_utils/utils.py (build_gram, build_texture_loss) over the CUDA Gram kernels."""
import ctypes
from collections import OrderedDict

import numpy as np

from .. import _lib

NP_DTYPE = np.float32


def _as_bf16_nhwc(feat):
    import torch
    if not isinstance(feat, torch.Tensor):
        feat = torch.as_tensor(np.asarray(feat))
    if feat.dim() != 4:
        raise ValueError("feature map must be [batch, height, width, channels], got %s" % (tuple(feat.shape),))
    _lib.require_cuda()
    feat = feat.to(device="cuda").contiguous()
    if feat.dtype == torch.bfloat16:
        return feat
    out = torch.empty(feat.shape, dtype=torch.bfloat16, device="cuda")
    src = feat.to(torch.float32)
    _lib.check(_lib.lib().stx_f32_to_bf16(src.data_ptr(), src.numel(), out.data_ptr(), _lib.stream_ptr()))
    return out


def build_gram(feat, mask=None, eps=1e-6, name="gram"):
    """utils.py:8-21:  G_b = F_b^T F_b / (H*W + eps)  -> torch.float32 CUDA tensor [B,C,C].
    feat: [B,H,W,C] CUDA tensor (bf16 natively; fp32 is rounded to bf16, exact for features produced by
    Vgg19.build). The spatial-mask variant (utils.py:11-14) is outside the built path: no reference caller passes
    a mask."""
    import torch
    if mask is not None:
        raise NotImplementedError("syntex_b200: build_gram(mask=...) is not built (no reference caller uses it)")
    f = _as_bf16_nhwc(feat)
    B, H, W, C = f.shape
    if B == 0 or H * W == 0:
        raise ValueError("build_gram: empty feature map")
    L = _lib.lib()
    ws_bytes = L.stx_gram_workspace_bytes(B, H * W, C)
    ws = torch.empty(max(ws_bytes // 4, 1), dtype=torch.float32, device="cuda")
    out = torch.empty((B, C, C), dtype=torch.float32, device="cuda")
    _lib.check(L.stx_gram_bf16(f.data_ptr(), B, H * W, C, ctypes.c_float(eps), out.data_ptr(), ws.data_ptr(),
                               _lib.stream_ptr()))
    return out


def build_texture_loss(a, b, coefs, is_gram_a=False, is_gram_b=True,
                       calc_grad=False, name="texture_loss"):
    """utils.py:23-57. a is the input (dict of features, or of Grams when is_gram_a), b the target.
    Returns (loss_overall [B], loss_layer {k: [B]}, grad_layer {k: d loss_layer[k] / d a[k]} or None)."""
    import torch
    gram_a = a if is_gram_a else OrderedDict((k, build_gram(a[k])) for k in coefs)
    gram_b = b if is_gram_b else OrderedDict((k, build_gram(b[k])) for k in coefs)
    loss_layer = OrderedDict()
    for k in coefs:
        gram_diff = gram_a[k] - gram_b[k]
        loss_layer[k] = torch.mean(gram_diff * gram_diff, dim=(1, 2))
    loss_overall = sum(coefs[k] * 1. / 4 * loss_layer[k] for k in loss_layer)
    if not calc_grad:
        return loss_overall, loss_layer, None
    if is_gram_a:
        raise ValueError("calc_grad needs features for `a`, not Gram matrices")
    grad_layer = OrderedDict()
    for k in coefs:
        # d mean((G - T)^2) / dF = 4 / (C^2 (HW + eps)) * F (G - T): one 1x1 tensor-core GEMM per layer
        f = _as_bf16_nhwc(a[k])
        B, H, W, C = f.shape
        diff = (gram_a[k] - gram_b[k]) * (4.0 / (C * C * (float(np.float32(H * W) + np.float32(1e-6)))))
        d16 = torch.empty((B, C, C), dtype=torch.bfloat16, device="cuda")
        diff = diff.expand(B, C, C).contiguous()
        L = _lib.lib()
        _lib.check(L.stx_f32_to_bf16(diff.data_ptr(), diff.numel(), d16.data_ptr(), _lib.stream_ptr()))
        out16 = torch.empty((B, H, W, C), dtype=torch.bfloat16, device="cuda")
        # per-image weights, whole batch in one launch (the matrix is symmetric: [cout][cin] = itself)
        _lib.check(L.stx_conv1x1_per_image_bf16(f.data_ptr(), B, H, W, C, d16.data_ptr(), C, out16.data_ptr(),
                                                _lib.stream_ptr()))
        grad_layer[k] = out16.to(torch.float32)
    return loss_overall, loss_layer, grad_layer


def gram_grad_double_bwd(feat, gram_diff, upstream, eps=1e-6, name="gram_grad_double_bwd"):
    """Second-order term of the per-layer style gradient (SURVEY.md section 8 rows a-14 / b): with
    g = alpha F D, D = F^T F / n - T, alpha = 4 / (C^2 n), n = HW + eps, and U = dl/dg, returns

        dl/dF = alpha [ U D + F (F^T U + U^T F) / n ]

    feat [B,H,W,C] (bf16 or fp32 CUDA), gram_diff = G - T [B,C,C] fp32, upstream U [B,H,W,C].
    F^T U is the cross-Gram tensor-core kernel, the two products with C x C matrices are 1x1 tensor-core GEMMs."""
    import torch
    f = _as_bf16_nhwc(feat)
    u = _as_bf16_nhwc(upstream)
    B, H, W, C = f.shape
    if tuple(u.shape) != (B, H, W, C) or tuple(gram_diff.shape) != (B, C, C):
        raise ValueError("gram_grad_double_bwd: shape mismatch")
    L = _lib.lib()
    st = _lib.stream_ptr()
    n = float(np.float32(H * W) + np.float32(eps))
    alpha = 4.0 / (C * C * n)
    D = gram_diff.to(device="cuda", dtype=torch.float32)
    ws = torch.empty(max(L.stx_gram_cross_workspace_bytes(B, H * W, C) // 4, 1), dtype=torch.float32, device="cuda")
    X = torch.empty((B, C, C), dtype=torch.float32, device="cuda")
    _lib.check(L.stx_gram_cross_bf16(f.data_ptr(), u.data_ptr(), B, H * W, C, ctypes.c_float(eps), X.data_ptr(),
                                     ws.data_ptr(), st))
    S = (alpha * (X + X.transpose(1, 2))).contiguous()        # alpha (F^T U + U^T F) / n, symmetric
    aD = (alpha * D).contiguous()
    out = torch.zeros((B, H, W, C), dtype=torch.float32, device="cuda")
    for src, mat in ((u, aD), (f, S)):
        m16 = torch.empty((B, C, C), dtype=torch.bfloat16, device="cuda")
        _lib.check(L.stx_f32_to_bf16(mat.data_ptr(), mat.numel(), m16.data_ptr(), st))
        o16 = torch.empty((B, H, W, C), dtype=torch.bfloat16, device="cuda")
        # per-image weights, one launch for the batch: both matrices are symmetric, so [cout][cin] = the matrix itself
        _lib.check(L.stx_conv1x1_per_image_bf16(src.data_ptr(), B, H, W, C, m16.data_ptr(), C, o16.data_ptr(), st))
        out += o16.float()
    return out
