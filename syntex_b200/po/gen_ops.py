"""Hand-written generator operators of the PO models as autograd nodes over the C ABI.

The reference builds its generators from tensorpack Conv2D / InstanceNorm and tf.pad(SYMMETRIC)
(po/progressive_model.py:77-206) and lets tf.gradients differentiate them. Here the trainable 3x3 layers run on the
same tcgen05 implicit-GEMM kernel as VGG19 (forward: VALID convolution of an explicitly replicated-edge input; input
gradient: FULL correlation with the rotated filter; filter gradient: split-K pixel contraction, csrc/wgrad.cu), the
normalisation / ReLU / padding passes on HBM-bound kernels (csrc/gen_ops.cu). Activations are NHWC bf16, parameters
stay fp32 `nn.Parameter`s in torch layout (OIHW), so the stock-torch route of po/layers.py and this one share weights
and checkpoints and can be compared against each other (tests/test_po_native_gpu.py).

There is no CPU route: every function raises without the CUDA library or a CUDA device.
"""
import ctypes

import torch

from .. import _lib

_PHASE = None


def _phase_matrices(device):
    """Nearest-neighbour x2 up-sampling followed by a 5x5 convolution (build_upsampling_nn,
    po/progressive_model.py:123-134) is, per output parity, a 3x3 convolution of the LOW-resolution map: along one axis
    the five taps of an even output pixel 2j read x[j-1], x[j-1], x[j], x[j], x[j+1], those of an odd pixel 2j+1 read
    x[j-1], x[j], x[j], x[j+1], x[j+1]. P[parity][a, t] = 1 where tap t lands on low-resolution neighbour a. The
    symmetric 2-pixel padding of the up-sampled map is exactly edge replication of the low-resolution one."""
    global _PHASE
    if _PHASE is None or _PHASE.device != device:
        even = [[1, 1, 0, 0, 0], [0, 0, 1, 1, 0], [0, 0, 0, 0, 1]]
        odd = [[1, 0, 0, 0, 0], [0, 1, 1, 0, 0], [0, 0, 0, 1, 1]]
        _PHASE = torch.tensor([even, odd], dtype=torch.float32, device=device)
    return _PHASE


def upsample_weights(w):
    """[Cout, Cin, k, k] (k = 5: build_upsampling_nn, po/progressive_model.py:123-134; k = 3:
    build_upsampling_nnconv, po/improved_model.py:185-199) -> [4 * Cout, Cin, 3, 3]: output channel
    (py * 2 + px) * Cout + o holds the filter that produces up-sampled pixel (2 j + py, 2 i + px) from the
    low-resolution neighbourhood of (j, i). Differentiable."""
    k = w.shape[-1]
    if k == 5:
        P = _phase_matrices(w.device).to(w.dtype)
    elif k == 3:
        # three taps on the up-sampled map: even pixel 2j reads x[j-1], x[j], x[j]; odd pixel 2j+1 reads x[j], x[j], x[j+1]
        P = torch.tensor([[[1, 0, 0], [0, 1, 1], [0, 0, 0]], [[0, 0, 0], [1, 1, 0], [0, 0, 1]]], dtype=w.dtype,
                         device=w.device)
    else:
        raise NotImplementedError("upsample_weights: kernel size %d" % k)
    w3 = torch.einsum("yat,xbu,oitu->yxoiab", P, P, w)
    return w3.reshape(4 * w.shape[0], w.shape[1], 3, 3)


def pixel_shuffle_nhwc(y, cout):
    """[B, h, w, 4 * cout] with channel (py * 2 + px) * cout + o -> [B, 2 h, 2 w, cout]."""
    B, h, w, _ = y.shape
    return y.view(B, h, w, 2, 2, cout).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * h, 2 * w, cout)


def _bf16(shape):
    return torch.empty(shape, dtype=torch.bfloat16, device="cuda")


def _f32(shape):
    return torch.empty(shape, dtype=torch.float32, device="cuda")


_PAD_MODES = {"rep": 0, "symmetric": 0, "reflect": 1}


class Conv3x3Function(torch.autograd.Function):
    """y = conv3x3(pad(act?(x)), W) + b on NHWC bf16. mode: "rep" / "symmetric" = tf.pad(SYMMETRIC) by one pixel +
    VALID (progressive_model.py:171-175), "reflect" = tf.pad(REFLECT) + VALID (improved_model.py:49-55), "zero" =
    Conv2D(padding="SAME") (single_model.py:170-171). relu_in: False / None = no activation, True = ReLU, a float =
    LeakyReLU with that slope, applied to x on the way into the padded copy."""

    @staticmethod
    def forward(ctx, x, weight, bias, mode, relu_in):
        _lib.require_cuda()
        L = _lib.lib()
        st = _lib.stream_ptr()
        if mode not in _PAD_MODES and mode != "zero":
            raise ValueError("Conv3x3Function: unknown padding mode %r" % (mode,))
        act_in = relu_in is not None and relu_in is not False
        slope = 0.0 if relu_in is True or not act_in else float(relu_in)
        if act_in and mode == "zero":
            raise ValueError("Conv3x3Function: the fused input activation needs a padded copy (not the zero mode)")
        x = x.contiguous()
        if x.dtype != torch.bfloat16 or not x.is_cuda:
            raise ValueError("Conv3x3Function expects a CUDA bf16 NHWC tensor")
        B, H, W, cin = x.shape
        cout = weight.shape[0]
        if tuple(weight.shape[1:]) != (cin, 3, 3):
            raise ValueError("Conv3x3Function: weight %s does not match %d input channels" % (tuple(weight.shape), cin))
        coutp = (cout + 63) // 64 * 64          # convlast (3 colours) runs as a 64-channel tile, extra filters zero
        w_oihw = weight.detach().to(torch.float32).contiguous()
        w_fwd, w_dgrad = _bf16((9, coutp, cin)), _bf16((9, cin, coutp))
        _lib.check(L.stx_pack_conv_weights_oihw(w_oihw.data_ptr(), cin, cout, coutp, w_fwd.data_ptr(), w_dgrad.data_ptr(),
                                                st))
        b = None
        if bias is not None:
            b = bias.detach().to(torch.float32)
            if coutp != cout:
                b = torch.nn.functional.pad(b, (0, coutp - cout))
            b = b.contiguous()
            if b.data_ptr() % 16:           # the epilogue reads the bias as float4; a flat parameter buffer need not align
                b = b.clone()
        if mode != "zero":
            xin = _bf16((B, H + 2, W + 2, cin))
            _lib.check(L.stx_pad2d_bf16(x.data_ptr(), B, H, W, cin, _PAD_MODES[mode], 1 if act_in else 0,
                                        ctypes.c_float(slope), xin.data_ptr(), st))
            pad = 0
        else:
            xin, pad = x, 1
        y = _bf16((B, H, W, coutp))
        _lib.check(L.stx_conv_bf16_pad(xin.data_ptr(), B, H, W, cin, w_fwd.data_ptr(), coutp, pad,
                                       b.data_ptr() if b is not None else None, None, 0, y.data_ptr(), 0, st))
        ctx.mode, ctx.slope, ctx.cout, ctx.has_bias = mode, slope, cout, bias is not None
        ctx.save_for_backward(xin, w_dgrad, x if act_in else None)
        return y if coutp == cout else y[..., :cout].contiguous()

    @staticmethod
    def backward(ctx, dy):
        L = _lib.lib()
        st = _lib.stream_ptr()
        xin, w_dgrad, x_mask = ctx.saved_tensors
        cin, coutp = w_dgrad.shape[1], w_dgrad.shape[2]
        cout = ctx.cout
        B, H, W = dy.shape[0], dy.shape[1], dy.shape[2]
        dy = dy.to(torch.bfloat16)
        if coutp != cout:
            dy = torch.nn.functional.pad(dy, (0, coutp - cout))
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _bf16((B, H, W, cin))
            if ctx.mode != "zero":
                # gradient w.r.t. the padded input = FULL correlation with the rotated filter; the ring then folds back
                gp = _bf16((B, H + 2, W + 2, cin))
                _lib.check(L.stx_conv_bf16_pad(dy.data_ptr(), B, H + 2, W + 2, coutp, w_dgrad.data_ptr(), cin, 2, None,
                                               None, 0, gp.data_ptr(), 0, st))
                _lib.check(L.stx_pad2d_bwd_bf16(gp.data_ptr(), x_mask.data_ptr() if x_mask is not None else None,
                                                B, H, W, cin, _PAD_MODES[ctx.mode], ctypes.c_float(ctx.slope),
                                                dx.data_ptr(), st))
            else:
                _lib.check(L.stx_conv_bf16_pad(dy.data_ptr(), B, H, W, coutp, w_dgrad.data_ptr(), cin, 1, None, None, 0,
                                               dx.data_ptr(), 0, st))
        if ctx.needs_input_grad[1]:
            ws = torch.empty(max(L.stx_conv_wgrad_workspace_bytes(B, H, W, cin, coutp) // 4, 1), dtype=torch.float32,
                             device="cuda")
            dw = _f32((cout, cin, 3, 3))          # torch's layout: adds straight into the parameter's .grad
            _lib.check(L.stx_conv_wgrad_oihw_bf16(xin.data_ptr(), dy.data_ptr(), B, H, W, cin, coutp,
                                                  1 if ctx.mode == "zero" else 0, cout, dw.data_ptr(), ws.data_ptr(), st))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy[..., :cout].to(torch.float32).sum(dim=(0, 1, 2))
        return dx, dw, db, None, None


class InstanceNormFunction(torch.autograd.Function):
    """out = act?(InstanceNorm(x) * gamma + beta) (+ shortcut): tensorpack InstanceNorm over H x W per image and
    channel, with the activation after it (relu: False / None = none, True = ReLU, a float = LeakyReLU slope) and the
    residual add of build_res_block (progressive_model.py:77-99) fused."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, relu, shortcut):
        act = relu is not None and relu is not False
        slope = 0.0 if relu is True or not act else float(relu)
        _lib.require_cuda()
        L = _lib.lib()
        st = _lib.stream_ptr()
        x = x.contiguous()
        if x.dtype != torch.bfloat16 or not x.is_cuda:
            raise ValueError("InstanceNormFunction expects a CUDA bf16 NHWC tensor")
        B, H, W, C = x.shape
        g = gamma.detach().to(torch.float32).contiguous()
        b = beta.detach().to(torch.float32).contiguous()
        mean, rstd = _f32((B, C)), _f32((B, C))
        ws = torch.empty(max(L.stx_inorm_workspace_bytes(B, H * W, C) // 4, 1), dtype=torch.float32, device="cuda")
        _lib.check(L.stx_inorm_stats_bf16(x.data_ptr(), B, H * W, C, ctypes.c_float(eps), mean.data_ptr(),
                                          rstd.data_ptr(), ws.data_ptr(), st))
        sc = None
        if shortcut is not None:
            sc = shortcut.contiguous()
            if sc.shape != x.shape or sc.dtype != torch.bfloat16:
                raise ValueError("InstanceNormFunction: shortcut must match the input")
        out = _bf16((B, H, W, C))
        _lib.check(L.stx_inorm_apply_bf16(x.data_ptr(), B, H * W, C, mean.data_ptr(), rstd.data_ptr(), g.data_ptr(),
                                          b.data_ptr(), 1 if act else 0, ctypes.c_float(slope),
                                          sc.data_ptr() if sc is not None else None, out.data_ptr(), st))
        ctx.relu, ctx.slope, ctx.has_shortcut = act, slope, shortcut is not None
        ctx.save_for_backward(x, mean, rstd, g, b)
        return out

    @staticmethod
    def backward(ctx, dy):
        L = _lib.lib()
        st = _lib.stream_ptr()
        x, mean, rstd, g, b = ctx.saved_tensors
        B, H, W, C = x.shape
        dy = dy.to(torch.bfloat16).contiguous()
        s1, s2 = _f32((B, C)), _f32((B, C))
        dx = _bf16((B, H, W, C))
        ws = torch.empty(max(L.stx_inorm_workspace_bytes(B, H * W, C) // 4, 1), dtype=torch.float32, device="cuda")
        _lib.check(L.stx_inorm_bwd_bf16(x.data_ptr(), dy.data_ptr(), B, H * W, C, mean.data_ptr(), rstd.data_ptr(),
                                        g.data_ptr(), b.data_ptr(), 1 if ctx.relu else 0, ctypes.c_float(ctx.slope),
                                        s1.data_ptr(), s2.data_ptr(), dx.data_ptr(), ws.data_ptr(), st))
        # per-image sums -> parameter gradients (at the reference's per-GPU batch of 1 there is nothing to add up)
        dgamma = (s2[0] if B == 1 else s2.sum(0)) if ctx.needs_input_grad[1] else None
        dbeta = (s1[0] if B == 1 else s1.sum(0)) if ctx.needs_input_grad[2] else None
        return dx, dgamma, dbeta, None, None, (dy if ctx.has_shortcut else None)


def conv3x3(x, weight, bias=None, mode="rep", relu_in=False):
    return Conv3x3Function.apply(x, weight, bias, mode, relu_in)


def instance_norm(x, gamma, beta, eps=1e-5, relu=False, shortcut=None):
    return InstanceNormFunction.apply(x, gamma, beta, eps, relu, shortcut)
