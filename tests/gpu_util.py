"""Helpers shared by the -m gpu tests: everything goes through the C-ABI (ctypes) of the built library."""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

from syntex_b200 import _lib


def L():
    return _lib.lib()


def st():
    return _lib.stream_ptr()


def relerr(a, b):
    a = torch.as_tensor(a).double().flatten().cpu()
    b = torch.as_tensor(b).double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


def gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


def pack(w_hwio, split=False):
    cin, cout = w_hwio.shape[2], w_hwio.shape[3]
    wf = torch.empty((9, cout, cin), dtype=torch.bfloat16, device="cuda")
    wl = torch.empty_like(wf) if split else None
    wd = torch.empty((9, cin, cout), dtype=torch.bfloat16, device="cuda")
    _lib.check(L().stx_pack_conv_weights_split(w_hwio.contiguous().data_ptr(), cin, cout, wf.data_ptr(),
                                               wl.data_ptr() if split else None, wd.data_ptr(), st()))
    return wf, wl, wd


def conv_ref(x_nhwc, w_hwio, b=None):
    """Plain PyTorch fp32/fp64 reference of the same op (SAME 3x3 cross-correlation, HWIO)."""
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2), w_hwio.permute(3, 2, 0, 1), b, padding=1)
    return y.permute(0, 2, 3, 1)


def ptr(t):
    return None if t is None else t.data_ptr()
