"""One launch of every hand-written generator kernel at two shapes of a 512x512 ProgressiveSynTex step (ncu target):
  ncu --set full --clock-control none --import-source on -k regex:"wgrad|in_|pad_rep|conv_patch" -o gpurun_out/prof_gen \
      python scripts/profile_gen_ops.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntex_b200 import _lib  # noqa: E402

L = _lib.lib()
st = _lib.stream_ptr


def bf(shape):
    return torch.randn(shape, device="cuda").to(torch.bfloat16)


for s, cin, cout in [(512, 64, 64), (64, 256, 256)]:
    x, dy = bf((1, s + 2, s + 2, cin)), bf((1, s, s, cout))
    wf, wd = bf((9, cout, cin)), bf((9, cin, cout))
    y = torch.empty((1, s, s, cout), dtype=torch.bfloat16, device="cuda")
    gp = torch.empty((1, s + 2, s + 2, cin), dtype=torch.bfloat16, device="cuda")
    dw = torch.empty((3, 3, cin, cout), dtype=torch.float32, device="cuda")
    ws = torch.empty(max(L.stx_conv_wgrad_workspace_bytes(1, s, s, cin, cout) // 4, 1), dtype=torch.float32, device="cuda")
    for _ in range(2):      # second round = warm instruction caches; profile with -s to skip the first
        _lib.check(L.stx_conv_bf16_pad(x.data_ptr(), 1, s, s, cin, wf.data_ptr(), cout, 0, None, None, 0, y.data_ptr(), 0, st()))
        _lib.check(L.stx_conv_bf16_pad(dy.data_ptr(), 1, s + 2, s + 2, cout, wd.data_ptr(), cin, 2, None, None, 0, gp.data_ptr(), 0, st()))
        _lib.check(L.stx_conv_wgrad_bf16(x.data_ptr(), dy.data_ptr(), 1, s, s, cin, cout, 0, dw.data_ptr(), ws.data_ptr(), st()))
        c = cout
        xi, o = bf((1, s, s, c)), torch.empty((1, s, s, c), dtype=torch.bfloat16, device="cuda")
        xp = torch.empty((1, s + 2, s + 2, c), dtype=torch.bfloat16, device="cuda")
        mean, rstd, s1, s2 = (torch.empty((1, c), dtype=torch.float32, device="cuda") for _ in range(4))
        gamma, beta = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        wsn = torch.empty(max(L.stx_inorm_workspace_bytes(1, s * s, c) // 4, 1), dtype=torch.float32, device="cuda")
        _lib.check(L.stx_inorm_stats_bf16(xi.data_ptr(), 1, s * s, c, ctypes.c_float(1e-5), mean.data_ptr(), rstd.data_ptr(), wsn.data_ptr(), st()))
        _lib.check(L.stx_inorm_apply_bf16(xi.data_ptr(), 1, s * s, c, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1, ctypes.c_float(0.0), None, o.data_ptr(), st()))
        _lib.check(L.stx_inorm_bwd_bf16(xi.data_ptr(), dy.data_ptr(), 1, s * s, c, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1, ctypes.c_float(0.0), s1.data_ptr(), s2.data_ptr(), o.data_ptr(), wsn.data_ptr(), st()))
        _lib.check(L.stx_pad_replicate_bf16(xi.data_ptr(), 1, s, s, c, 1, xp.data_ptr(), st()))
        _lib.check(L.stx_pad_replicate_bwd_bf16(xp.data_ptr(), xi.data_ptr(), 1, s, s, c, o.data_ptr(), st()))
        torch.cuda.synchronize()
print("ok")
