"""Stand-alone timings of the hand-written generator kernels at the shapes of a 512x512 ProgressiveSynTex step
(SURVEY.md section 8 row f-1): tensor-pipe TFLOP/s for the padded-mode convolutions and the filter gradient,
HBM GB/s for the InstanceNorm / padding passes. CUDA events, L2 flushed between iterations."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntex_b200 import _lib  # noqa: E402

L = _lib.lib()
st = _lib.stream_ptr
FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    ms = 0.0
    for _ in range(reps):
        FLUSH.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms += a.elapsed_time(b)
    return ms / reps


def bf(shape):
    return torch.randn(shape, device="cuda").to(torch.bfloat16)


def main():
    out = []
    shapes = [(512, 64, 64), (256, 64, 64), (128, 128, 128), (64, 256, 256), (32, 512, 512),
              (256, 64, 256), (128, 128, 256), (64, 256, 512), (32, 512, 1024)]
    for s, cin, cout in shapes:
        x = bf((1, s + 2, s + 2, cin))
        dy = bf((1, s, s, cout))
        wf, wd = bf((9, cout, cin)), bf((9, cin, cout))
        y = torch.empty((1, s, s, cout), dtype=torch.bfloat16, device="cuda")
        gp = torch.empty((1, s + 2, s + 2, cin), dtype=torch.bfloat16, device="cuda")
        dw = torch.empty((3, 3, cin, cout), dtype=torch.float32, device="cuda")
        ws = torch.empty(max(L.stx_conv_wgrad_workspace_bytes(1, s, s, cin, cout) // 4, 1), dtype=torch.float32, device="cuda")
        flop = 2.0 * s * s * 9 * cin * cout
        t_f = timed(lambda: _lib.check(L.stx_conv_bf16_pad(x.data_ptr(), 1, s, s, cin, wf.data_ptr(), cout, 0, None, None, 0, y.data_ptr(), 0, st())))
        t_d = timed(lambda: _lib.check(L.stx_conv_bf16_pad(dy.data_ptr(), 1, s + 2, s + 2, cout, wd.data_ptr(), cin, 2, None, None, 0, gp.data_ptr(), 0, st())))
        t_w = timed(lambda: _lib.check(L.stx_conv_wgrad_bf16(x.data_ptr(), dy.data_ptr(), 1, s, s, cin, cout, 0, dw.data_ptr(), ws.data_ptr(), st())))
        out.append({"op": "conv3x3", "size": s, "cin": cin, "cout": cout, "gflop": flop / 1e9,
                    "fwd_us": t_f * 1e3, "fwd_tflops": flop / t_f / 1e9, "dgrad_us": t_d * 1e3, "dgrad_tflops": flop / t_d / 1e9,
                    "wgrad_us": t_w * 1e3, "wgrad_tflops": flop / t_w / 1e9})
        print(json.dumps(out[-1]), flush=True)
    for s, c in [(512, 64), (256, 64), (128, 128), (64, 256), (32, 512)]:
        x, dy = bf((1, s, s, c)), bf((1, s, s, c))
        o = torch.empty_like(x)
        xp = torch.empty((1, s + 2, s + 2, c), dtype=torch.bfloat16, device="cuda")
        mean, rstd, s1, s2 = (torch.empty((1, c), dtype=torch.float32, device="cuda") for _ in range(4))
        gamma, beta = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        ws = torch.empty(max(L.stx_inorm_workspace_bytes(1, s * s, c) // 4, 1), dtype=torch.float32, device="cuda")
        nb = x.numel() * 2
        t_s = timed(lambda: _lib.check(L.stx_inorm_stats_bf16(x.data_ptr(), 1, s * s, c, ctypes.c_float(1e-5), mean.data_ptr(), rstd.data_ptr(), ws.data_ptr(), st())))
        t_a = timed(lambda: _lib.check(L.stx_inorm_apply_bf16(x.data_ptr(), 1, s * s, c, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1, ctypes.c_float(0.0), None, o.data_ptr(), st())))
        t_b = timed(lambda: _lib.check(L.stx_inorm_bwd_bf16(x.data_ptr(), dy.data_ptr(), 1, s * s, c, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1, ctypes.c_float(0.0), s1.data_ptr(), s2.data_ptr(), o.data_ptr(), ws.data_ptr(), st())))
        t_p = timed(lambda: _lib.check(L.stx_pad_replicate_bf16(x.data_ptr(), 1, s, s, c, 1, xp.data_ptr(), st())))
        t_q = timed(lambda: _lib.check(L.stx_pad_replicate_bwd_bf16(xp.data_ptr(), x.data_ptr(), 1, s, s, c, o.data_ptr(), st())))
        out.append({"op": "elementwise", "size": s, "c": c, "tensor_mb": nb / 1e6,
                    "stats_us": t_s * 1e3, "stats_gbs": nb / t_s / 1e6, "apply_us": t_a * 1e3, "apply_gbs": 2 * nb / t_a / 1e6,
                    "bwd_us": t_b * 1e3, "bwd_gbs": 5 * nb / t_b / 1e6, "pad_us": t_p * 1e3, "pad_gbs": 2 * nb / t_p / 1e6,
                    "pad_bwd_us": t_q * 1e3, "pad_bwd_gbs": 3 * nb / t_q / 1e6})
        print(json.dumps(out[-1]), flush=True)


if __name__ == "__main__":
    main()
