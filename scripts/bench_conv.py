"""Time every kernel/tile variant of the 3x3 convolution on the VGG19 layer shapes (CUDA events).
python scripts/bench_conv.py [batch] [size]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntex_b200 import _lib
L = _lib.lib()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
LAYERS = [("conv1_2", 1, 64, 64), ("conv2_1", 2, 64, 128), ("conv2_2", 2, 128, 128), ("conv3_1", 4, 128, 256),
          ("conv3_2", 4, 256, 256), ("conv4_1", 8, 256, 512), ("conv4_2", 8, 512, 512)]
st = _lib.stream_ptr()

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3   # us

g = torch.Generator(device="cuda").manual_seed(0)
print("batch %d size %d   (us per launch; TF/s = algorithmic 2MNK / time)" % (batch, size))
for name, ds, ci, co in LAYERS:
    H = size // ds
    x = torch.randn((batch, H, H, ci), device="cuda", generator=g).relu()
    xh = x.to(torch.bfloat16); xl = (x - xh.float()).to(torch.bfloat16)
    w = (torch.randn((3, 3, ci, co), device="cuda", generator=g) / np.sqrt(9 * ci)).float()
    wh = torch.empty((9, co, ci), dtype=torch.bfloat16, device="cuda"); wl = torch.empty_like(wh)
    wd = torch.empty((9, ci, co), dtype=torch.bfloat16, device="cuda")
    _lib.check(L.stx_pack_conv_weights_split(w.contiguous().data_ptr(), ci, co, wh.data_ptr(), wl.data_ptr(), wd.data_ptr(), st))
    oh = torch.empty((batch, H, H, co), dtype=torch.bfloat16, device="cuda"); ol = torch.empty_like(oh)
    gy = torch.randn((batch, H, H, co), device="cuda", generator=g).to(torch.bfloat16)
    dx = torch.empty((batch, H, H, ci), dtype=torch.bfloat16, device="cuda")
    flops = 2.0 * batch * H * H * 9 * ci * co
    line = "%-8s H%-3d %3d->%3d |" % (name, H, ci, co)
    for label, variants in (("fwd x3", [0, 64, 128]), ("fwd bf16", [0, 11128, 21128, 2128, 2256]),
                            ("dgrad", [0, 11064, 21064, 2064, 11128, 21128, 2128, 2256])):
        res = []
        for v in variants:
            bn = v % 1000
            nn = co if label != "dgrad" else ci
            if bn and nn % bn: continue
            try:
                if label == "fwd x3":
                    fn = lambda: _lib.check(L.stx_conv_bf16x3(xh.data_ptr(), xl.data_ptr(), batch, H, H, ci, wh.data_ptr(), wl.data_ptr(), co, None, 1, oh.data_ptr(), ol.data_ptr(), v, st))
                elif label == "fwd bf16":
                    fn = lambda: _lib.check(L.stx_conv_bf16(xh.data_ptr(), batch, H, H, ci, wh.data_ptr(), co, 9, None, None, None, 1, oh.data_ptr(), v, st))
                else:
                    fn = lambda: _lib.check(L.stx_conv_bf16(gy.data_ptr(), batch, H, H, co, wd.data_ptr(), ci, 9, None, None, xh.data_ptr(), 0, dx.data_ptr(), v, st))
                t = timeit(fn)
                res.append("%d:%.0fus/%.0fTF" % (v, t, flops / t / 1e6))
            except Exception as ex:
                res.append("%d:ERR" % v)
        print(line, label, " ".join(res), flush=True)
