"""Run one convolution variant a few times (for ncu): python scripts/one_conv.py layer variant [batch] [mode]
layer: conv1_2 conv2_1 conv2_2 conv3_1 conv3_2 conv4_1 conv4_2; mode: dgrad | fwd3"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntex_b200 import _lib
L = _lib.lib()
LAYERS = {"conv1_2": (1, 64, 64), "conv2_1": (2, 64, 128), "conv2_2": (2, 128, 128), "conv3_1": (4, 128, 256),
          "conv3_2": (4, 256, 256), "conv4_1": (8, 256, 512), "conv4_2": (8, 512, 512)}
name, variant = sys.argv[1], int(sys.argv[2])
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 9
mode = sys.argv[4] if len(sys.argv) > 4 else "dgrad"
ds, ci, co = LAYERS[name]
H = 256 // ds
st = _lib.stream_ptr()
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((batch, H, H, ci), device="cuda", generator=g).relu()
xh = x.to(torch.bfloat16); xl = (x - xh.float()).to(torch.bfloat16)
w = (torch.randn((3, 3, ci, co), device="cuda", generator=g) / np.sqrt(9 * ci)).float()
wh = torch.empty((9, co, ci), dtype=torch.bfloat16, device="cuda"); wl = torch.empty_like(wh)
wd = torch.empty((9, ci, co), dtype=torch.bfloat16, device="cuda")
_lib.check(L.stx_pack_conv_weights_split(w.contiguous().data_ptr(), ci, co, wh.data_ptr(), wl.data_ptr(), wd.data_ptr(), st))
oh = torch.empty((batch, H, H, co), dtype=torch.bfloat16, device="cuda"); ol = torch.empty_like(oh)
gy = torch.randn((batch, H, H, co), device="cuda", generator=g).to(torch.bfloat16)
dx = torch.empty((batch, H, H, ci), dtype=torch.bfloat16, device="cuda")
for _ in range(4):
    if mode == "dgrad":
        _lib.check(L.stx_conv_bf16(gy.data_ptr(), batch, H, H, co, wd.data_ptr(), ci, 9, None, None, xh.data_ptr(), 0, dx.data_ptr(), variant, st))
    else:
        _lib.check(L.stx_conv_bf16x3(xh.data_ptr(), xl.data_ptr(), batch, H, H, ci, wh.data_ptr(), wl.data_ptr(), co, None, 1, oh.data_ptr(), ol.data_ptr(), variant, st))
torch.cuda.synchronize()
print("ok")
