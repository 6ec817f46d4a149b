"""Which descriptor encoding addresses rows of a swizzled patch at arbitrary 128 B offsets? (GPU bring-up)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntex_b200 import _lib
L = _lib.lib()
g = torch.Generator(device="cuda").manual_seed(0)
rows = 180
P = torch.randn((rows, 64), device="cuda", generator=g).to(torch.bfloat16)
B = torch.randn((64, 64), device="cuda", generator=g).to(torch.bfloat16)
out = torch.empty((128, 64), dtype=torch.float32, device="cuda")
def expect(start, sbo):
    pitch = sbo // 128
    idx = [start + (m >> 3) * pitch + (m & 7) for m in range(128)]
    ok = max(idx) < rows
    A = P[torch.tensor([min(i, rows - 1) for i in idx], device="cuda")].float()
    return A @ B.float().t(), ok
for sbo in (1024, 1280, 2304):
    for start in (0, 1, 2, 3, 7, 8, 10, 11, 12, 21, 22):
        ref, ok = expect(start, sbo)
        if not ok: continue
        res = []
        for mode in (0, 1):
            out.zero_()
            _lib.check(L.stx_debug_patch_probe(P.data_ptr(), rows, B.data_ptr(), start, sbo, mode, out.data_ptr(), None))
            torch.cuda.synchronize()
            res.append(float((out - ref).norm() / ref.norm()))
        print("sbo %4d start_row %2d : base_offset=0 err %.2e | base_offset=(addr>>7)&7 err %.2e" % (sbo, start, res[0], res[1]), flush=True)
