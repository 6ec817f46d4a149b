"""B200 check of the CTA-pair (tcgen05 cta_group::2) instances of the halo-patch convolution: one f/g evaluation with
stx_debug_set key 15 = 0 / 1 (forward) / 2 (input gradient) / 3 (both, the default) - results must be bit-identical (the pair's MMA
accumulates the same products in the same order), per-class kernel times next to each other.

  python tests/tools/pair_check.py [batch] [size]
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import syntex_oracle as so  # noqa: E402
from syntex_b200 import _lib  # noqa: E402
from syntex_b200.gatys import Synthesizer  # noqa: E402
from syntex_b200.synthetic import synthetic_vgg_weights  # noqa: E402

NAMES = ["conv_fwd", "conv_dgrad", "gram_grad", "gram", "gram_finish", "conv1_1", "pool", "other"]


def run(batch, S, mode, prec):
    raw = synthetic_vgg_weights(0)
    L = _lib.lib()
    st = _lib.stream_ptr()
    _lib.check(L.stx_debug_set(15, mode))
    try:
        syn = Synthesizer({"vgg19_npy_path": raw, "image_size": S, "maxiter": 1, "maxcor": 20, "precision": prec, "batch": batch})
        syn.build()
    finally:
        _lib.check(L.stx_debug_set(15, 3))
    tg = np.stack([so.synthetic_texture(S, seed=100 + j) for j in range(batch)]).astype(np.float32)
    x0 = np.stack([so.init_input((S, S, 3), False, j) for j in range(batch)]).astype(np.float32)
    syn.compute_gram(tg, update=True)
    xd = torch.from_numpy(x0).cuda()
    g = torch.empty_like(xd)
    f = torch.empty(batch, dtype=torch.float32, device="cuda")
    ms = (ctypes.c_float * 8)()
    cnt = (ctypes.c_int * 8)()
    acc = np.zeros(8)
    for i in range(13):
        _lib.check(L.stx_plan_step_timed(syn._plan.handle, xd.data_ptr(), 0, f.data_ptr(), g.data_ptr(), ms, cnt, st))
        if i >= 3:
            acc += np.array(list(ms))
    acc /= 10
    torch.cuda.synchronize()
    print("%-6s pair=%d batch %d size %d:" % (prec, mode, batch, S),
          "  ".join("%s %.3f" % (n, a) for n, a in zip(NAMES, acc) if a > 0), " total %.3f ms" % acc.sum(), flush=True)
    out = (f.cpu().numpy().copy(), g.cpu().numpy().copy())
    syn.close()
    return out


if __name__ == "__main__":
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    for prec in ("f16f8", "bf16"):
        base = run(batch, S, 0, prec)
        for mode in (1, 2, 3):
            got = run(batch, S, mode, prec)
            same = np.array_equal(base[0], got[0]) and np.array_equal(base[1], got[1])
            print("   identical to pair=0:", same, "" if same else
                  " max |df|/f %.3e  |dg| rel %.3e" % (np.max(np.abs(base[0] - got[0]) / base[0]),
                                                       np.linalg.norm(base[1] - got[1]) / np.linalg.norm(base[1])), flush=True)
