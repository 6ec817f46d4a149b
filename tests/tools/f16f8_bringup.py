"""B200 bring-up of the f16 + f8 forward: Gram / loss / input-gradient error against the float64 oracle next to the
bf16x3 and plain-bf16 forwards, and the per-class kernel times of one f/g evaluation.

  python tests/tools/f16f8_bringup.py [batch] [sizes, comma separated]
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


def accuracy():
    raw = synthetic_vgg_weights(0)
    img = np.load(os.path.join(ROOT, "tests", "golden", "exemplar_flowerbeds0008_256.npz"))["image"]
    print("%-8s %5s %4s  %10s %10s %10s" % ("prec", "size", "seed", "gram_rel", "loss_rel", "grad_rel"))
    for S in SIZES:
        tgt = so.resize_crop(img / 255.0, S)
        orc = so.SynthesizerOracle(so.load_vgg_weights(raw), S, dtype=torch.float64)
        orc.compute_gram(tgt, update=True)
        for seed in (0, 1):
            x0 = so.init_input((S, S, 3), False, seed)
            ref = orc.step(x0)
            rg = orc.compute_gram(x0)
            for prec in ("f16f8", "bf16x3", "bf16"):
                syn = Synthesizer({"vgg19_npy_path": raw, "image_size": S, "maxiter": 1, "maxcor": 20, "precision": prec})
                syn.build()
                syn.compute_gram(tgt, update=True)
                got = syn.step(x0)
                gg = syn.compute_gram(x0)
                ge = max(float(np.linalg.norm(np.asarray(gg[k]) - rg[k]) / np.linalg.norm(rg[k])) for k in rg)
                fe = abs(float(got["func"].sum()) - float(ref["func"].sum())) / float(ref["func"].sum())
                gr = float(np.linalg.norm(got["grad"] - ref["grad"]) / np.linalg.norm(ref["grad"]))
                print("%-8s %5d %4d  %10.2e %10.2e %10.2e" % (prec, S, seed, ge, fe, gr), flush=True)
                syn.close()


def timing(batch):
    raw = synthetic_vgg_weights(0)
    L = _lib.lib()
    st = _lib.stream_ptr()
    names = ["conv_fwd", "conv_dgrad", "gram_grad", "gram", "gram_finish", "conv1_1", "pool", "other"]
    for prec in ("f16f8", "bf16x3"):
        S = 256
        syn = Synthesizer({"vgg19_npy_path": raw, "image_size": S, "maxiter": 1, "maxcor": 20, "precision": prec, "batch": batch})
        syn.build()
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
        print("%-8s batch %d:" % (prec, batch), "  ".join("%s %.3f" % (n, a) for n, a in zip(names, acc) if a > 0),
              " total %.3f ms" % acc.sum(), flush=True)
        syn.close()


SIZES = (64, 128, 256, 512)

if __name__ == "__main__":
    if len(sys.argv) > 2:
        SIZES = tuple(int(a) for a in sys.argv[2].split(","))
    accuracy()
    timing(int(sys.argv[1]) if len(sys.argv) > 1 else 18)
