"""CPU emulation of candidate forward precisions for the VGG19 f/g evaluation (no GPU needed).

The float64 oracle is re-run with the OPERANDS of every 3x3 convolution rounded the way a tensor-core scheme would
see them (straight-through in the backward pass, so the gradient that comes out is the one a kernel with these
forward roundings, exact accumulation and an exact backward would compute: ReLU masks and Gram differences from
the rounded forward). Reports Gram / loss / input-gradient error against the unrounded float64 oracle.

  python tests/tools/precision_emulation.py [sizes ...]

Schemes (MMAs per algorithmic MAC at kind::f16 rate in brackets):
  bf16            A, W rounded to bf16                                            [1]
  bf16x3          A, W as bf16 hi+lo pairs, lo*lo dropped                          [3]   (round-1 default)
  fp16            A, W rounded to fp16                                            [1]
  tf32            A, W rounded to 10 explicit mantissa bits                         [2: kind::tf32 runs at half rate]
  fp16_a2         A as fp16 hi+lo, W fp16                                           [2]
  fp16_w2         A fp16, W as fp16 hi+lo                                           [2]
  fp16x3          A, W as fp16 hi+lo pairs, lo*lo dropped                          [3]
  fp16_fp8corr    fp16(A)*fp16(W) + [e4m3(A)*e4m3(W_lo) + e4m3(A_lo)*e4m3(W)]       [2: the correction runs as ONE
                  kind::f8f6f4 MMA over the concatenated K at twice the f16 rate]
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import syntex_oracle as so  # noqa: E402


def r_bf16(x):
    return x.to(torch.float32).to(torch.bfloat16).to(torch.float64)


def r_fp16(x):
    return x.to(torch.float32).to(torch.float16).to(torch.float64)


def r_tf32(x):
    v = x.to(torch.float32).contiguous().view(torch.int32)
    v = (v + 0x1000) & ~0x1FFF            # round to nearest, 10 explicit mantissa bits (ties away)
    return v.view(torch.float32).to(torch.float64)


def r_e4m3(x, scale):
    """e4m3 of x * scale (saturating), returned unscaled."""
    y = (x * scale).to(torch.float32).clamp(-448.0, 448.0).to(torch.float8_e4m3fn).to(torch.float64)
    return y / scale


def pow2_scale(t, target):
    m = float(t.abs().max())
    if m == 0.0:
        return 1.0
    return 2.0 ** np.floor(np.log2(target / m))


def ste(x, xq):
    return x + (xq - x).detach()


class Scheme:
    def __init__(self, name):
        self.name = name

    def conv(self, a, w, b):
        n = self.name
        if n == "exact":
            return F.conv2d(a, w, b, padding=1)
        if n in ("bf16", "fp16", "tf32"):
            r = {"bf16": r_bf16, "fp16": r_fp16, "tf32": r_tf32}[n]
            return F.conv2d(ste(a, r(a)), r(w), b, padding=1)
        if n in ("bf16x3", "fp16x3"):
            r = r_bf16 if n == "bf16x3" else r_fp16
            ah, wh = r(a), r(w)
            al, wl = r(a - ah), r(w - wh)
            y = F.conv2d(ste(a, ah + al), wh + wl, b, padding=1)
            return y - F.conv2d(al, wl, None, padding=1).detach()
        if n == "fp16_a2":
            ah = r_fp16(a)
            al = r_fp16(a - ah)
            return F.conv2d(ste(a, ah + al), r_fp16(w), b, padding=1)
        if n == "fp16_w2":
            wh = r_fp16(w)
            wl = r_fp16(w - wh)
            return F.conv2d(ste(a, r_fp16(a)), wh + wl, b, padding=1)
        if n == "fp16_fp8corr":
            ah, wh = r_fp16(a), r_fp16(w)
            al, wl = a - ah, w - wh
            # power-of-two scales into e4m3's range: activations are data dependent, so a FIXED scale per plane kind
            # (a kernel cannot look at the maximum first): A * 2^-3 (max 3584), A_lo * 2^9; weights per layer.
            a8 = r_e4m3(a, 2.0 ** -3)
            al8 = r_e4m3(al, 2.0 ** 9)
            w8 = r_e4m3(w, pow2_scale(w, 224.0))
            wl8 = r_e4m3(wl, pow2_scale(wl, 224.0))
            y = F.conv2d(ste(a, ah), wh, b, padding=1)
            corr = F.conv2d(a8, wl8, None, padding=1) + F.conv2d(al8, w8, None, padding=1)
            return y + corr.detach()
        raise ValueError(n)


def run(S, scheme, raw, tgt, x0):
    W = so.TorchWeights(raw, torch.float64)
    sch = Scheme(scheme)

    def feats(img):
        x = img * 255.0 - torch.tensor(so.VGG_MEAN_RGB, dtype=torch.float64)
        x = x.permute(0, 3, 1, 2)
        out = {}
        for layer in so.PARAM_SHAPE:
            if layer.startswith("conv"):
                w, b = W[layer]
                # conv1_1 runs on CUDA cores in fp32 from the fp32 image in every scheme
                x = F.relu(F.conv2d(x, w, b, padding=1) if layer == "conv1_1" else sch.conv(x, w, b))
            else:
                x = F.avg_pool2d(x, 2, 2)
            out[layer] = x.permute(0, 2, 3, 1)
        return out

    with torch.no_grad():
        ft = feats(torch.as_tensor(tgt[None]))
        target = {k: so.build_gram(ft[k]) for k in so.DEFAULT_COEFS}
    x = torch.as_tensor(x0[None]).requires_grad_(True)
    fx = feats(x)
    loss, _, _ = so.build_texture_loss(fx, target, so.DEFAULT_COEFS)
    (g,) = torch.autograd.grad(loss.sum(), x)
    grams = {k: so.build_gram(fx[k]).detach() for k in so.DEFAULT_COEFS}
    return float(loss.sum()), g.numpy(), grams


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [64, 128]
    raw = so.load_vgg_weights(so.synthetic_vgg_weights(0))
    img = np.load(os.path.join(ROOT, "tests", "golden", "exemplar_flowerbeds0008_256.npz"))["image"]
    schemes = ["bf16", "bf16x3", "fp16", "tf32", "fp16_a2", "fp16_w2", "fp16x3", "fp16_fp8corr"]
    print("%-14s %5s %4s  %10s %10s %10s" % ("scheme", "size", "seed", "gram_rel", "loss_rel", "grad_rel"))
    for S in sizes:
        tgt = so.resize_crop(img / 255.0, S)
        for seed in (0, 1):
            x0 = so.init_input((S, S, 3), False, seed)
            f0, g0, G0 = run(S, "exact", raw, tgt, x0)
            for sc in schemes:
                f, g, G = run(S, sc, raw, tgt, x0)
                ge = max(float((G[k] - G0[k]).norm() / G0[k].norm()) for k in G0)
                print("%-14s %5d %4d  %10.2e %10.2e %10.2e" % (sc, S, seed, ge, abs(f - f0) / f0,
                                                               np.linalg.norm(g - g0) / np.linalg.norm(g0)), flush=True)


if __name__ == "__main__":
    torch.set_num_threads(4)
    main()
