"""Generate the committed golden fixtures (run in the build container; the GPU box only reads the outputs).

  python tests/golden/make_golden.py

Writes
  tests/golden/exemplar_flowerbeds0008_256.npz   the reference's default exemplar (gatys/synthesizer.py:199,
                                                 images/flower_beds_256/FlowerBeds0008_256.jpg) decoded to uint8 RGB
  tests/golden/path_S{64,128,256}.npz            float64 oracle outputs for fixed seeds: Gram summaries, per-layer and
                                                 overall loss, gradient norm and strided probes
The oracle itself is unpinned by the reference (it ships no vectors); these files pin the oracle against
regressions and give the CUDA path fixed numbers to hit.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import syntex_oracle as so  # noqa: E402

REF_IMG = "/root/reference/images/flower_beds_256/FlowerBeds0008_256.jpg"


def exemplar():
    out = os.path.join(HERE, "exemplar_flowerbeds0008_256.npz")
    if os.path.exists(REF_IMG):
        import cv2
        img = cv2.imread(REF_IMG, cv2.IMREAD_COLOR)[:, :, ::-1]  # RGB, as imread(pilmode="RGB")
        np.savez_compressed(out, image=np.ascontiguousarray(img))
    return np.load(out)["image"]


def summarise(grams):
    d = {}
    for k, g in grams.items():
        g = np.asarray(g)[0]
        d["gram_trace_" + k] = np.trace(g)
        d["gram_fro_" + k] = np.linalg.norm(g)
        d["gram_corner_" + k] = g[:8, :8].copy()
        d["gram_rowsum_" + k] = g.sum(axis=1)[:64].copy()
    return d


def main():
    img = exemplar()
    W = so.load_vgg_weights(so.synthetic_vgg_weights(0))
    for S in (64, 128, 256):
        tgt = so.resize_crop(img / 255.0, S)
        x0 = so.init_input((S, S, 3), False, 0)
        orc = so.SynthesizerOracle(W, S, dtype=torch.float64)
        gt = orc.compute_gram(tgt, update=True)
        r = orc.step(x0, calc_grad=True)
        gx = orc.compute_gram(x0)
        d = {"S": S, "func": r["func"], "grad_norm": np.linalg.norm(r["grad"]),
             "grad_absmean": np.abs(r["grad"]).mean(),
             "grad_probe": r["grad"].ravel()[::max(1, r["grad"].size // 256)][:256].copy()}
        for k, v in r["loss_layer"].items():
            d["loss_" + k] = v
        for k, v in r["grad_layer"].items():
            d["gradlayer_norm_" + k] = np.linalg.norm(v)
        d.update({"t_" + k: v for k, v in summarise(gt).items()})
        d.update({"x_" + k: v for k, v in summarise(gx).items()})
        # pre-image mode (sigmoid parametrisation, synthesizer.py:59)
        p0 = so.init_input((S, S, 3), True, 0)
        rp = orc.step(p0, is_preimage=True)
        d["pre_func"] = rp["func"]
        d["pre_grad_norm"] = np.linalg.norm(rp["grad"])
        d["pre_grad_probe"] = rp["grad"].ravel()[::max(1, rp["grad"].size // 256)][:256].copy()
        np.savez_compressed(os.path.join(HERE, "path_S%d.npz" % S), **d)
        print("S", S, "func", r["func"], "pre_func", rp["func"], "|grad|", d["grad_norm"])


if __name__ == "__main__":
    main()
