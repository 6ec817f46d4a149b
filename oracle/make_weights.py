"""Calibrate / write the synthetic VGG19 weights used by the tests and the benchmark (TEST INFRASTRUCTURE).

  python oracle/make_weights.py --calibrate      # recompute oracle/synth_scale.json (mean activation ~1 per layer)
  python oracle/make_weights.py --out w.npz      # write the weights in the reference npz layout
                                                 #  (texture_utils/caffe2npz.py:16-18: {layer: [w, b]}, object arrays)
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import syntex_oracle as so  # noqa: E402

SCALE_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "syntex_b200", "synth_scale.json")


def calibrate(seed=0, size=128):
    raw = so.synthetic_vgg_weights(seed, scales=None)
    img = torch.as_tensor(np.random.RandomState(1234).uniform(0, 1, size=(1, size, size, 3)))
    scales = []
    w = so.load_vgg_weights(raw)
    x = (img * 255.0 - torch.tensor(so.VGG_MEAN_RGB)).permute(0, 3, 1, 2)
    for layer in so.PARAM_SHAPE:
        if layer.startswith("conv"):
            y = so._conv_layer(x, w[layer][0], w[layer][1])
            m = float(y.mean())
            s = 1.0 / m
            scales.append([s, s])
            x = y * s
        else:
            x = torch.nn.functional.avg_pool2d(x, 2, 2)
    return scales


from syntex_b200.synthetic import save_vgg_npz as save_npz  # noqa: E402


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--calibrate", action="store_true")
    ap.add_argument("--out")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    if a.calibrate:
        sc = calibrate(a.seed)
        json.dump({"seed": a.seed, "scales": sc}, open(SCALE_PATH, "w"), indent=1)
        print("wrote", SCALE_PATH)
    if a.out:
        sc = json.load(open(SCALE_PATH))["scales"]
        save_npz(a.out, so.synthetic_vgg_weights(a.seed, sc))
        print("wrote", a.out)
