"""Seeded synthetic inputs: VGG19 weights in the reference's npz layout, exemplar textures and initial images.

The reference ships neither weights (texture_utils/README.md:8-9) nor a way to fetch them offline, so tests and
benchmarks run on seeded He-normal weights rescaled per layer to a mean activation of ~1 (the property of the
"normalised" VGG19 the reference expects). A real vgg19_normalised.npz is used instead whenever a path is given.
"""
import json
import os
from collections import OrderedDict

import numpy as np

_CONV_SHAPES = OrderedDict([
    ("conv1_1", (3, 3, 3, 64)), ("conv1_2", (3, 3, 64, 64)),
    ("conv2_1", (3, 3, 64, 128)), ("conv2_2", (3, 3, 128, 128)),
    ("conv3_1", (3, 3, 128, 256)), ("conv3_2", (3, 3, 256, 256)), ("conv3_3", (3, 3, 256, 256)),
    ("conv3_4", (3, 3, 256, 256)),
    ("conv4_1", (3, 3, 256, 512)), ("conv4_2", (3, 3, 512, 512)), ("conv4_3", (3, 3, 512, 512)),
    ("conv4_4", (3, 3, 512, 512)),
])


def _default_scales(seed):
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "synth_scale.json")
    d = json.load(open(path))
    if d["seed"] != seed:
        raise ValueError("synth_scale.json was calibrated for seed %d" % d["seed"])
    return d["scales"]


def synthetic_vgg_weights(seed=0, scales="default"):
    """{layer: [w (3,3,Cin,Cout) float32, b (Cout,) float32]} as stored on disk by texture_utils/caffe2npz.py:16-18,
    i.e. BEFORE the RGB flip of vgg19.py:124-127. w ~ N(0, 2/(9 Cin)) * scale, b ~ N(0, 0.05) * scale."""
    if isinstance(scales, str):
        scales = _default_scales(seed)
    rng = np.random.RandomState(seed)
    out = OrderedDict()
    for i, (name, wshape) in enumerate(_CONV_SHAPES.items()):
        cin = wshape[2]
        w = rng.normal(0.0, np.sqrt(2.0 / (9 * cin)), size=wshape)
        b = rng.normal(0.0, 0.05, size=(wshape[3],))
        if scales is not None:
            w = w * scales[i][0]
            b = b * scales[i][1]
        out[name] = [w.astype(np.float32), b.astype(np.float32)]
    return out


def save_vgg_npz(path, weights):
    """Write weights the way caffe2npz.py does: one object array [w, b] per layer."""
    arrs = {}
    for k, (w, b) in weights.items():
        o = np.empty(2, dtype=object)
        o[0], o[1] = w, b
        arrs[k] = o
    np.savez(path, **arrs)


def init_input(shape, is_preimage, seed):
    """gatys/synthesizer.py:179-182 with a fixed seed (the reference draws from the global numpy state)."""
    rng = np.random.RandomState(seed)
    if is_preimage:
        return rng.uniform(-1.0, 1.0, size=shape)
    return rng.uniform(1.0 / 255, 1.0 - 1.0 / 255, size=shape)


def synthetic_texture(size, seed=0, channels=3):
    """Seeded band-limited noise texture in [0,1]: a stand-in exemplar where the reference JPEGs cannot be read."""
    rng = np.random.RandomState(seed)
    acc = np.zeros((size, size, channels))
    for octave, amp in ((4, 1.0), (8, 0.7), (16, 0.5), (32, 0.35), (64, 0.25)):
        if octave > size:
            break
        coarse = rng.normal(size=(octave, octave, channels))
        reps = size // octave
        up = np.kron(coarse, np.ones((reps, reps, 1)))
        k = max(reps // 2, 1)
        for ax in (0, 1):
            up = sum(np.roll(up, s, axis=ax) for s in range(-k, k + 1)) / (2 * k + 1)
        acc += amp * up
    acc = (acc - acc.min()) / (acc.max() - acc.min() + 1e-12)
    return acc
