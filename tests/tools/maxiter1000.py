"""The reference's DEFAULT budget (gatys/synthesizer.py:43-49: maxiter 1000, maxcor 20) on BASELINE config 1's problem:
how do the jobs of the device optimiser end? (Round 1's projected L-BFGS gave up on 6 of 18 jobs within 600 cycles.)

  python tests/tools/maxiter1000.py [seeds]      # B200
Prints per seed: iterations, evaluations, reported loss, float64-oracle loss at the returned image.
The same budget by scipy over the float32 CPU oracle: tests/golden/optimizer_config1_maxiter1000.npz (if present)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import syntex_oracle as so  # noqa: E402
from syntex_b200.gatys import Synthesizer  # noqa: E402
from syntex_b200.synthetic import synthetic_vgg_weights  # noqa: E402

S, IT = 256, 1000
nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 8
raw = synthetic_vgg_weights(0)
img = np.load(os.path.join(ROOT, "tests", "golden", "exemplar_flowerbeds0008_256.npz"))["image"]
tgt = so.resize_crop(img / 255.0, S)
syn = Synthesizer({"vgg19_npy_path": raw, "image_size": S, "maxiter": IT, "maxcor": 20, "batch": nseeds})
syn.build()
x0 = np.stack([so.init_input((S, S, 3), False, s) for s in range(nseeds)])
x, fun = syn.synthesize(x0, np.stack([tgt] * nseeds))
nit, nfev = syn.last_result["nit"], syn.last_result["nfev"]
o64 = so.SynthesizerOracle(so.load_vgg_weights(raw), S, dtype=torch.float64)
o64.compute_gram(tgt, update=True)
gold = None
gp = os.path.join(ROOT, "tests", "golden", "optimizer_config1_maxiter1000.npz")
if os.path.exists(gp):
    gold = np.load(gp)
for j in range(nseeds):
    f64 = float(o64.step(x[j])["func"].sum())
    line = "seed %d: device nit %4d nfev %4d reported %.5e float64-oracle %.5e" % (j, nit[j], nfev[j], fun[j], f64)
    if gold is not None and j < len(gold["scipy_f64"]):
        line += " | scipy nit %d nfev %d float64-oracle %.5e ratio %.3f" % (gold["scipy_nit"][j], gold["scipy_nfev"][j],
                                                                       gold["scipy_f64"][j], f64 / gold["scipy_f64"][j])
    print(line, flush=True)
print("jobs that ran the whole budget: %d of %d" % (int((nit == IT).sum()), nseeds))
