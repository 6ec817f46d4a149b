"""Per-image parity of a large batch against single-image plans (GPU bring-up)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntex_b200.gatys import Synthesizer
from syntex_b200.synthetic import init_input, synthetic_texture, synthetic_vgg_weights
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
raw = synthetic_vgg_weights(0)
tg = np.stack([synthetic_texture(S, seed=100 + j) for j in range(B)]).astype(np.float32)
x0 = np.stack([init_input((S, S, 3), False, j) for j in range(B)]).astype(np.float32)
synb = Synthesizer({"vgg19_npy_path": raw, "image_size": S, "maxiter": 5, "maxcor": 20, "batch": B}); synb.build()
gb = synb.compute_gram(tg, update=True)
rb = synb.step(x0)
syn1 = Synthesizer({"vgg19_npy_path": raw, "image_size": S, "maxiter": 5, "maxcor": 20}); syn1.build()
for j in list(range(0, B, max(1, B // 8))) + [B - 1]:
    g1 = syn1.compute_gram(tg[j], update=True)
    r1 = syn1.step(x0[j])
    dg = max(float(np.abs(g1[i][0] - gb[i][j]).max() / np.abs(g1[i][0]).max()) for i in range(len(g1)))
    print("img %2d: func b %.6e single %.6e | grad rel diff %.3e | max gram rel diff %.2e" % (
        j, rb["func"][j], r1["func"][0], np.linalg.norm(rb["grad"][j] - r1["grad"][0]) / np.linalg.norm(r1["grad"][0]), dg), flush=True)
# optimiser behaviour per job
synb.options["maxiter"] = 30
x, fun = synb.synthesize(x0, tg)
print("nit", synb.last_result["nit"]); print("nfev", synb.last_result["nfev"]); print("fun", np.round(fun, 1))
