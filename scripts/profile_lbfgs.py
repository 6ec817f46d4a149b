"""A short device-resident L-BFGS run for ncu (kernels lbfgs_dots / lbfgs_decide / lbfgs_update inside the graph)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntex_b200.gatys import Synthesizer
from syntex_b200.synthetic import init_input, synthetic_texture, synthetic_vgg_weights
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 24
S = 256
syn = Synthesizer({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": S, "maxiter": iters, "maxcor": 20, "batch": batch})
syn.build()
tg = np.stack([synthetic_texture(S, seed=100 + j) for j in range(batch)])
x0 = np.stack([init_input((S, S, 3), False, j) for j in range(batch)])
x, fun = syn.synthesize(x0, tg)
print("ok", np.round(np.atleast_1d(fun)[:2], 2), syn.last_result["nfev"][:2])
