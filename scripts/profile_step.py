"""One f/g evaluation workload for ncu: `python scripts/profile_step.py [batch] [size] [evals] [precision]`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntex_b200 import _lib  # noqa: E402
from syntex_b200.gatys import Synthesizer  # noqa: E402
from syntex_b200.synthetic import init_input, synthetic_texture, synthetic_vgg_weights  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
evals = int(sys.argv[3]) if len(sys.argv) > 3 else 3
prec = sys.argv[4] if len(sys.argv) > 4 else "bf16x3"
syn = Synthesizer({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": size, "maxiter": 1, "maxcor": 20,
                   "batch": batch, "precision": prec})
syn.build()
tg = np.stack([synthetic_texture(size, seed=100 + j) for j in range(batch)]).astype(np.float32)
x0 = np.stack([init_input((size, size, 3), False, j) for j in range(batch)]).astype(np.float32)
syn.compute_gram(tg, update=True)
xd = torch.from_numpy(x0).cuda()
g = torch.empty_like(xd)
f = torch.empty(batch, dtype=torch.float32, device="cuda")
L = _lib.lib()
for _ in range(evals):
    _lib.check(L.stx_plan_step(syn._plan.handle, xd.data_ptr(), 0, f.data_ptr(), g.data_ptr(), None, _lib.stream_ptr()))
torch.cuda.synchronize()
print("ok", f.cpu().numpy()[:2])
