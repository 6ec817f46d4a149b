"""Where a device-optimiser cycle goes at a given job batch: f/g evaluation alone (back-to-back stx_plan_step launches,
CUDA events) vs the whole cycle (f/g + L-BFGS dots/decide/update, graph replay).  python scripts/time_cycle_parts.py [batch]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntex_b200 import _lib
from syntex_b200.gatys import Synthesizer
from syntex_b200.synthetic import init_input, synthetic_texture, synthetic_vgg_weights
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 18
S = 256
L = _lib.lib()
syn = Synthesizer({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": S, "maxiter": 10, "maxcor": 20, "batch": batch})
syn.build()
tg = np.stack([synthetic_texture(S, seed=100 + j) for j in range(batch)]).astype(np.float32)
x0 = np.stack([init_input((S, S, 3), False, j) for j in range(batch)]).astype(np.float32)
syn.compute_gram(tg, update=True)
xd = torch.from_numpy(x0).cuda()
g = torch.empty_like(xd); f = torch.empty(batch, dtype=torch.float32, device="cuda")
st = _lib.stream_ptr()
def ev():
    return torch.cuda.Event(enable_timing=True)
for _ in range(5):
    _lib.check(L.stx_plan_step(syn._plan.handle, xd.data_ptr(), 0, f.data_ptr(), g.data_ptr(), None, st))
torch.cuda.synchronize()
a, b = ev(), ev(); a.record()
N = 50
for _ in range(N):
    _lib.check(L.stx_plan_step(syn._plan.handle, xd.data_ptr(), 0, f.data_ptr(), g.data_ptr(), None, st))
b.record(); torch.cuda.synchronize()
fg = a.elapsed_time(b) / N
opt = syn._get_lbfgs(True, 0.0, 1.0)
_lib.check(L.stx_lbfgs_reset(opt, xd.data_ptr(), 0, st))
_lib.check(L.stx_lbfgs_run(opt, 1 << 30, 40, st))          # history full (m = 20)
torch.cuda.synchronize()
a, b = ev(), ev(); a.record()
_lib.check(L.stx_lbfgs_run(opt, 1 << 30, 100, st))
b.record(); torch.cuda.synchronize()
cyc = a.elapsed_time(b) / 100
print("batch %d: f/g evaluation alone %.3f ms, whole cycle %.3f ms, optimiser + gaps %.3f ms" % (batch, fg, cyc, cyc - fg))
