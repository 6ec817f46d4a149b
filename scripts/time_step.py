"""Time plain-stream f/g evaluations (no CUDA graph): python scripts/time_step.py [batch] [size] [reps]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntex_b200 import _lib
from syntex_b200.gatys import Synthesizer
from syntex_b200.synthetic import init_input, synthetic_texture, synthetic_vgg_weights
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 9
size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
syn = Synthesizer({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": size, "maxiter": 1, "maxcor": 20, "batch": batch})
syn.build()
tg = np.stack([synthetic_texture(size, seed=100 + j) for j in range(batch)]).astype(np.float32)
x0 = np.stack([init_input((size, size, 3), False, j) for j in range(batch)]).astype(np.float32)
syn.compute_gram(tg, update=True)
xd = torch.from_numpy(x0).cuda(); g = torch.empty_like(xd); f = torch.empty(batch, dtype=torch.float32, device="cuda")
L = _lib.lib()
def step(): _lib.check(L.stx_plan_step(syn._plan.handle, xd.data_ptr(), 0, f.data_ptr(), g.data_ptr(), None, _lib.stream_ptr()))
for pdl in (1, 0, 1, 0):
    L.stx_debug_set(9, pdl)
    for _ in range(5): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): step()
    b.record(); torch.cuda.synchronize()
    print("batch %d size %d pdl=%d: %.1f us per f/g evaluation" % (batch, size, pdl, a.elapsed_time(b) / reps * 1e3))
