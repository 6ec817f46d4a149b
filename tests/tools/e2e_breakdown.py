"""Where the wall time of Synthesizer.synthesize() with host arrays goes (the bench's `e2e` leg) beyond the optimiser
cycles themselves: host-side phases timed one by one at the bench's configuration.

  python tests/tools/e2e_breakdown.py [jobs] [iterations]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from syntex_b200 import _lib  # noqa: E402
from syntex_b200.gatys import Synthesizer  # noqa: E402
from syntex_b200.synthetic import init_input, synthetic_texture, synthetic_vgg_weights  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
S = 256
syn = Synthesizer({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": S, "maxiter": iters, "maxcor": 20, "batch": batch})
syn.build()
tg = np.stack([synthetic_texture(S, seed=100 + j) for j in range(batch)]).astype(np.float32)
x0 = np.stack([init_input((S, S, 3), False, j) for j in range(batch)]).astype(np.float32)


def T():
    torch.cuda.synchronize()
    return time.perf_counter()


syn.synthesize(x0, tg)
for rep in range(2):
    t0 = T()
    x, fun = syn.synthesize(x0, tg)
    t1 = T()
    nfev = float(np.mean(syn.last_result["nfev"]))
    print("synthesize(): %.1f ms for %d iterations x %d jobs = %.0f job-iterations/s; %.1f evaluations per job"
          % ((t1 - t0) * 1e3, iters, batch, batch * iters / (t1 - t0), nfev), flush=True)
# phases
t0 = T()
syn._set_target(tg)
t1 = T()
x0b = syn._to_batch(x0)
t2 = T()
opt = syn._get_lbfgs(True, 0.0, 1.0)
L = _lib.lib()
st = _lib.stream_ptr()
xd = torch.from_numpy(x0b).cuda()
t3 = T()
_lib.check(L.stx_lbfgs_reset(opt, xd.data_ptr(), 0, st))
_lib.check(L.stx_lbfgs_run(opt, iters, 0, st))
t4 = T()
xo = np.empty(syn.shape_input, dtype=np.float32)
fo = np.empty((batch,), dtype=np.float32)
ni = np.empty((batch,), dtype=np.int32)
nf = np.empty((batch,), dtype=np.int32)
_lib.check(L.stx_lbfgs_minimize_host(opt, x0b.ctypes.data, 0, iters, xo.ctypes.data, fo.ctypes.data, ni.ctypes.data,
                                     nf.ctypes.data, st))
t5 = T()
x64 = xo.astype(np.float64)
t6 = T()
print("set_target (H2D exemplars + forward + Grams) %.1f ms | to_batch %.1f | H2D x0 (torch, pageable) %.1f | "
      "reset + run on device %.1f | minimize_host (H2D + run + D2H) %.1f | astype(float64) %.1f"
      % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t5 - t4) * 1e3, (t6 - t5) * 1e3))
