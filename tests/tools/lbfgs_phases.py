"""Where a cycle of the device optimiser spends its time: per-kernel times are in the ncu launch lists; this prints
the phases INSIDE the decision kernel (stx_lbfgs_phase_cycles, SM clocks) after a run with a full history, and the
event-timed cycle.

  python tests/tools/lbfgs_phases.py [jobs]
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from syntex_b200 import _lib  # noqa: E402
from syntex_b200.gatys import Synthesizer  # noqa: E402
from syntex_b200.synthetic import init_input, synthetic_texture, synthetic_vgg_weights  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 18
S = 256
syn = Synthesizer({"vgg19_npy_path": synthetic_vgg_weights(0), "image_size": S, "maxiter": 1000, "maxcor": 20,
                   "batch": batch})
syn.build()
tg = np.stack([synthetic_texture(S, seed=100 + j) for j in range(batch)]).astype(np.float32)
x0 = np.stack([init_input((S, S, 3), False, j) for j in range(batch)]).astype(np.float32)
syn.compute_gram(tg, update=True)
L = _lib.lib()
st = _lib.stream_ptr()
opt = syn._get_lbfgs(True, 0.0, 1.0)
xd = torch.from_numpy(x0).cuda()
_lib.check(L.stx_lbfgs_reset(opt, xd.data_ptr(), 0, st))
_lib.check(L.stx_lbfgs_run(opt, 1 << 30, 40, st))
names = ["sums", "serial", "tables+K", "eliminate", "back-subst", "write-back"]
for rep in range(3):
    _lib.check(L.stx_lbfgs_run(opt, 1 << 30, 1, st))
    buf = (ctypes.c_longlong * 16)()
    _lib.check(L.stx_lbfgs_phase_cycles(opt, buf))
    t = list(buf)
    print("decide kernel, job 0, %d jobs: " % batch +
          "  ".join("%s %d" % (n, t[i + 1] - t[i]) for i, n in enumerate(names)), " total %d cycles" % (t[6] - t[0]),
          flush=True)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
_lib.check(L.stx_lbfgs_run(opt, 1 << 30, 50, st))
ev1.record()
torch.cuda.synchronize()
print("cycle %.3f ms at %d jobs (50 cycles, full history)" % (ev0.elapsed_time(ev1) / 50, batch))
