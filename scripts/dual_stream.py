"""Experiment: G independent optimiser instances (each `batch` jobs, own plan, own CUDA stream, own host thread) on
one GPU. python scripts/dual_stream.py [groups] [batch] [steps]"""
import os, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntex_b200 import _lib
from syntex_b200.gatys import Synthesizer
from syntex_b200.synthetic import init_input, synthetic_texture, synthetic_vgg_weights
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 9
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
size = 256
L = _lib.lib()
raw = synthetic_vgg_weights(0)
insts = []
for g in range(G):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        syn = Synthesizer({"vgg19_npy_path": raw, "image_size": size, "maxiter": steps, "maxcor": 20, "batch": batch})
        syn.build()
        tg = np.stack([synthetic_texture(size, seed=100 + g * batch + j) for j in range(batch)]).astype(np.float32)
        x0 = np.stack([init_input((size, size, 3), False, g * batch + j) for j in range(batch)]).astype(np.float32)
        syn.compute_gram(tg, update=True)
        xd = torch.from_numpy(x0).cuda()
        opt = syn._get_lbfgs(True, 0.0, 1.0)
        _lib.check(L.stx_lbfgs_reset(opt, xd.data_ptr(), 0, _lib.stream_ptr()))
        _lib.check(L.stx_lbfgs_run(opt, 10, 0, _lib.stream_ptr()))
    insts.append((st, syn, opt, xd))
torch.cuda.synchronize()

def worker(g, upto):
    st, syn, opt, xd = insts[g]
    with torch.cuda.stream(st):
        _lib.check(L.stx_lbfgs_run(opt, upto, 0, _lib.stream_ptr()))
        st.synchronize()

for rep, upto in enumerate((10 + steps, 10 + 2 * steps)):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ths = [threading.Thread(target=worker, args=(g, upto)) for g in range(G)]
    for t in ths: t.start()
    for t in ths: t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("groups %d x batch %d: %d steps in %.1f ms -> %.0f job-iterations/s" % (G, batch, steps, dt * 1e3, G * batch * steps / dt))
