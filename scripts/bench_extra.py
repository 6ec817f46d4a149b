"""Secondary measurements for the other BASELINE.json configs (one JSON line each; the headline is bench.py).

  config 1  Gatys 256^2, 1 job, 100 L-BFGS iterations            -> iterations/s (single-job latency regime)
  config 2  Gatys 512^2, 1 job, 500 iterations                   -> iterations/s, final loss vs scipy on the same f/g
  config 3  PO inference hot-path content: VGG19 forward (+ 5 Grams) at batch 64, 256^2 -> images/s
"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntex_b200 import _lib
from syntex_b200.gatys import Synthesizer
from syntex_b200.synthetic import init_input, synthetic_texture, synthetic_vgg_weights

L = _lib.lib()
raw = synthetic_vgg_weights(0)


def gatys(size, iters, batch=1, precision="bf16x3"):
    tg = np.stack([synthetic_texture(size, seed=100 + j) for j in range(batch)]).astype(np.float32)
    x0 = np.stack([init_input((size, size, 3), False, j) for j in range(batch)]).astype(np.float32)
    syn = Synthesizer({"vgg19_npy_path": raw, "image_size": size, "maxiter": iters, "maxcor": 20, "batch": batch,
                       "precision": precision})
    syn.build()
    syn.options["maxiter"] = 10
    syn.synthesize(x0, tg)
    syn.options["maxiter"] = iters
    torch.cuda.synchronize()
    t = time.perf_counter()
    x, fun = syn.synthesize(x0, tg)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    return {"workload": "gatys_lbfgs_%d" % size, "jobs": batch, "iterations": iters, "precision": precision,
            "e2e_iterations_per_s": batch * iters / dt, "seconds": dt, "final_loss": np.atleast_1d(fun).tolist(),
            "nfev": syn.last_result["nfev"].tolist()}


def vgg_forward(batch, size, reps=10):
    img = torch.rand((batch, size, size, 3), device="cuda")
    from syntex_b200.texture_utils import Vgg19Extractor
    ext = Vgg19Extractor(raw)
    plan = ext.get_plan(batch, size, size, "pool4", Synthesizer.DEFAULT_COEFS)
    st = _lib.stream_ptr()
    for _ in range(3):
        _lib.check(L.stx_plan_forward(plan.handle, img.data_ptr(), 0, 1, st))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        _lib.check(L.stx_plan_forward(plan.handle, img.data_ptr(), 0, 1, st))
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    flops = 47.20e9 * (size / 256.0) ** 2 * batch
    return {"workload": "po_vgg19_forward_5grams", "batch": batch, "size": size, "images_per_s": batch / (ms * 1e-3),
            "ms_per_batch": ms, "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
            "note": "hot-path content of PO inference (config 3): sigmoid -> VGG19 forward to pool4 + 5 Gram matrices, bf16x3"}


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c3"]
    if "c1" in which:
        print(json.dumps(gatys(256, 100)), flush=True)
        print(json.dumps(gatys(256, 100, precision="bf16")), flush=True)
    if "c2" in which:
        print(json.dumps(gatys(512, 500)), flush=True)
    if "c3" in which:
        print(json.dumps(vgg_forward(64, 256)), flush=True)
