"""Generate the optimiser fixtures of BASELINE.json configs 1 and 2 (run in the build container, CPU only).

  python tests/golden/make_optimizer_golden.py [config1] [config2]

The reference path (gatys/synthesizer.py:127-151) is scipy L-BFGS-B (bounds [0,1], maxcor 20, ftol = gtol = 0) over
the float32 f/g evaluation; here that f/g is the float32 CPU oracle (oracle/syntex_oracle.py, SynthesizerOracle).
For every seed this script records the final loss of that path EVALUATED BY THE FLOAT64 ORACLE AT THE RETURNED
IMAGE, nit / nfev and the share of pixels at a bound. tests/test_optimizer_parity_gpu.py
compares the device optimiser against these numbers; the GPU box never runs the CPU optimisation itself.

With `with_model=True` (diagnostic, `model_*` keys): the same budget run by the numpy model of the device optimiser
(oracle/lbfgs_model.py) over the same float32 oracle f/g. Single seeds scatter by +-10 % (L-BFGS trajectories are
chaotic), hence 24 seeds for config 1: the standard error of the geometric mean is then ~1.5 %.

  config 1: 256^2, 100 iterations, exemplar FlowerBeds0008_256, seeds 0..23      -> optimizer_config1.npz
  config 2: 512^2, 500 iterations, the same exemplar through resize_crop(512), seeds 0..5 -> optimizer_config2.npz
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lbfgs_model as lm  # noqa: E402
import syntex_oracle as so  # noqa: E402


def target_image(S):
    img = np.load(os.path.join(HERE, "exemplar_flowerbeds0008_256.npz"))["image"]
    return so.resize_crop(img / 255.0, S)


def run(S, maxiter, seeds, out, with_model=True):
    raw = so.load_vgg_weights(so.synthetic_vgg_weights(0))
    o32 = so.SynthesizerOracle(raw, S, dtype=torch.float32)
    o64 = so.SynthesizerOracle(raw, S, dtype=torch.float64)
    tgt = target_image(S)
    o64.compute_gram(tgt, update=True)
    rec = {"size": S, "maxiter": maxiter, "seeds": np.asarray(seeds)}
    res = {k: [] for k in ("scipy_f64", "scipy_fun", "scipy_nit", "scipy_nfev", "scipy_at_bound",
                           "model_f64", "model_fun", "model_nit", "model_nfev", "model_curve")}
    for seed in seeds:
        x0 = so.init_input((S, S, 3), False, seed)
        t0 = time.time()
        x, fun, r = o32.synthesize(x0, tgt, maxiter=maxiter)
        f64 = float(o64.step(x)["func"].sum())
        res["scipy_f64"].append(f64)
        res["scipy_fun"].append(float(fun))
        res["scipy_nit"].append(int(r["nit"]))
        res["scipy_nfev"].append(int(r["nfev"]))
        res["scipy_at_bound"].append(float(np.mean((x <= 0.0) | (x >= 1.0))))
        print("S %d seed %d scipy: f64 %.6e fun %.6e nit %d nfev %d at-bound %.3f (%.0f s)" % (
            S, seed, f64, fun, r["nit"], r["nfev"], res["scipy_at_bound"][-1], time.time() - t0), flush=True)
        if with_model:
            t0 = time.time()
            o32.compute_gram(tgt, update=True)

            def fg(v):
                ret = o32.step(v.reshape(1, S, S, 3))
                return float(ret["func"].sum()), ret["grad"].ravel()
            mcurve = []
            m = lm.minimize(fg, x0, maxiter, callback=lambda opt: mcurve.append(opt.f))
            mf64 = float(o64.step(m.x.reshape(S, S, 3))["func"].sum())
            res["model_f64"].append(mf64)
            res["model_fun"].append(float(m.f))
            res["model_nit"].append(int(m.nit))
            res["model_nfev"].append(int(m.nfev))
            res["model_curve"].append(np.asarray(mcurve + [np.nan] * (maxiter - len(mcurve))))
            print("S %d seed %d model: f64 %.6e fun %.6e nit %d nfev %d ratio %.3f (%.0f s)" % (
                S, seed, mf64, m.f, m.nit, m.nfev, mf64 / f64, time.time() - t0), flush=True)
    for k, v in res.items():
        if v:
            rec[k] = np.asarray(v)
    np.savez_compressed(out, **rec)
    print("wrote", out)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ["config1", "config2"]
    if "config1" in which:
        run(256, 100, list(range(24)), os.path.join(HERE, "optimizer_config1.npz"), with_model=False)
    if "config2" in which:
        run(512, 500, list(range(6)), os.path.join(HERE, "optimizer_config2.npz"), with_model=False)
