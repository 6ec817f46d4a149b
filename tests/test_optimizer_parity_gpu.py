"""Final-loss parity of the device optimiser with the reference path on BASELINE.json's own configurations.

Reference path (gatys/synthesizer.py:127-151): scipy L-BFGS-B (bounds [0,1], maxcor 20, ftol = gtol = 0) over the
float32 f/g evaluation. It was run once in the build container over the float32 CPU oracle
(tests/golden/make_optimizer_golden.py -> tests/golden/optimizer_config{1,2}.npz); the fixture holds, per seed, the
loss at the image it returned, evaluated by the FLOAT64 oracle. Here the device optimiser
(csrc/lbfgs.cu over the CUDA f/g) runs the same budget from the same starting images, its returned images are
evaluated by the same float64 oracle on the host, and the geometric mean over seeds of loss(device) / loss(reference)
must be within BASELINE.json's 5 %. (Single seeds scatter by +-10 %: L-BFGS trajectories are chaotic - scipy against
itself under another summation order does the same - so the bar is on the mean, with a 25 % guard per seed.)"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import syntex_oracle as so  # noqa: E402
from syntex_b200.gatys import Synthesizer  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run(raw_weights, exemplar, fixture, max_seeds=None):
    if not os.path.exists(os.path.join(GOLDEN, fixture)):
        pytest.skip("fixture %s not generated (tests/golden/make_optimizer_golden.py)" % fixture)
    gold = np.load(os.path.join(GOLDEN, fixture))
    S, maxiter = int(gold["size"]), int(gold["maxiter"])
    seeds = [int(s) for s in gold["seeds"]][:max_seeds]
    tgt = so.resize_crop(exemplar / 255.0, S)
    B = len(seeds)
    syn = Synthesizer({"vgg19_npy_path": raw_weights, "image_size": S, "maxiter": maxiter, "maxcor": 20, "batch": B})
    syn.build()
    x0 = np.stack([so.init_input((S, S, 3), False, s) for s in seeds])
    x, fun = syn.synthesize(x0 if B > 1 else x0[0], np.stack([tgt] * B) if B > 1 else tgt)
    x = np.asarray(x).reshape(B, S, S, 3)
    nit = np.asarray(syn.last_result["nit"]).ravel()
    nfev = np.asarray(syn.last_result["nfev"]).ravel()
    assert x.min() >= 0.0 and x.max() <= 1.0
    o64 = so.SynthesizerOracle(so.load_vgg_weights(raw_weights), S, dtype=torch.float64)
    o64.compute_gram(tgt, update=True)
    f64 = np.array([float(o64.step(x[j])["func"].sum()) for j in range(B)])
    ref = gold["scipy_f64"][:B]
    ratios = f64 / ref
    gm = float(np.exp(np.mean(np.log(ratios))))
    print("\n%s: size %d, %d iterations, seeds %s" % (fixture, S, maxiter, seeds))
    for j in range(B):
        print("  seed %d: device %.5e (nit %d nfev %d, reported %.5e)  reference %.5e (nit %d nfev %d)  ratio %.4f" % (
            seeds[j], f64[j], nit[j], nfev[j], np.asarray(fun).ravel()[j], ref[j], gold["scipy_nit"][j],
            gold["scipy_nfev"][j], ratios[j]))
    print("  geometric mean of loss(device) / loss(reference) = %.4f" % gm)
    return gm, ratios, nit, maxiter


def test_config1_256_100_iterations_final_loss_within_5_percent(raw_weights, exemplar):
    """BASELINE.json configs[0]: 256^2, FlowerBeds0008, 100 L-BFGS-B iterations, seeds 0..7."""
    gm, ratios, nit, maxiter = _run(raw_weights, exemplar, "optimizer_config1.npz")
    assert (nit == maxiter).all(), nit
    assert 0.95 < gm < 1.05, (gm, ratios)
    assert ratios.max() < 1.25 and ratios.min() > 0.75, ratios


def test_config2_512_500_iterations_final_loss_within_5_percent(raw_weights, exemplar):
    """BASELINE.json configs[1]: 512^2, one exemplar, 500 L-BFGS-B iterations, every seed the fixture holds (a reference
    run is 10-15 minutes of CPU: 6 seeds; single seeds scatter by +-6 % and re-roll with every change of the kernels'
    rounding, so the mean over fewer seeds than that is not a stable statistic)."""
    gm, ratios, nit, maxiter = _run(raw_weights, exemplar, "optimizer_config2.npz")
    assert (nit == maxiter).all(), nit
    assert 0.95 < gm < 1.05, (gm, ratios)
    assert ratios.max() < 1.25 and ratios.min() > 0.75, ratios
