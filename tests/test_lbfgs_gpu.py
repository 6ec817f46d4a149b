"""The device-resident L-BFGS against its numpy model and against the reference's optimiser call (scipy
L-BFGS-B), all driven by the same GPU f/g evaluation."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import lbfgs_model as lm  # noqa: E402
import syntex_oracle as so  # noqa: E402
from syntex_b200.gatys import Synthesizer  # noqa: E402
from syntex_b200.gatys import test_single as run_single  # noqa: E402  (not a pytest test)


def make(raw, S, it, **kw):
    a = {"vgg19_npy_path": raw, "image_size": S, "maxiter": it, "maxcor": 20}
    a.update(kw)
    syn = Synthesizer(a)
    syn.build()
    return syn


def test_device_optimizer_runs_the_budget_and_respects_bounds(raw_weights, exemplar):
    S, IT = 64, 40
    tgt = so.resize_crop(exemplar / 255.0, S)
    x0 = so.init_input((S, S, 3), False, 0)
    syn = make(raw_weights, S, IT)
    x, fun = syn.synthesize(x0, tgt)
    assert x.shape == (S, S, 3) and x.min() >= 0.0 and x.max() <= 1.0
    assert int(syn.last_result["nit"][0]) == IT
    assert IT <= int(syn.last_result["nfev"][0]) <= 3 * IT
    f0 = syn.step(x0)["func"][0]
    assert fun < 0.01 * f0
    assert abs(syn.step(x)["func"][0] - fun) <= 1e-5 * fun         # reported loss is the loss at the returned image
    x2, fun2 = syn.synthesize(x0, tgt)                               # deterministic end to end
    assert fun2 == fun and np.array_equal(x, x2)


def test_device_optimizer_tracks_the_numpy_model(raw_weights):
    """Same algorithm, same f/g: the first iterations must agree closely (later ones drift apart chaotically:
    the two differ in summation order, and L-BFGS amplifies that by roughly a digit per few iterations)."""
    S, IT = 64, 3
    tgt, x0 = so.synthetic_texture(S, seed=3), so.init_input((S, S, 3), False, 0)
    syn = make(raw_weights, S, IT)
    x, fun = syn.synthesize(x0, tgt)

    def fg(v):
        r = syn.step(v.reshape(1, S, S, 3))
        return float(r["func"].sum()), r["grad"].ravel()
    model = lm.minimize(fg, x0, IT)
    assert model.nit == IT
    assert abs(model.f - fun) / fun < 1e-2
    moved = np.linalg.norm(x - x0)
    assert np.linalg.norm(model.x.reshape(S, S, 3) - x) < 5e-2 * moved
    assert model.nfev == int(syn.last_result["nfev"][0])


def test_final_loss_comparable_to_scipy_lbfgsb(raw_weights, exemplar):
    """The device optimiser against scipy L-BFGS-B driven by the SAME GPU f/g (optimizer="scipy"), small and quick: the
    geometric mean over four seeds of loss(device) / loss(scipy) within 12 % of 1 (single seeds scatter by +-5..10 %:
    L-BFGS trajectories are chaotic). The 5 % bar of BASELINE.json on its own configurations, against scipy over the
    CPU oracle and with enough seeds for the mean to be a stable statistic, is tests/test_optimizer_parity_gpu.py."""
    S, IT = 64, 60
    tgt = so.resize_crop(exemplar / 255.0, S)
    dev = make(raw_weights, S, IT)
    ref = make(raw_weights, S, IT, optimizer="scipy")
    ratios = []
    for seed in range(4):
        x0 = so.init_input((S, S, 3), False, seed)
        _, f_dev = dev.synthesize(x0, tgt)
        _, f_ref = ref.synthesize(x0, tgt)
        ratios.append(f_dev / f_ref)
    gm = float(np.exp(np.mean(np.log(ratios))))
    assert 0.88 < gm < 1.12, ratios
    assert max(ratios) < 1.4, ratios


@pytest.mark.parametrize("maxcor", [1, 3, 32])
def test_history_sizes_from_one_pair_to_the_largest(raw_weights, maxcor):
    """The reduced system is 2k x 2k for k = min(iterations, maxcor) pairs: one pair, a short wrapping history and the
    largest the kernels are sized for (32) all run their budget and make progress."""
    S, IT = 64, 40
    tgt = so.synthetic_texture(S, seed=6)
    x0 = so.init_input((S, S, 3), False, 2)
    syn = make(raw_weights, S, IT, maxcor=maxcor)
    x, fun = syn.synthesize(x0, tgt)
    assert int(syn.last_result["nit"][0]) == IT
    assert x.min() >= 0.0 and x.max() <= 1.0
    f0 = syn.step(x0)["func"][0]
    assert fun < (0.2 if maxcor == 1 else 0.05) * f0
    if maxcor == 32:        # more memory than iterations: same trajectory as any history that never wraps
        syn20 = make(raw_weights, S, 15, maxcor=20)
        syn32 = make(raw_weights, S, 15, maxcor=32)
        xa, fa = syn20.synthesize(x0, tgt)
        xb, fb = syn32.synthesize(x0, tgt)
        assert fa == fb and np.array_equal(xa, xb)


def test_preimage_mode_is_unbounded(raw_weights):
    S, IT = 64, 20
    tgt = so.synthetic_texture(S, seed=4)
    syn = make(raw_weights, S, IT)
    img, fun = run_single(syn, tgt, True, seed=0)
    assert img.min() > 0.0 and img.max() < 1.0 and np.isfinite(fun)
    assert int(syn.last_result["nit"][0]) == IT


def test_batched_jobs_are_independent_problems(raw_weights):
    S, IT, B = 64, 15, 3
    tg = np.stack([so.synthetic_texture(S, seed=20 + j) for j in range(B)])
    x0 = np.stack([so.init_input((S, S, 3), False, j) for j in range(B)])
    synb = make(raw_weights, S, IT, batch=B)
    xb, fb = synb.synthesize(x0, tg)
    assert xb.shape == (B, S, S, 3) and fb.shape == (B,)
    syn1 = make(raw_weights, S, IT)
    for j in range(B):
        x1, f1 = syn1.synthesize(x0[j], tg[j])
        assert f1 == fb[j]
        np.testing.assert_array_equal(x1, xb[j])


def test_scipy_optimizer_path_matches_reference_call(raw_weights):
    S, IT = 64, 10
    tgt, x0 = so.synthetic_texture(S, seed=3), so.init_input((S, S, 3), False, 0)
    syn = make(raw_weights, S, IT, optimizer="scipy")
    x, fun = syn.synthesize(x0, tgt)
    assert x.shape == (S, S, 3) and x.min() >= 0 and x.max() <= 1
    assert int(syn.last_result["nit"][0]) == IT


def test_launch_modes_give_identical_results(raw_weights):
    """CUDA-graph replay (default), plain stream launches, and both with / without programmatic dependent launch
    (stx_debug_set keys 9 and 10) only change how the same kernels are enqueued: bit-identical trajectories."""
    from syntex_b200 import _lib
    L = _lib.lib()
    S, IT = 64, 12
    tgt = so.synthetic_texture(S, seed=5)
    x0 = so.init_input((S, S, 3), False, 1)
    results = []
    try:
        for pdl, no_graph in ((1, 0), (0, 0), (1, 1), (0, 1)):
            L.stx_debug_set(9, pdl)
            L.stx_debug_set(10, no_graph)
            syn = make(raw_weights, S, IT)
            x, fun = syn.synthesize(x0, tgt)
            results.append((x, fun, int(syn.last_result["nfev"][0])))
    finally:
        L.stx_debug_set(9, 1)
        L.stx_debug_set(10, 0)
    for x, fun, nfev in results[1:]:
        assert fun == results[0][1] and nfev == results[0][2] and np.array_equal(x, results[0][0])
