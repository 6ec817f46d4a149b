"""oracle/lbfgsb_model.py (the numpy restatement of L-BFGS-B) pinned against scipy's L-BFGS-B - the optimiser the
reference calls (gatys/synthesizer.py:149-150) - and the DEVICE variant (what csrc/lbfgs.cu implements) measured
against it on the texture objective."""
import numpy as np
import pytest
import torch
from scipy.optimize import minimize as sp_minimize

import lbfgsb_model as lb
import syntex_oracle as so


def _bounded_quadratic(n=300, seed=0):
    rng = np.random.RandomState(seed)
    A = rng.normal(size=(n, n))
    H = A @ A.T / n + 0.05 * np.eye(n)
    b = rng.normal(size=n) * 2

    def fg(x):
        x = np.asarray(x, np.float64)
        return 0.5 * x @ H @ x - b @ x, H @ x - b
    return fg, rng.uniform(0, 1, n)


def _rosen(x):
    x = np.asarray(x, np.float64)
    f = np.sum(100 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2)
    g = np.zeros_like(x)
    g[:-1] = -400 * x[:-1] * (x[1:] - x[:-1] ** 2) - 2 * (1 - x[:-1])
    g[1:] += 200 * (x[1:] - x[:-1] ** 2)
    return f, g


@pytest.mark.parametrize("it", [1, 2, 3, 5, 10, 20, 40])
def test_same_iterates_as_scipy_on_a_bounded_quadratic(it):
    """83 % of the variables end at a bound: exercises the breakpoint walk of the generalized Cauchy point, the
    subspace minimisation, the projected step and both stopping tests (scipy stops at iteration 22)."""
    fg, x0 = _bounded_quadratic()
    r = sp_minimize(fg, x0, jac=True, method="L-BFGS-B", bounds=[(0, 1)] * x0.size,
                    options={"maxiter": it, "maxcor": 20, "ftol": 0, "gtol": 0})
    m = lb.minimize(fg, x0, it)
    assert (m.nit, m.nfev) == (r.nit, r.nfev)
    assert abs(m.f - r.fun) <= 1e-12 * abs(r.fun)
    np.testing.assert_allclose(m.x, r.x, atol=1e-12)


@pytest.mark.parametrize("it", [1, 3, 10, 30])
def test_same_iterates_as_scipy_on_a_bounded_rosenbrock(it):
    rng = np.random.RandomState(1)
    x0 = rng.uniform(0, 0.9, 50)
    r = sp_minimize(_rosen, x0, jac=True, method="L-BFGS-B", bounds=[(0, 0.8)] * 50,
                    options={"maxiter": it, "maxcor": 5, "ftol": 0, "gtol": 0})
    m = lb.minimize(_rosen, x0, it, m=5, lo=0.0, hi=0.8)
    assert (m.nit, m.nfev) == (r.nit, r.nfev)
    np.testing.assert_allclose(m.x, r.x, atol=1e-10)
    assert abs(m.f - r.fun) <= 1e-10 * abs(r.fun)


def test_same_iterates_as_scipy_unbounded():
    rng = np.random.RandomState(2)
    x0 = rng.uniform(-1, 1, 40)
    r = sp_minimize(_rosen, x0, jac=True, method="L-BFGS-B", options={"maxiter": 25, "maxcor": 7, "ftol": 0, "gtol": 0})
    m = lb.minimize(_rosen, x0, 25, m=7, bounded=False)
    assert (m.nit, m.nfev) == (r.nit, r.nfev)
    np.testing.assert_allclose(m.x, r.x, atol=1e-9)


def _texture_problem(S, raw_weights, exemplar):
    o32 = so.SynthesizerOracle(so.load_vgg_weights(raw_weights), S, dtype=torch.float32)
    tgt = so.resize_crop(exemplar / 255.0, S)
    o32.compute_gram(tgt, update=True)

    def fg(v):
        ret = o32.step(np.asarray(v).reshape(1, S, S, 3))
        return float(ret["func"].sum()), ret["grad"].ravel().astype(np.float64)
    return o32, tgt, fg


def test_first_iterations_on_the_texture_objective_track_scipy(raw_weights, exemplar):
    """The reference path: scipy over the float32 f/g. The float64 numpy restatement sees the same f/g, so the
    first iterations agree to rounding (later ones drift apart chaotically, as scipy does against itself)."""
    S, IT = 32, 6
    o32, tgt, fg = _texture_problem(S, raw_weights, exemplar)
    x0 = so.init_input((S, S, 3), False, 0)
    x, fun, r = o32.synthesize(x0, tgt, maxiter=IT)
    o32.compute_gram(tgt, update=True)
    m = lb.minimize(fg, x0, IT)
    assert (m.nit, m.nfev) == (r["nit"], r["nfev"])
    assert abs(m.f - fun) <= 1e-4 * fun
    assert np.abs(m.x - x.ravel()).max() < 1e-4


def test_device_variant_reaches_scipys_loss_on_the_texture_objective(raw_weights, exemplar):
    """The variant the CUDA optimiser implements (no breakpoint walk, float32 vectors) against scipy, 32^2,
    60 iterations, 6 seeds: geometric mean of the final-loss ratio within 10 % (per-seed ratios scatter by +-8 %:
    L-BFGS trajectories are chaotic; at 64^2 / 100 iterations / 8 seeds the mean is 1.008)."""
    S, IT = 32, 60
    o32, tgt, fg = _texture_problem(S, raw_weights, exemplar)
    ratios = []
    for seed in range(6):
        x0 = so.init_input((S, S, 3), False, seed)
        _, fun, _ = o32.synthesize(x0, tgt, maxiter=IT)
        o32.compute_gram(tgt, update=True)
        m = lb.minimize(fg, x0, IT, **lb.DEVICE)
        assert m.x.min() >= 0.0 and m.x.max() <= 1.0
        ratios.append(m.f / fun)
    gm = float(np.exp(np.mean(np.log(ratios))))
    assert 0.9 < gm < 1.1, ratios
