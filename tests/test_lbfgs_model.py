"""The numpy model of the device optimiser (oracle/lbfgs_model.py over oracle/lbfgsb_model.py): line-search
conditions, agreement with scipy L-BFGS-B on problems whose solution is unique, bound handling."""
import numpy as np
import pytest
from scipy.optimize import minimize as sp_minimize

import lbfgs_model as lm


def _search(phi, dphi, stp0):
    f0, g0 = phi(0.0), dphi(0.0)
    st = lm.ls_init(f0, g0, stp0)
    for _ in range(40):
        res = lm.ls_update(st, phi(st["stp"]), dphi(st["stp"]))
        if res != "again":
            return res, st, f0, g0
    return "again", st, f0, g0


@pytest.mark.parametrize("stp0", [1e-3, 0.1, 1.0, 10.0, 1e3])
def test_more_thuente_strong_wolfe(stp0):
    phi = lambda a: -a / (a * a + 2.0)                       # More' & Thuente test function 1 (beta = 2)
    dphi = lambda a: (a * a - 2.0) / (a * a + 2.0) ** 2
    res, st, f0, g0 = _search(phi, dphi, stp0)
    assert res == "accept"
    a = st["stp"]
    assert phi(a) <= f0 + lm.FTOL * a * g0
    assert abs(dphi(a)) <= lm.GTOL * abs(g0)


def test_line_search_quadratic_accepts_unit_step():
    phi = lambda a: (a - 1.0) ** 2
    dphi = lambda a: 2 * (a - 1.0)
    res, st, _, _ = _search(phi, dphi, 1.0)
    assert res == "accept" and st["nev"] == 1 and st["stp"] == 1.0


def _quadratic(n, seed, cond=50.0):
    rng = np.random.RandomState(seed)
    Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    A = Q @ np.diag(np.linspace(1.0, cond, n)) @ Q.T
    xs = rng.uniform(-0.5, 1.5, size=n)           # unconstrained minimiser partly outside [0,1]

    def fg(x):
        x = np.asarray(x, np.float64)
        r = x - xs
        return 0.5 * float(r @ A @ r), (A @ r)
    return fg


def test_bounded_quadratic_matches_scipy():
    n = 60
    fg = _quadratic(n, 0)
    x0 = np.full(n, 0.5)
    ref = sp_minimize(lambda x: fg(x), x0, jac=True, method="L-BFGS-B", bounds=[(0, 1)] * n,
                      options={"maxiter": 500, "ftol": 0, "gtol": 1e-10})
    opt = lm.minimize(lambda x: fg(x), x0, maxiter=500, m=20)
    assert np.all(opt.x >= 0) and np.all(opt.x <= 1)
    assert opt.f <= ref.fun * (1 + 1e-4) + 1e-6
    np.testing.assert_allclose(opt.x, ref.x, atol=2e-3)


def test_unbounded_matches_scipy_iterates_closely():
    n = 40
    fg = _quadratic(n, 1, cond=20.0)
    x0 = np.zeros(n)
    ref = sp_minimize(lambda x: fg(x), x0, jac=True, method="L-BFGS-B",
                      options={"maxiter": 25, "ftol": 0, "gtol": 0, "maxcor": 20})
    opt = lm.minimize(lambda x: fg(x), x0, maxiter=25, m=20, bounded=False)
    assert opt.nit == 25
    # same algorithm without bounds (L-BFGS + More'-Thuente): same rate of convergence
    assert opt.f < 1e-6 and ref.fun < 1e-6


def test_history_wraps_and_budget():
    n = 30
    fg = _quadratic(n, 2)
    opt = lm.minimize(lambda x: fg(x), np.full(n, 0.5), maxiter=40, m=5)
    assert opt.nh <= 5 and opt.nit <= 40 and opt.nfev <= 3 * 40


def test_stationary_start_stops():
    fg = lambda x: (0.0, np.zeros_like(x))
    opt = lm.minimize(fg, np.full(8, 0.5), maxiter=10)
    assert opt.done and opt.nit == 0
