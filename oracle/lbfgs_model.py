"""NumPy model of the device-resident projected L-BFGS (syntex_b200/csrc/lbfgs.cu) — TEST INFRASTRUCTURE ONLY.

The CUDA optimiser is a state machine that advances by exactly one f/g evaluation per cycle. This file is the
same state machine written with numpy (fp32 vectors, fp64 scalars) so that (a) the algorithm can be compared
with the reference's optimiser call, scipy L-BFGS-B (gatys/synthesizer.py:149-150), on the CPU oracle
objective, and (b) the CUDA kernels can be checked cycle by cycle against it.

Algorithm (one independent problem):
  direction   d = -P H P g, P = projector onto the free variables (not at a bound with the gradient pushing
              outwards), H = L-BFGS inverse Hessian in compact form (Byrd, Nocedal, Schnabel 1994, eq. 3.1)
              built from the last m accepted (s, y) pairs with s'y > eps y'y; first step length min(1, 1/|Pg|).
  line search More'-Thuente (1994) with L-BFGS-B's constants ftol 1e-3, gtol 0.9, xtol 0.1, <= 20 evaluations,
              along the projected path x(a) = clip(x + a d), directional derivative taken over unclipped
              coordinates. On failure the memory is dropped and the search restarts from steepest descent
              (as L-BFGS-B does).
"""
from __future__ import annotations

import numpy as np

FTOL, GTOL, XTOL = 1e-3, 0.9, 0.1
XTRAPL, XTRAPU = 1.1, 4.0
STPMAX = 1e10
MAX_LS = 20
EPS = 2.2e-16


def dcstep(st, stp, fp, dp, stpmin, stpmax):
    """One safeguarded cubic/quadratic interpolation step (More' & Thuente 1994, section 4).
    st: dict with stx, fx, dx, sty, fy, dy, brackt. Returns the new trial step."""
    stx, fx, dx, sty, fy, dy, brackt = st["stx"], st["fx"], st["dx"], st["sty"], st["fy"], st["dy"], st["brackt"]
    sgnd = dp * (dx / abs(dx))
    if fp > fx:
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * np.sqrt(max(0.0, (theta / s) ** 2 - (dx / s) * (dp / s)))
        if stp < stx:
            gamma = -gamma
        p = (gamma - dx) + theta
        q = ((gamma - dx) + gamma) + dp
        r = p / q
        stpc = stx + r * (stp - stx)
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx)
        stpf = stpc if abs(stpc - stx) < abs(stpq - stx) else stpc + (stpq - stpc) / 2.0
        brackt = True
    elif sgnd < 0.0:
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * np.sqrt(max(0.0, (theta / s) ** 2 - (dx / s) * (dp / s)))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = ((gamma - dp) + gamma) + dx
        r = p / q
        stpc = stp + r * (stx - stp)
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
        brackt = True
    elif abs(dp) < abs(dx):
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * np.sqrt(max(0.0, (theta / s) ** 2 - (dx / s) * (dp / s)))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = (gamma + (dx - dp)) + gamma
        r = p / q
        if r < 0.0 and gamma != 0.0:
            stpc = stp + r * (stx - stp)
        elif stp > stx:
            stpc = stpmax
        else:
            stpc = stpmin
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        if brackt:
            stpf = stpc if abs(stpc - stp) < abs(stpq - stp) else stpq
            if stp > stx:
                stpf = min(stp + 0.66 * (sty - stp), stpf)
            else:
                stpf = max(stp + 0.66 * (sty - stp), stpf)
        else:
            stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
            stpf = min(stpmax, max(stpmin, stpf))
    else:
        if brackt:
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp
            s = max(abs(theta), abs(dy), abs(dp))
            gamma = s * np.sqrt(max(0.0, (theta / s) ** 2 - (dy / s) * (dp / s)))
            if stp > sty:
                gamma = -gamma
            p = (gamma - dp) + theta
            q = ((gamma - dp) + gamma) + dy
            r = p / q
            stpf = stp + r * (sty - stp)
        elif stp > stx:
            stpf = stpmax
        else:
            stpf = stpmin
    if fp > fx:
        sty, fy, dy = stp, fp, dp
    else:
        if sgnd < 0.0:
            sty, fy, dy = stx, fx, dx
        stx, fx, dx = stp, fp, dp
    st.update(stx=stx, fx=fx, dx=dx, sty=sty, fy=fy, dy=dy, brackt=brackt)
    return stpf


def ls_init(f0, g0, stp):
    return dict(brackt=False, stage=1, finit=f0, ginit=g0, gtest=FTOL * g0, width=STPMAX, width1=2 * STPMAX,
                stx=0.0, fx=f0, dx=g0, sty=0.0, fy=f0, dy=g0, stmin=0.0, stmax=stp + XTRAPU * stp, stp=stp, nev=0)


def ls_update(st, f, g):
    """Returns 'accept', 'fail' or 'again' (with st['stp'] set to the next trial step)."""
    stp = st["stp"]
    st["nev"] += 1
    ftest = st["finit"] + stp * st["gtest"]
    if st["stage"] == 1 and f <= ftest and g >= 0.0:
        st["stage"] = 2
    warn = False
    if st["brackt"] and (stp <= st["stmin"] or stp >= st["stmax"]):
        warn = True
    if st["brackt"] and st["stmax"] - st["stmin"] <= XTOL * st["stmax"]:
        warn = True
    if stp == STPMAX and f <= ftest and g <= st["gtest"]:
        warn = True
    if stp == 0.0 and (f > ftest or g >= st["gtest"]):
        warn = True
    if f <= ftest and abs(g) <= GTOL * (-st["ginit"]):
        return "accept"
    if warn or st["nev"] >= MAX_LS:
        return "fail"
    if st["stage"] == 1 and f <= st["fx"] and f > ftest:
        gt = st["gtest"]
        m = dict(stx=st["stx"], fx=st["fx"] - st["stx"] * gt, dx=st["dx"] - gt, sty=st["sty"],
                 fy=st["fy"] - st["sty"] * gt, dy=st["dy"] - gt, brackt=st["brackt"])
        new = dcstep(m, stp, f - stp * gt, g - gt, st["stmin"], st["stmax"])
        st.update(stx=m["stx"], fx=m["fx"] + m["stx"] * gt, dx=m["dx"] + gt, sty=m["sty"],
                  fy=m["fy"] + m["sty"] * gt, dy=m["dy"] + gt, brackt=m["brackt"])
    else:
        new = dcstep(st, stp, f, g, st["stmin"], st["stmax"])
    if st["brackt"]:
        if abs(st["sty"] - st["stx"]) >= 0.66 * st["width1"]:
            new = st["stx"] + 0.5 * (st["sty"] - st["stx"])
        st["width1"] = st["width"]
        st["width"] = abs(st["sty"] - st["stx"])
        st["stmin"] = min(st["stx"], st["sty"])
        st["stmax"] = max(st["stx"], st["sty"])
    else:
        st["stmin"] = new + XTRAPL * (new - st["stx"])
        st["stmax"] = new + XTRAPU * (new - st["stx"])
    new = min(STPMAX, max(0.0, new))
    if st["brackt"] and (new <= st["stmin"] or new >= st["stmax"] or
                         st["stmax"] - st["stmin"] <= XTOL * st["stmax"]):
        new = st["stx"]
    st["stp"] = new
    return "again"


class ProjectedLbfgs:
    """One problem; drive with: opt.start(x0, f0, g0); while ...: f, g = fg(opt.xt); opt.feed(f, g)."""

    def __init__(self, n, m=20, lo=0.0, hi=1.0, bounded=True, masked_history=False):
        self.n, self.m = n, m
        self.lo, self.hi, self.bounded = np.float32(lo), np.float32(hi), bounded
        self.masked_history = masked_history
        self.S = np.zeros((m, n), np.float32)
        self.Y = np.zeros((m, n), np.float32)
        self.nh = 0
        self.SY = np.zeros((m, m))   # SY[i, j] = s_i . y_j  (chronological order 0 = oldest)
        self.YY = np.zeros((m, m))
        self.nit = 0
        self.nfev = 0
        self.done = False

    # -- helpers
    def _free(self, x, g):
        if not self.bounded:
            return np.ones(self.n, bool)
        return ~(((x <= self.lo) & (g > 0)) | ((x >= self.hi) & (g < 0)))

    def _project(self, z):
        return np.clip(z, self.lo, self.hi) if self.bounded else z

    def _direction(self):
        x, g = self.x, self.g
        free = self._free(x, g)
        pg = np.where(free, g, np.float32(0))
        pgn = float(np.sqrt(np.dot(pg.astype(np.float64), pg.astype(np.float64))))
        k = self.nh
        if k == 0:
            d = -pg
            stp0 = min(1.0, 1.0 / pgn) if pgn > 0 else 1.0
        else:
            S, Y = self.S[:k], self.Y[:k]
            if self.masked_history:
                Sm, Ym = S * free, Y * free
                SY = Sm.astype(np.float64) @ Ym.astype(np.float64).T
                YY = Ym.astype(np.float64) @ Ym.astype(np.float64).T
                S, Y = Sm, Ym
            else:
                SY, YY = self.SY[:k, :k], self.YY[:k, :k]
            gamma = SY[k - 1, k - 1] / YY[k - 1, k - 1]
            R = np.triu(SY)
            D = np.diag(np.diag(SY))
            p1 = S.astype(np.float64) @ pg.astype(np.float64)
            p2 = gamma * (Y.astype(np.float64) @ pg.astype(np.float64))
            Rinv_p1 = np.linalg.solve(R, p1)
            a = np.linalg.solve(R.T, (D + gamma * YY) @ Rinv_p1 - p2)
            b = -Rinv_p1
            hv = gamma * pg.astype(np.float64) + S.T.astype(np.float64) @ a + gamma * (Y.T.astype(np.float64) @ b)
            d = np.where(free, -hv, 0.0).astype(np.float32)
            stp0 = 1.0
        dg0 = float(np.dot(g.astype(np.float64), d.astype(np.float64)))
        if not (dg0 < 0.0):
            if k > 0:
                self.nh = 0
                return self._direction()
            self.done = True
            return
        self.d = d
        self.ls = ls_init(self.f, dg0, stp0)
        self.xt = self._project(x + np.float32(stp0) * d)

    # -- driver interface
    def start(self, x0, f0, g0):
        self.x = self._project(np.asarray(x0, np.float32).ravel())
        self.f = float(f0)
        self.g = np.asarray(g0, np.float32).ravel()
        self.nfev = 1
        self._direction()

    def feed(self, ft, gt):
        """Consume f/g at self.xt."""
        gt = np.asarray(gt, np.float32).ravel()
        self.nfev += 1
        step = self.ls["stp"]
        z = self.x + np.float32(step) * self.d
        unclipped = (z >= self.lo) & (z <= self.hi) if self.bounded else np.ones(self.n, bool)
        dphi = float(np.dot(np.where(unclipped, gt, 0).astype(np.float64), self.d.astype(np.float64)))
        res = ls_update(self.ls, float(ft), dphi)
        if res == "accept":
            s = self.xt - self.x
            y = gt - self.g
            ys = float(np.dot(s.astype(np.float64), y.astype(np.float64)))
            yy = float(np.dot(y.astype(np.float64), y.astype(np.float64)))
            if ys > EPS * yy:
                if self.nh == self.m:
                    self.S[:-1] = self.S[1:]
                    self.Y[:-1] = self.Y[1:]
                    self.SY[:-1, :-1] = self.SY[1:, 1:]
                    self.YY[:-1, :-1] = self.YY[1:, 1:]
                    self.nh -= 1
                k = self.nh
                self.S[k], self.Y[k] = s, y
                Sd, Yd = self.S[:k + 1].astype(np.float64), self.Y[:k + 1].astype(np.float64)
                self.SY[:k + 1, k] = Sd @ y.astype(np.float64)
                self.SY[k, :k + 1] = Yd @ s.astype(np.float64)
                self.YY[:k + 1, k] = Yd @ y.astype(np.float64)
                self.YY[k, :k + 1] = self.YY[:k + 1, k]
                self.nh = k + 1
            self.x, self.f, self.g = self.xt, float(ft), gt
            self.nit += 1
            self._direction()
        elif res == "fail":
            if self.nh > 0:
                self.nh = 0
                self._direction()
            else:
                self.done = True
        else:
            self.xt = self._project(self.x + np.float32(self.ls["stp"]) * self.d)


def minimize(fg, x0, maxiter, m=20, lo=0.0, hi=1.0, bounded=True, max_evals=None, masked_history=False,
             callback=None):
    x0 = np.asarray(x0, np.float32).ravel()
    opt = ProjectedLbfgs(x0.size, m, lo, hi, bounded, masked_history)
    x0 = opt._project(x0)
    f0, g0 = fg(x0)
    opt.start(x0, f0, g0)
    max_evals = max_evals or 3 * maxiter
    while not opt.done and opt.nit < maxiter and opt.nfev < max_evals:
        ft, gt = fg(opt.xt)
        nit = opt.nit
        opt.feed(ft, gt)
        if callback is not None and opt.nit != nit:
            callback(opt)
    return opt
