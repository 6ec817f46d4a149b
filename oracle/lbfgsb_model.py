"""NumPy restatement of L-BFGS-B (bound-constrained limited-memory BFGS) — TEST INFRASTRUCTURE ONLY.

The reference's optimiser (gatys/synthesizer.py:30-36, :146-150) is scipy.optimize.minimize(method="L-BFGS-B"),
a third-party dependency whose version the reference does not pin; scipy 1.18.1 (the one in this image) ships a C
translation of L-BFGS-B 3.0 as a binary extension, so its source is not under /root/reference either. This file
restates the PUBLISHED algorithm:

  Byrd, Lu, Nocedal, Zhu, "A limited memory algorithm for bound constrained optimization", SIAM J. Sci. Comput.
  16 (1995): generalized Cauchy point (section 4, algorithm CP), subspace minimisation by the direct primal method
  (section 5.1), compact representation B = theta I - W M W', W = [Y, theta S].
  Morales, Nocedal, "Remark on algorithm 778" (2011): the projected subspace step of version 3.0.
  More', Thuente, "Line search algorithms with guaranteed sufficient decrease" (1994): dcsrch / dcstep, with
  L-BFGS-B's constants ftol 1e-3, gtol 0.9, xtol 0.1, at most 20 evaluations per search (lbfgs_model.py holds the
  line search; it is shared).

and is PINNED against scipy itself in tests/test_lbfgsb_model.py: same iterates (to rounding) on bounded quadratics
and on the float32 oracle objective. It serves two purposes: (a) the cycle-by-cycle model of the device optimiser
(syntex_b200/csrc/lbfgs.cu, L-BFGS-B mode), (b) a place to measure which parts of L-BFGS-B the final loss of a
texture synthesis depends on (switches `gcp` and `subspace`).

Differences from the Fortran that do not change iterates: the 2m x 2m middle matrix and the reduced system are
solved densely in float64 (numpy) instead of through the LEL' / Cholesky factorisations of formk / formt; the
breakpoints are sorted once instead of heap-extracted.
"""
from __future__ import annotations

import numpy as np

from lbfgs_model import ls_init, ls_update

EPSMCH = np.finfo(np.float64).eps


class LbfgsB:
    """One bound-constrained problem, reverse communication: opt.start(x0, f0, g0); while not opt.done:
    f, g = fg(opt.xt); opt.feed(f, g).  Vectors float64 (as scipy) unless `vec_dtype` says otherwise."""

    def __init__(self, n, m=20, lo=0.0, hi=1.0, bounded=True, vec_dtype=np.float64, gcp="exact", subspace=True,
                 x_dtype=np.float64, backtrack="lbfgsb", hist="exact"):
        self.n, self.m = n, m
        self.bounded = bounded
        self.lo = np.full(n, lo, np.float64)
        self.hi = np.full(n, hi, np.float64)
        self.vt = vec_dtype
        self.hist = hist            # "f16": history vectors kept as fp16 with a power-of-two scale per vector (device)
        self.gcp = gcp              # "exact": algorithm CP; "first": first segment only (no breakpoint walk);
                                    # "none": no Cauchy step once there is a history (xcp = x, active set = variables at a
                                    # bound whose gradient points outwards) - what the CUDA optimiser does
        self.subspace = subspace
        self.xdt = x_dtype          # dtype of x, g, d and the trial point (float32 on the device)
        self.backtrack = backtrack  # "lbfgsb": truncate at the first bound when the projected step is not a descent
                                    # direction (subsm of version 3.0); "reset": drop the memory instead (device)
        self.S = np.zeros((m, n), vec_dtype)
        self.Y = np.zeros((m, n), vec_dtype)
        self.col = 0
        self.theta = 1.0
        self.SY = np.zeros((m, m))
        self.SS = np.zeros((m, m))
        self.nit = 0
        self.nfev = 0
        self.done = False
        self.status = ""
        self.nseg_total = 0
        self.nskip = 0
        self.nreset = 0
        self.nproj = 0
        self.nbacktrack = 0

    # ------------------------------------------------------------------ compact representation
    def _middle(self):
        """M = [[-D, L'], [L, theta S'S]]^-1 (2col x 2col), BLNZ eq. (3.4)."""
        k = self.col
        SY, SS = self.SY[:k, :k], self.SS[:k, :k]
        D = np.diag(np.diag(SY))
        L = np.tril(SY, -1)
        Minv = np.block([[-D, L.T], [L, self.theta * SS]])
        return Minv

    def _W(self, idx=None):
        """W = [Y, theta S] as a (2col, n) array (rows = columns of W)."""
        k = self.col
        Y = self.Y[:k].astype(np.float64)
        S = self.S[:k].astype(np.float64)
        if idx is not None:
            Y, S = Y[:, idx], S[:, idx]
        return np.concatenate([Y, self.theta * S], axis=0)

    # ------------------------------------------------------------------ generalized Cauchy point
    def _cauchy(self, x, g):
        n, k = self.n, self.col
        lo, hi = self.lo, self.hi
        if self.bounded:
            at_lo = (x - lo) <= 0.0
            at_hi = (hi - x) <= 0.0
            fixed = (at_lo & (g >= 0.0)) | (at_hi & (g <= 0.0))     # iwhere 1 / 2: at a bound, gradient pushing outwards
        else:
            fixed = np.zeros(n, bool)
        d = np.where(fixed, 0.0, -g)
        iwhere_fixed = fixed.copy()
        xcp = x.copy()
        if self.gcp == "none" and k > 0:
            return xcp, np.zeros(2 * k), iwhere_fixed
        if not np.any(d != 0.0):
            return xcp, np.zeros(2 * k), iwhere_fixed
        if self.bounded:
            with np.errstate(divide="ignore", invalid="ignore"):
                t = np.where(d < 0.0, (x - lo) / (-d), np.where(d > 0.0, (hi - x) / d, np.inf))
        else:
            t = np.full(n, np.inf)
        W = self._W() if k > 0 else None
        p = W @ d if k > 0 else np.zeros(0)
        c = np.zeros(2 * k)
        f1 = -float(d @ d)
        f2 = -self.theta * f1
        f2_org = f2
        if k > 0:
            Minv = self._middle()
            f2 -= float(p @ np.linalg.solve(Minv, p))
        dtm = -f1 / f2
        tsum = 0.0
        nseg = 1
        finite = np.flatnonzero(np.isfinite(t))
        all_break = finite.size == n
        if finite.size > 0 and self.gcp == "exact":
            order = finite[np.argsort(t[finite], kind="stable")]
            tj = 0.0
            nleft = order.size
            hit_all = False
            for ibp in order:
                tj0 = tj
                tj = t[ibp]
                dt = tj - tj0
                if dtm < dt:
                    break
                tsum += dt
                nleft -= 1
                dibp = d[ibp]
                d[ibp] = 0.0
                if dibp > 0.0:
                    zibp = hi[ibp] - x[ibp]
                    xcp[ibp] = hi[ibp]
                else:
                    zibp = lo[ibp] - x[ibp]
                    xcp[ibp] = lo[ibp]
                iwhere_fixed[ibp] = True
                if nleft == 0 and all_break:
                    dtm = dt
                    hit_all = True
                    break
                nseg += 1
                dibp2 = dibp * dibp
                f1 = f1 + dt * f2 + dibp2 - self.theta * dibp * zibp
                f2 = f2 - self.theta * dibp2
                if k > 0:
                    c += dt * p
                    wbp = W[:, ibp]
                    v = np.linalg.solve(Minv, wbp)
                    wmc, wmp, wmw = float(c @ v), float(p @ v), float(wbp @ v)
                    p = p - dibp * wbp
                    f1 += dibp * wmc
                    f2 += 2.0 * dibp * wmp - dibp2 * wmw
                f2 = max(EPSMCH * f2_org, f2)
                if nleft > 0:
                    dtm = -f1 / f2
                else:
                    # every finite breakpoint passed; variables without a breakpoint keep moving
                    dtm = -f1 / f2 if np.any(d != 0.0) else 0.0
            self.nseg_total += nseg
            if hit_all:
                if k > 0:
                    c += dtm * p
                return xcp, c, iwhere_fixed
        dtm = max(dtm, 0.0)
        tsum += dtm
        xcp = xcp + tsum * d
        if self.gcp != "exact" and self.bounded:
            # "first": the unconstrained Cauchy step of the first segment, clipped; whatever crossed a bound is fixed
            crossed = (xcp < lo) | (xcp > hi)
            xcp = np.clip(xcp, lo, hi)
            iwhere_fixed |= crossed
            if k > 0:
                c = W @ (xcp - x)
                return xcp, c, iwhere_fixed
        if k > 0:
            c += dtm * p
        return xcp, c, iwhere_fixed

    # ------------------------------------------------------------------ subspace minimisation (direct primal)
    def _subspace(self, x, g, xcp, c, fixed):
        k = self.col
        free = np.flatnonzero(~fixed)
        if free.size == 0 or k == 0 or not self.subspace:
            return xcp
        theta = self.theta
        Minv = self._middle()
        Wf = self._W(free)
        # r = -Z'(B (xcp - x) + g)
        r = -g[free] - theta * (xcp[free] - x[free]) + Wf.T @ np.linalg.solve(Minv, c)
        # d = B_ff^-1 r,  B_ff = theta I - W_f M W_f'   (Sherman-Morrison-Woodbury)
        if self.subspace == "php":
            W = self._W()
            K = Minv - (W @ W.T) / theta          # inverse of the FULL B, restricted to the free set on both sides
        else:
            K = Minv - (Wf @ Wf.T) / theta
        try:
            wv = np.linalg.solve(K, Wf @ r)
        except np.linalg.LinAlgError:
            return None
        d = r / theta + (Wf.T @ wv) / (theta * theta)
        xs = xcp.copy()
        if not self.bounded:
            xs[free] = xcp[free] + d
            return xs
        trial = xcp[free] + d
        proj = np.clip(trial, self.lo[free], self.hi[free])
        xs[free] = proj
        if np.any(proj != trial):
            self.nproj += 1
            if float((xs - x) @ g) > 0.0:
                # projected point is not a descent direction: classic truncation of the step at the first bound
                self.nbacktrack += 1
                if self.backtrack == "reset":
                    return None
                xf = xcp[free]
                with np.errstate(divide="ignore", invalid="ignore"):
                    a = np.where(d < 0.0, (self.lo[free] - xf) / d, np.where(d > 0.0, (self.hi[free] - xf) / d, np.inf))
                a = np.where(np.isnan(a), np.inf, a)
                ibd = int(np.argmin(a))
                alpha = min(1.0, max(float(a[ibd]), 0.0))
                xs = xcp.copy()
                xs[free] = xf + alpha * d
                if alpha < 1.0:
                    xs[free[ibd]] = self.hi[free[ibd]] if d[ibd] > 0 else self.lo[free[ibd]]
        return xs

    # ------------------------------------------------------------------ one iteration set-up
    def _direction(self):
        x, g = self.x, self.g
        x64, g64 = x.astype(np.float64), g.astype(np.float64)
        while True:
            if not self.bounded and self.col > 0:
                xcp, c, fixed = x64.copy(), np.zeros(2 * self.col), np.zeros(self.n, bool)
            else:
                xcp, c, fixed = self._cauchy(x64, g64)
            z = self._subspace(x64, g64, xcp, c, fixed)
            if z is None:
                self._reset()
                continue
            z = z.astype(self.xdt)
            d = z - x
            gd = float(g.astype(np.float64) @ d.astype(np.float64))
            if not (gd < 0.0):
                if self.col > 0:
                    self._reset()
                    continue
                self.done = True
                self.status = "ABNORMAL: no descent direction"
                return
            break
        self.d, self.z, self.gd = d, z, gd
        # lnsrlb: largest feasible step
        stpmx = 1e10
        if self.bounded:
            if self.nit == 0:
                stpmx = 1.0
            else:
                with np.errstate(divide="ignore", invalid="ignore"):
                    d64 = d.astype(np.float64)
                    a = np.where(d64 < 0.0, (self.lo - x64) / d64, np.where(d64 > 0.0, (self.hi - x64) / d64, np.inf))
                a = np.where(np.isnan(a), np.inf, a)
                stpmx = min(stpmx, max(float(a.min()), 0.0))
                if self.xdt != np.float64:
                    stpmx = max(stpmx, 1.0)     # d = z - x is feasible at step 1 by construction; rounding may say 0.9999999
        dnorm = float(np.sqrt(d.astype(np.float64) @ d.astype(np.float64)))
        stp = min(1.0 / dnorm, stpmx) if (self.nit == 0 and not self.bounded) else 1.0
        self.ls = ls_init(self.f, gd, stp)
        self.ls["stmax_user"] = stpmx
        self.stpmx = stpmx
        self._trial(stp)

    def _trial(self, stp):
        self.stp = stp
        if stp == 1.0:
            self.xt = self.z.copy()
        else:
            xt = self.x + self.xdt(stp) * self.d
            self.xt = np.clip(xt, self.lo.astype(self.xdt), self.hi.astype(self.xdt)) if self.bounded else xt

    def _reset(self):
        self.col = 0
        self.theta = 1.0
        self.nreset += 1

    # ------------------------------------------------------------------ driver interface
    def start(self, x0, f0, g0):
        x = np.asarray(x0, self.xdt).ravel().copy()
        if self.bounded:
            x = np.clip(x, self.lo.astype(self.xdt), self.hi.astype(self.xdt))
        self.x, self.f, self.g = x, float(f0), np.asarray(g0, self.xdt).ravel().copy()
        self.nfev = 1
        self._direction()

    def feed(self, ft, gt):
        gt = np.asarray(gt, self.xdt).ravel()
        self.nfev += 1
        gdt = float(gt.astype(np.float64) @ self.d.astype(np.float64))
        res = ls_update_bounded(self.ls, float(ft), gdt, self.stpmx)
        if res == "again":
            self._trial(self.ls["stp"])
            return
        if res == "fail":
            if self.col == 0:
                self.done = True
                self.status = "ABNORMAL_TERMINATION_IN_LNSRCH"
                return
            self._reset()
            self._direction()
            return
        # accept (CONVERGENCE or WARNING of dcsrch: lnsrlb takes both as the new iterate)
        stp = self.stp
        r = (gt - self.g).astype(np.float64)
        if self.xdt != np.float64:
            # the device forms s = xt - x (after rounding and clipping) and takes s'y from the stored vectors
            s_dev = (self.xt - self.x).astype(np.float64)
            dr_dev = float(s_dev @ r)
        if stp == 1.0:
            dr = gdt - self.gd
            ddum = -self.gd
            s = self.d
        else:
            dr = (gdt - self.gd) * stp
            s = self.d * stp
            ddum = -self.gd * stp
        if self.xdt != np.float64:
            s, dr = s_dev, dr_dev
        fold = self.f
        self.x, self.f, self.g = self.xt, float(ft), gt.copy()
        self.nit += 1
        if fold - self.f <= 0.0:
            # factr = 0 (the reference passes ftol = 0): mainlb still stops when f did not decrease at all
            self.done = True
            self.status = "CONVERGENCE: REL_REDUCTION_OF_F_<=_FACTR*EPSMCH"
            return
        if dr <= EPSMCH * ddum:
            self.nskip += 1
        else:
            self._update(s, r, dr)
        self._direction()

    @property
    def nh(self):
        return self.col

    def _stored(self, v):
        """A vector as the history holds it."""
        v = v.astype(self.vt)
        if self.hist != "f16":
            return v
        v32 = v.astype(np.float32)
        sc = np.float32(f16_scale(float(np.dot(v32.astype(np.float64), v32.astype(np.float64)))))
        return ((v32 / sc).astype(np.float16).astype(np.float32) * sc).astype(self.vt)

    def _update(self, s, y, dr):
        m = self.m
        if self.col == m:
            self.S[:-1] = self.S[1:]
            self.Y[:-1] = self.Y[1:]
            self.SY[:-1, :-1] = self.SY[1:, 1:]
            self.SS[:-1, :-1] = self.SS[1:, 1:]
            self.col -= 1
        k = self.col
        self.S[k], self.Y[k] = self._stored(s), self._stored(y)
        Sd, Yd = self.S[:k + 1].astype(np.float64), self.Y[:k + 1].astype(np.float64)
        if self.hist == "f16":      # the tables see the pair as it was measured, the older vectors as they are stored
            Sd[k], Yd[k] = s.astype(self.vt), y.astype(self.vt)
        sk, yk = Sd[k], Yd[k]
        self.SY[:k + 1, k] = Sd @ yk
        self.SY[k, :k + 1] = Yd @ sk
        self.SS[:k + 1, k] = Sd @ sk
        self.SS[k, :k + 1] = self.SS[:k + 1, k]
        self.SY[k, k] = dr
        self.theta = float(yk @ yk) / dr
        self.col = k + 1


def f16_scale(norm2):
    """Power of two that maps a vector of squared 2-norm `norm2` below 2^15 in every element (csrc/lbfgs.cu,
    hist_scale): stored = v / scale fits fp16 (max 65504) whatever the distribution of its elements."""
    nrm = float(np.sqrt(norm2))
    if not (nrm > 0.0) or not np.isfinite(nrm):
        return 1.0
    _, e = np.frexp(nrm)            # nrm = f * 2^e, 0.5 <= f < 1
    return float(np.ldexp(1.0, int(e) - 15))


def ls_update_bounded(st, f, g, stpmx):
    """dcsrch as lnsrlb drives it: stpmax is the largest feasible step and dcsrch's WARNING outcomes (rounding errors
    prevent progress, xtol test satisfied, step at stpmax / stpmin) END the search with the current point ACCEPTED
    (lnsrlb: task = NEW_X unless the status is neither CONVERGENCE nor WARNING). Returns 'accept' | 'again' | 'fail'."""
    from lbfgs_model import FTOL, GTOL, XTOL, XTRAPL, XTRAPU, MAX_LS, dcstep
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
    if stp == stpmx and f <= ftest and g <= st["gtest"]:
        warn = True
    if stp == 0.0 and (f > ftest or g >= st["gtest"]):
        warn = True
    if f <= ftest and abs(g) <= GTOL * (-st["ginit"]):
        return "accept"
    if warn:
        return "accept"
    if st["nev"] >= MAX_LS:
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
    new = min(stpmx, max(0.0, new))
    if st["brackt"] and (new <= st["stmin"] or new >= st["stmax"] or
                         st["stmax"] - st["stmin"] <= XTOL * st["stmax"]):
        new = st["stx"]
    st["stp"] = new
    return "again"


DEVICE = dict(gcp="none", vec_dtype=np.float32, x_dtype=np.float32, backtrack="reset", hist="f16")
"""The variant syntex_b200/csrc/lbfgs.cu implements (keyword arguments of LbfgsB): float32 vectors, the (s, y) history
stored as fp16 with a power-of-two scale per vector (half the bytes of the two history passes; final losses against an
fp32 history on the texture objective, 128^2, 100 iterations, 8 seeds: geometric mean 1.00, single seeds +-5 % as between
any two chaotic trajectories), the subspace step
taken from x itself with the active set of the current point (no breakpoint walk: measured on the texture objective,
tests/test_lbfgsb_model.py, final losses are indistinguishable from the exact generalized Cauchy point), and a
memory reset in the (never observed) case that the projected step is not a descent direction."""


def minimize(fg, x0, maxiter, m=20, lo=0.0, hi=1.0, bounded=True, max_evals=None, callback=None, **kw):
    x0 = np.asarray(x0, np.float64).ravel()
    opt = LbfgsB(x0.size, m, lo, hi, bounded, **kw)
    if bounded:
        x0 = np.clip(x0, lo, hi)
    f0, g0 = fg(x0)
    opt.start(x0, f0, g0)
    max_evals = max_evals or 15000
    while not opt.done and opt.nit < maxiter and opt.nfev < max_evals:
        ft, gt = fg(opt.xt)
        nit = opt.nit
        opt.feed(ft, gt)
        if callback is not None and opt.nit != nit:
            callback(opt)
    return opt
