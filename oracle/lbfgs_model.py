"""More'-Thuente line search (dcsrch / dcstep) shared by the numpy optimiser models — TEST INFRASTRUCTURE ONLY.

oracle/lbfgsb_model.py restates L-BFGS-B (the reference's optimiser call, scipy L-BFGS-B, gatys/synthesizer.py:149-150)
and is pinned against scipy; its DEVICE options describe the variant syntex_b200/csrc/lbfgs.cu implements. This file
holds the line search both use (More' & Thuente 1994, with L-BFGS-B's constants ftol 1e-3, gtol 0.9, xtol 0.1, at
most 20 evaluations) and `minimize`, the numpy model of the device optimiser, so that (a) the algorithm can be
compared with scipy on the CPU oracle objective and (b) the CUDA kernels can be checked against it.
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


def minimize(fg, x0, maxiter, m=20, lo=0.0, hi=1.0, bounded=True, max_evals=None, callback=None):
    """The numpy model of the device optimiser: L-BFGS-B with the DEVICE options of lbfgsb_model (float32 vectors,
    subspace step from x with the active set of the current point, memory reset instead of the truncated step)."""
    import lbfgsb_model as lb
    return lb.minimize(fg, x0, maxiter, m=m, lo=lo, hi=hi, bounded=bounded, max_evals=max_evals or 3 * maxiter + 20,
                       callback=callback, **lb.DEVICE)
