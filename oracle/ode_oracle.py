"""CPU restatement of ODE.jl's adaptive explicit Runge-Kutta drivers -- TEST INFRASTRUCTURE ONLY.

The reference's evolutions hand their right-hand-side closure to `ode45` / `ode78` of ODE.jl
(src/pdes.jl:62-68, 113-119, 206-213).  ODE.jl is a THIRD-PARTY dependency that is not under
/root/reference: pinned version 2.4.0 (Manifest.toml:184-188).  There is no network and no Julia in
this image, so this file restates the published algorithm of that version from its source as I know
it (ODE.jl/src/runge_kutta.jl: `oderk_adapt`, `rk_embedded_step!`, `calc_next_k!`, `stepsize_hw92!`,
`hinit`, `hermite_interp!`; tableaus `bt_dopri5` for ode45 = ode45_dp and `bt_feh78` for ode78).

PARITY STATUS: unpinned against ODE.jl itself.  Anchors: (a) the tableaus satisfy the Runge-Kutta
order conditions they claim (tests/test_oracle_pins.py checks the conditions up to order 5 and the
empirical orders), (b) the reference's own assertions at these call sites, test/solvers.jl:26-77
(sqrt(E) ~ 2 pi to 1e-7 in 1-D with ode45 in the position and hierarchical bases and with ode78;
sqrt(E) ~ sqrt(2) pi to 1e-4 and an energy drop in (0, 1e-8) in 2-D sparse), are asserted against this
restatement.  Defaults as in ODE.jl 2.4.0: reltol = 1e-5, abstol = 1e-8, norm = 2-norm of the scaled
error vector (NOT divided by sqrt(n)), maxstep = |tspan| / 2.5, minstep = |tspan| / 1e18,
points = :all, no step-size increase for 5 steps after a rejection.
"""
from __future__ import annotations

import math
from fractions import Fraction as Fr

import numpy as np


def _tab(rows):
    return [[float(Fr(x)) for x in r] for r in rows]


# Dormand-Prince 5(4)  (ODE.jl bt_dopri5, order (5, 4); first b row = 5th order solution)
DOPRI5 = {
    "name": "dopri5", "order": (5, 4),
    "a": _tab([
        ["0"] * 7,
        ["1/5", "0", "0", "0", "0", "0", "0"],
        ["3/40", "9/40", "0", "0", "0", "0", "0"],
        ["44/45", "-56/15", "32/9", "0", "0", "0", "0"],
        ["19372/6561", "-25360/2187", "64448/6561", "-212/729", "0", "0", "0"],
        ["9017/3168", "-355/33", "46732/5247", "49/176", "-5103/18656", "0", "0"],
        ["35/384", "0", "500/1113", "125/192", "-2187/6784", "11/84", "0"]]),
    "b": _tab([
        ["35/384", "0", "500/1113", "125/192", "-2187/6784", "11/84", "0"],
        ["5179/57600", "0", "7571/16695", "393/640", "-92097/339200", "187/2100", "1/40"]]),
    "c": [float(Fr(x)) for x in ["0", "1/5", "3/10", "4/5", "8/9", "1", "1"]],
}

# Fehlberg 7(8)  (ODE.jl bt_feh78, order (7, 8); first b row = 7th order solution)
FEH78 = {
    "name": "feh78", "order": (7, 8),
    "a": _tab([
        ["0"] * 13,
        ["2/27"] + ["0"] * 12,
        ["1/36", "1/12"] + ["0"] * 11,
        ["1/24", "0", "1/8"] + ["0"] * 10,
        ["5/12", "0", "-25/16", "25/16"] + ["0"] * 9,
        ["1/20", "0", "0", "1/4", "1/5"] + ["0"] * 8,
        ["-25/108", "0", "0", "125/108", "-65/27", "125/54"] + ["0"] * 7,
        ["31/300", "0", "0", "0", "61/225", "-2/9", "13/900"] + ["0"] * 6,
        ["2", "0", "0", "-53/6", "704/45", "-107/9", "67/90", "3"] + ["0"] * 5,
        ["-91/108", "0", "0", "23/108", "-976/135", "311/54", "-19/60", "17/6", "-1/12"] + ["0"] * 4,
        ["2383/4100", "0", "0", "-341/164", "4496/1025", "-301/82", "2133/4100", "45/82", "45/164", "18/41", "0", "0", "0"],
        ["3/205", "0", "0", "0", "0", "-6/41", "-3/205", "-3/41", "3/41", "6/41", "0", "0", "0"],
        ["-1777/4100", "0", "0", "-341/164", "4496/1025", "-289/82", "2193/4100", "51/82", "33/164", "12/41", "0", "1", "0"]]),
    "b": _tab([
        ["41/840", "0", "0", "0", "0", "34/105", "9/35", "9/35", "9/280", "9/280", "41/840", "0", "0"],
        ["0", "0", "0", "0", "0", "34/105", "9/35", "9/35", "9/280", "9/280", "0", "41/840", "41/840"]]),
    "c": [float(Fr(x)) for x in ["0", "2/27", "1/9", "1/6", "5/12", "1/2", "5/6", "1/6", "2/3", "1/3", "1", "0", "1"]],
}

TABLEAUS = {"45": DOPRI5, "78": FEH78}


def is_fsal(bt) -> bool:
    """isFSAL(btab): last stage = first b row, evaluated at the end of the step."""
    return bt["a"][-1] == bt["b"][0] and bt["c"][-1] == 1.0


def hinit(F, x0, t0, tend, p, reltol, abstol):
    """First step, direction of integration and F(t0, x0) (Hairer & Wanner II.4 p169)."""
    tdir = math.copysign(1.0, tend - t0)
    if tend == t0:
        raise ValueError("Zero time span")
    tau = max(reltol * np.linalg.norm(x0, np.inf), abstol)
    d0 = np.linalg.norm(x0, np.inf) / tau
    f0 = F(t0, x0)
    d1 = np.linalg.norm(f0, np.inf) / tau
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = 1e-6
    else:
        h0 = 0.01 * (d0 / d1)
    x1 = x0 + tdir * h0 * f0                           # Euler step
    f1 = F(t0 + tdir * h0, x1)
    d2 = np.linalg.norm(f1 - f0, np.inf) / (tau * h0)  # second-derivative estimate
    if max(d1, d2) <= 1e-15:
        h1 = max(1e-6, 1e-3 * h0)
    else:
        pw = -(2.0 + math.log10(max(d1, d2))) / (p + 1.0)
        h1 = 10.0 ** pw
    return tdir * min(100 * h0, h1, tdir * (tend - t0)), tdir, f0


def rk_embedded_step(ks, y, F, t, dt, bt):
    """One embedded step; ks[0] must hold F(t, y).  Returns (ytrial, yerr); fills ks[1:]."""
    a, b, c = bt["a"], bt["b"], bt["c"]
    S = len(c)
    ytrial = b[0][0] * ks[0]
    yerr = b[1][0] * ks[0]
    for s in range(1, S):
        ytmp = y.copy()                               # calc_next_k!: ytmp[d] += dt * ks[ss][d] * a[s, ss]
        for ss in range(s):
            ytmp += dt * ks[ss] * a[s][ss]
        ks[s] = F(t + c[s] * dt, ytmp)
        ytrial = ytrial + b[0][s] * ks[s]
        yerr = yerr + b[1][s] * ks[s]
    yerr = dt * (ytrial - yerr)
    ytrial = y + dt * ytrial
    return ytrial, yerr


def stepsize_hw92(dt, tdir, x0, xtrial, xerr, order, timeout, abstol, reltol, maxstep):
    """Error estimate and new step size (Hairer & Wanner 1992 p167, as modified in ODE.jl)."""
    timout_after_nan = 5
    fac = 0.8
    facmax = 5.0
    facmin = 1.0 / facmax
    if np.isnan(xtrial).any():                          # isoutofdomain
        return 10.0, dt * facmin, timout_after_nan
    xerr = xerr / (abstol + np.maximum(np.abs(x0), np.abs(xtrial)) * reltol)     # Eq 4.10
    err = float(np.linalg.norm(xerr, 2))                                          # Eq 4.11
    with np.errstate(divide="ignore"):
        newdt = min(maxstep, tdir * dt * max(facmin, fac * (1.0 / err) ** (1.0 / (order + 1)))) if err > 0 else \
            min(maxstep, tdir * dt * max(facmin, math.inf))
    if timeout > 0:
        newdt = min(newdt, dt)
        timeout -= 1
    return err, tdir * newdt, timeout


def hermite_interp(tquery, t, dt, y0, y1, f0, f1):
    theta = (tquery - t) / dt
    return ((1 - theta) * y0 + theta * y1 + theta * (theta - 1) *
            ((1 - 2 * theta) * (y1 - y0) + (theta - 1) * dt * f0 + theta * dt * f1))


def oderk_adapt(F, y0, tspan, bt, reltol=1.0e-5, abstol=1.0e-8, points="all", maxstep=None, minstep=None, initstep=0.0,
                stats=None):
    """ODE.jl oderk_adapt: returns (tout, yout) exactly as `ode45(F, y0, tspan)` does (lists)."""
    tspan = [float(t) for t in tspan]
    tstart, tend = tspan[0], tspan[-1]
    if maxstep is None:
        maxstep = abs(tend - tstart) / 2.5
    if minstep is None:
        minstep = abs(tend - tstart) / 1e18
    order = min(bt["order"])
    S = len(bt["c"])
    timeout_const = 5
    fsal = is_fsal(bt)
    y = np.array(y0, dtype=np.float64)
    ks = [None] * S
    ys = [y.copy()]
    tout = [tstart] if points == "all" else list(tspan)
    if points != "all":
        ys = [y.copy()] + [None] * (len(tspan) - 1)
    iter_fixed = 1
    it = 1
    t = tstart
    dt, tdir, ks[0] = hinit(F, y, tstart, tend, order, reltol, abstol)
    if initstep != 0:
        dt = initstep
    nacc = nrej = 0
    laststep = False
    timeout = 0
    while True:
        ytrial, yerr = rk_embedded_step(ks, y, F, t, dt, bt)
        err, newdt, timeout = stepsize_hw92(dt, tdir, y, ytrial, yerr, order, timeout, abstol, reltol, maxstep)
        if err <= 1.0:
            nacc += 1
            f0 = ks[0]
            f1 = ks[S - 1] if fsal else F(t + dt, ytrial)
            if points == "specified":
                while it < len(tspan) and (tdir * tspan[it] < tdir * (t + dt) or laststep):
                    ys[it] = hermite_interp(tspan[it], t, dt, y, ytrial, f0, f1)
                    it += 1
            else:
                while iter_fixed < len(tspan) and tdir * t < tdir * tspan[iter_fixed] < tdir * (t + dt):
                    ys.append(hermite_interp(tspan[iter_fixed], t, dt, y, ytrial, f0, f1))
                    tout.append(tspan[iter_fixed])
                    iter_fixed += 1
                ys.append(ytrial.copy())
                tout.append(t + dt)
            ks[0] = f1
            if laststep:
                break
            y = ytrial
            t += dt
            dt = newdt
            if tdir * (t + dt * 1.01) >= tdir * tend:
                dt = tend - t
                laststep = True
        elif abs(newdt) < minstep:
            break
        else:
            laststep = False
            nrej += 1
            dt = newdt
            timeout = timeout_const
    if stats is not None:
        stats.update(accepted=nacc, rejected=nrej)
    return tout, ys


def ode45(F, y0, tspan, **kw):
    return oderk_adapt(F, y0, tspan, DOPRI5, **kw)


def ode78(F, y0, tspan, **kw):
    return oderk_adapt(F, y0, tspan, FEH78, **kw)
