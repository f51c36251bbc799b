"""CPU ORACLE (test infrastructure, not product code) for Shockwave's per-round schedule solve.

This file restates, on scipy's bundled HiGHS (`scipy.optimize.milp`), the mixed-integer program
that the reference builds with cvxpy and hands to Gurobi in
`scheduler/shockwave.py:504-711` (`dynamic_eisenberg_gale_scheduling`).  Gurobi/cvxpy are not
installable here (no package, no licence, no network) so every number produced by this module is a
"HiGHS stand-in for Gurobi".  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it.

Pinning status: the reference holds NO known-answer test for the solver, so at
the x-matrix level this oracle is "parity unpinned".  What pins it: (1) `tests/golden/make_solve_fixtures.py` (via oracle/ref_harness.py)
runs the UNMODIFIED reference simulator (`scheduler/scheduler.py`) with this module substituted for
the Gurobi call and compares the end-to-end metrics with the golden pickle the reference ships
(`scheduler/reproduce/pickles/tacc_32gpus/shockwave_*.pickle`); (2) the forecast half is checked
value-for-value against the reference's own `JobMetaData` (see `oracle/jobmeta.py`).

Variable / constraint map (reference file:line -> here):
  x[j][t] boolean                         shockwave.py:288-294   -> columns [0, J*T)
  sum_j g_j x[j][t] <= G                  shockwave.py:297-319   -> rows "cap"
  p_j >= 0 ; dbar_j p_j <= D sum_t x      shockwave.py:369-377   -> columns P, rows "prog"
  lambda_{j,b} >= 0, sum lambda*base = (c_j+p_j)/E_j, sum lambda = 1
                                          shockwave.py:384-402   -> columns L, rows "pwl_u","pwl_1"
  z_{j,b} boolean SOS2 rows               shockwave.py:403-419   -> optional (with_sos2=True);
        redundant for a concave maximise (the LP already picks adjacent breakpoints) — verified by
        tests/test_oracle_milp.py::test_sos2_redundant
  rem_j = max(0, R_j - dbar_j p_j)        shockwave.py:555-563
  maximise sum_j w_j plog_j/(J*T) - k*max_j rem_j
                                          shockwave.py:565-568   -> scalar column M, rows "mk"
  D(r+T) + rem_j/share <= rhomax*FTobj_j  shockwave.py:573-597   -> rows "ftf"
  fallback priorities                     shockwave.py:830-911   -> relax_priorities()
  second MILP (re-rank rounds)            shockwave.py:714-793   -> rank_in_schedule()
  rounding + work-conserving back-fill    shockwave.py:213-285   -> construct_schedules()
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import scipy.sparse as sp
from scipy.optimize import Bounds, LinearConstraint, milp

STATUS_FTF_FEASIBLE = 0   # first solve (with finish-time-fairness rows) had a solution
STATUS_FALLBACK = 1       # FTF rows infeasible -> relaxed objective + re-rank (shockwave.py:631-706)


def pwl_log_values(bases, origin):
    """log(base_b) with log(0) -> log(origin[0.0])   (shockwave.py:339-352)."""
    assert bases[0] == 0.0
    vals = []
    for b in bases:
        assert 0.0 <= b <= 1.0
        vals.append(math.log(origin[0.0]) if b == 0.0 else math.log(b))
    assert all(p < n for p, n in zip(vals, vals[1:]))
    return np.asarray(vals, dtype=np.float64)


def plog(u, bases, logv):
    """Piecewise-linear interpolation of log through (base_b, logv_b); u in [0,1]."""
    return np.interp(u, np.asarray(bases, dtype=np.float64), logv)


def finish_time_momentumed_average(series, rnd, momentum=0.9):
    """shockwave.py:480-501, value-for-value (same float64 operation order)."""
    assert len(series) > 0
    irounds = [ir for ir, _ in series]
    assert max(irounds) <= rnd
    irounds += [rnd]
    ftwindows = np.diff(irounds)
    if max(ftwindows) == 0:
        ftprobs = [1.0]
    else:
        ftprobs = (ftwindows / np.sum(ftwindows)).tolist()
    ftvals = [val for _, val in series]
    assert len(ftprobs) == len(ftvals)
    running_average = 0.0
    for prob, val in zip(ftprobs, ftvals):
        running_average += prob * val
    return momentum * running_average + (1.0 - momentum) * ftvals[-1]


def ftf_caps(R, ftobj, G, J, T, D, r, rhomax):
    """Upper bound on rem_j implied by the FTF row (shockwave.py:573-597):
    D(r+T) + rem_j/share <= rhomax*FTobj_j   <=>   rem_j <= share*(rhomax*FTobj_j - D(r+T))."""
    share = min(1.0, G / J)
    return share * (rhomax * np.asarray(ftobj, dtype=np.float64) - D * (r + T))


def evaluate(x, g, E, c, dbar, R, w, G, T, D, k, bases, logv):
    """MILP objective of an integral (or fractional) x, with p_j at its best value
    p_j = min(D n_j / dbar_j, E_j - c_j).  Returns (objective, welfare, M, n, cap_ok)."""
    x = np.asarray(x, dtype=np.float64)
    J = x.shape[0]
    n = x.sum(axis=1)
    p = np.minimum(D * n / dbar, (E - c).astype(np.float64))
    u = (c + p) / E
    welfare = float(np.sum(w * plog(u, bases, logv)) / (J * T))
    rem = np.maximum(0.0, R - dbar * p)
    M = float(rem.max()) if J else 0.0
    cap_ok = bool(np.all(x.T @ g.astype(np.float64) <= G + 1e-9))
    return welfare - k * M, welfare, M, n, cap_ok


def _solve(g, E, c, dbar, R, w, G, T, D, k, bases, logv, ftf_cap, rel_gap, time_limit,
           relax=False, with_sos2=False):
    """One MILP (or its LP relaxation).  Returns (status_ok, x[J,T], p[J], objective)."""
    J = len(g)
    B = len(bases)
    nx, npv, nl = J * T, J, J * B
    nz = J * B if with_sos2 else 0
    oX, oP, oL, oZ = 0, nx, nx + npv, nx + npv + nl
    oM = oZ + nz
    nvar = oM + 1

    cost = np.zeros(nvar)
    # maximise sum_j w_j sum_b lambda_jb logv_b /(J T) - k M   ->  minimise the negation
    cost[oL:oL + nl] = -(np.repeat(w, B) * np.tile(logv, J)) / (J * T)
    cost[oM] = k

    # constraint blocks, built vectorised (COO triplets); row order: cap, prog, pwl_u, pwl_1, [sos2], mk, [ftf]
    R_, C_, V_, lo, hi = [], [], [], [], []
    nrow = 0
    jj = np.arange(J)
    gf = g.astype(float)
    Ef, cf = E.astype(float), c.astype(float)

    def block(r, cidx, v, lo_, hi_, n):
        nonlocal nrow
        R_.append(np.asarray(r) + nrow); C_.append(np.asarray(cidx)); V_.append(np.asarray(v, dtype=float))
        lo.append(np.broadcast_to(np.asarray(lo_, dtype=float), (n,))); hi.append(np.broadcast_to(np.asarray(hi_, dtype=float), (n,)))
        nrow += n

    # cap: sum_j g_j x_jt <= G                                  (row t)
    tt = np.arange(T)
    block(np.repeat(tt, J), oX + np.tile(jj * T, T) + np.repeat(tt, J), np.tile(gf, T), -np.inf, float(G), T)
    # prog: dbar_j p_j - D sum_t x_jt <= 0                        (row j)
    block(np.concatenate([jj, np.repeat(jj, T)]),
          np.concatenate([oP + jj, oX + np.arange(J * T)]),
          np.concatenate([dbar, np.full(J * T, -float(D))]), -np.inf, 0.0, J)
    # pwl_u: sum_b lambda_jb base_b - p_j/E_j = c_j/E_j           (row 2j)
    # pwl_1: sum_b lambda_jb = 1                                  (row 2j+1)
    # interleaved per job: the row order is part of what the end-to-end pin (tests/golden/tacc32_oracle_pin.json) was
    # recorded with — HiGHS' branch and bound returns a different, equally optimal-within-gap incumbent otherwise
    rhs = np.empty(2 * J)
    rhs[0::2] = cf / Ef
    rhs[1::2] = 1.0
    block(np.concatenate([np.repeat(2 * jj, B), 2 * jj, np.repeat(2 * jj + 1, B)]),
          np.concatenate([oL + np.arange(J * B), oP + jj, oL + np.arange(J * B)]),
          np.concatenate([np.tile(np.asarray(bases, dtype=float), J), -1.0 / Ef, np.ones(J * B)]), rhs, rhs, 2 * J)
    if with_sos2:
        for j in range(J):
            zb = oZ + j * B
            block([0] * B, list(range(zb, zb + B)), [1.0] * B, -np.inf, 2.0, 1)
            for b in range(B):
                block([0, 0], [oL + j * B + b, zb + b], [1.0, -1.0], -np.inf, 0.0, 1)
            for l in range(0, B - 2):
                for rr in range(l + 2, B):
                    block([0, 0], [zb + l, zb + rr], [1.0, 1.0], -np.inf, 1.0, 1)
    # mk: M + dbar_j p_j >= R_j  (M >= 0 through its bound: rem_j = max(0, .))
    block(np.concatenate([jj, jj]), np.concatenate([np.full(J, oM), oP + jj]),
          np.concatenate([np.ones(J), dbar]), R, np.inf, J)
    # ftf: rem_j <= cap_j  <=>  cap_j >= 0  and  dbar_j p_j >= R_j - cap_j
    if ftf_cap is not None:
        if np.any(ftf_cap < 0):
            return False, None, None, None
        block(jj, oP + jj, dbar, R - ftf_cap, np.inf, J)
    rows, cols, vals = np.concatenate(R_), np.concatenate(C_), np.concatenate(V_)
    lo, hi = np.concatenate(lo), np.concatenate(hi)

    A = sp.csr_matrix((vals, (rows, cols)), shape=(nrow, nvar))
    lb = np.zeros(nvar)
    ub = np.full(nvar, np.inf)
    ub[oX:oX + nx] = 1.0
    if with_sos2:
        ub[oZ:oZ + nz] = 1.0
    integrality = np.zeros(nvar)
    if not relax:
        integrality[oX:oX + nx] = 1
        if with_sos2:
            integrality[oZ:oZ + nz] = 1
    opts = {"mip_rel_gap": rel_gap, "disp": False}
    if time_limit and time_limit > 0:
        opts["time_limit"] = float(time_limit)
    res = milp(cost, constraints=LinearConstraint(A, np.asarray(lo), np.asarray(hi)),
               integrality=integrality, bounds=Bounds(lb, ub), options=opts)
    if res.x is None:
        return False, None, None, None
    x = res.x[oX:oX + nx].reshape(J, T)
    if not relax:
        x = np.round(x)
    return True, x, res.x[oP:oP + npv], -float(res.fun)


def relax_priorities(R, ftobj, G, J, D, r, rhomax, lam):
    """shockwave.py:830-911: ratio_j = (D r + R_j/share)/FTobj_j ; prio = ratio^lam if ratio>rhomax
    (ratio^100 when R_j < D) else 1."""
    share = min(1.0, G / J)
    ratio = (D * r + np.asarray(R, dtype=np.float64) / share) / np.asarray(ftobj, dtype=np.float64)
    prio = np.ones(J)
    for j in range(J):
        if ratio[j] > rhomax:
            prio[j] = ratio[j] ** lam
            if R[j] < D:
                prio[j] = ratio[j] ** 1e2
    return prio, ratio


def rank_objective(y, prio):
    """sum_j prio_j * mean round index of job j   (shockwave.py:761-779)."""
    y = np.asarray(y, dtype=np.float64)
    n = y.sum(axis=1)
    t = np.arange(y.shape[1], dtype=np.float64)
    tot = 0.0
    for j in range(y.shape[0]):
        if n[j] > 0:
            tot += prio[j] * float(y[j] @ t) / n[j]
    return tot


def rank_in_schedule(x, prio, g, G, rel_gap, time_limit):
    """shockwave.py:714-793: keep per-job round counts, re-order rounds to minimise
    sum_j prio_j * (sum_t t*y_jt)/n_j under the same per-round capacity."""
    J, T = x.shape
    n = x.sum(axis=1)
    if not np.any(n > 0):
        return x
    nvar = J * T
    cost = np.zeros(nvar)
    for j in range(J):
        if n[j] > 0:
            cost[j * T:(j + 1) * T] = prio[j] * np.arange(T) / n[j]
    rows, cols, vals, lo, hi = [], [], [], [], []
    for j in range(J):
        rows += [j] * T; cols += list(range(j * T, (j + 1) * T)); vals += [1.0] * T
        lo.append(float(n[j])); hi.append(float(n[j]))
    jj = np.arange(J)
    for t in range(T):
        rows += [J + t] * J; cols += (jj * T + t).tolist(); vals += g.astype(float).tolist()
        lo.append(-np.inf); hi.append(float(G))
    A = sp.csr_matrix((vals, (rows, cols)), shape=(J + T, nvar))
    opts = {"mip_rel_gap": rel_gap, "disp": False}
    if time_limit and time_limit > 0:
        opts["time_limit"] = float(time_limit)
    res = milp(cost, constraints=LinearConstraint(A, np.asarray(lo), np.asarray(hi)),
               integrality=np.ones(nvar), bounds=Bounds(np.zeros(nvar), np.ones(nvar)), options=opts)
    if res.x is None:
        return x
    return np.round(res.x.reshape(J, T))


def dynamic_eisenberg_gale(g, E, c, dbar, R, ftobj, G, T, D, r, k, lam, rhomax, bases, logv,
                           rel_gap=1e-3, time_limit=15.0, relax=False, with_sos2=False,
                           do_rank=True):
    """Whole solve incl. the infeasible -> relaxed -> re-rank fallback (shockwave.py:504-711).
    Returns dict(x, status, objective, weights, ftf_cap)."""
    g = np.asarray(g, dtype=np.int64); E = np.asarray(E, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64); dbar = np.asarray(dbar, dtype=np.float64)
    R = np.asarray(R, dtype=np.float64); ftobj = np.asarray(ftobj, dtype=np.float64)
    J = len(g)
    ones = np.ones(J)
    cap = ftf_caps(R, ftobj, G, J, T, D, r, rhomax)
    ok, x, p, obj = _solve(g, E, c, dbar, R, ones, G, T, D, k, bases, logv, cap, rel_gap,
                           time_limit, relax, with_sos2)
    if ok:
        return dict(x=x, status=STATUS_FTF_FEASIBLE, objective=obj, weights=ones, ftf_cap=cap, p=p)
    prio, _ = relax_priorities(R, ftobj, G, J, D, r, rhomax, lam)
    ok, x, p, obj = _solve(g, E, c, dbar, R, prio, G, T, D, k, bases, logv, None, rel_gap,
                           time_limit, relax, with_sos2)
    assert ok, "relaxed problem must have a solution (shockwave.py:671)"
    if do_rank and not relax:
        x = rank_in_schedule(x, prio, g, G, rel_gap, time_limit)
    return dict(x=x, status=STATUS_FALLBACK, objective=obj, weights=prio, ftf_cap=cap, p=p)


def construct_schedules(x, jobids, g, R, round_ptr, G):
    """shockwave.py:213-285: per round, the jobs with round(x)==1, then back-fill idle GPUs with the
    not-yet-scheduled jobs in DESCENDING remaining-runtime order that still fit."""
    J, T = x.shape
    sched = OrderedDict()
    order = sorted(range(J), key=lambda i: R[i], reverse=True)  # stable, like Python's sorted()
    for t in range(T):
        cur = [j for j in range(J) if round(float(x[j, t])) == 1.0]
        cur_set = set(cur)
        idle = G - int(sum(g[j] for j in cur))
        if idle > 0:
            for j in order:
                if j in cur_set:
                    continue
                if g[j] <= idle:
                    idle -= int(g[j])
                    cur.append(j)
                if idle <= 0:
                    break
        sched[round_ptr + t] = [jobids[j] for j in cur]
    return sched
