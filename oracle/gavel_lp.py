"""CPU ORACLE (test infrastructure, not product code) for the Gavel policies' get_allocation().

Restates on scipy's HiGHS (`linprog`) the programs the reference hands to cvxpy -> ECOS / Gurobi:
  base constraints      scheduler/policies/policy.py:58-65        x >= 0, sum_j sf_j x_jw <= N_w, sum_w x_jw <= 1
  MaxMinFairness_Perf   scheduler/policies/max_min_fairness.py:53-113
  FinishTimeFairness_Perf  scheduler/policies/finish_time_fairness.py:66-157  (min of a max of ratios; solved by
                        bisection on rho over LP feasibility — the cvxpy model is the same quasi-convex program)
  MinTotalDuration_Perf scheduler/policies/min_total_duration.py:55-135  (the reference's own bisection on T)
  ThroughputNormalizedByCostSum_PerfSLOs   scheduler/policies/max_sum_throughput.py:49-108
  Proportional / Isolated throughputs      proportional.py:26-43, isolated.py:35-55
Pinning: the reference has no known-answer test for these -> "parity unpinned" at the
x level; LP optima are degenerate so parity is on the objective and on constraints.
General in the number of worker types W.
"""
import numpy as np
from scipy.optimize import linprog


def proportional_throughputs(thr, N):
    m = thr.shape[0]
    x = np.tile(np.asarray(N, dtype=float) / m, (m, 1))
    x = x / x.sum(axis=1).max()
    return (thr * x).sum(axis=1)


def isolated_allocation(thr, sf, N):
    m = thr.shape[0]
    x = np.tile(np.asarray(N, dtype=float) / m, (m, 1)) / np.asarray(sf, dtype=float)[:, None]
    rs = np.maximum(x.sum(axis=1), 1.0)
    return x / rs[:, None]


def _base(J, W, sf, N):
    """A_ub, b_ub rows for the base constraints on x flattened row-major [J*W]."""
    A = np.zeros((W + J, J * W))
    for w in range(W):
        A[w, w::W] = sf
    for j in range(J):
        A[W + j, j * W:(j + 1) * W] = 1.0
    return A, np.concatenate([np.asarray(N, dtype=float), np.ones(J)])


def max_min(coef, sf, N):
    """max z  s.t.  sum_w coef_jw x_jw >= z  + base.  Returns (z, x[J,W])."""
    J, W = coef.shape
    A, b = _base(J, W, sf, N)
    A = np.hstack([A, np.zeros((A.shape[0], 1))])
    C = np.zeros((J, J * W + 1))
    for j in range(J):
        C[j, j * W:(j + 1) * W] = -coef[j]
        C[j, -1] = 1.0
    cost = np.zeros(J * W + 1); cost[-1] = -1.0
    res = linprog(cost, A_ub=np.vstack([A, C]), b_ub=np.concatenate([b, np.zeros(J)]),
                  bounds=[(0, None)] * (J * W) + [(None, None)], method="highs")
    assert res.status == 0, res.message
    return float(res.x[-1]), res.x[:-1].reshape(J, W)


def max_min_fairness_perf(thr, sf, priority, N):
    pw = (1.0 / np.asarray(priority, dtype=float)) / proportional_throughputs(thr, N)
    coef = thr * pw[:, None] * np.asarray(sf, dtype=float)[:, None]
    return max_min(coef, sf, N)


def feasible_rates(thr, sf, N, need):
    """LP feasibility of sum_w thr_jw x_jw >= need_j + base; returns x or None."""
    J, W = thr.shape
    A, b = _base(J, W, sf, N)
    C = np.zeros((J, J * W))
    for j in range(J):
        C[j, j * W:(j + 1) * W] = -thr[j]
    res = linprog(np.zeros(J * W), A_ub=np.vstack([A, C]), b_ub=np.concatenate([b, -np.asarray(need, float)]),
                  bounds=[(0, None)] * (J * W), method="highs")
    return res.x.reshape(J, W) if res.status == 0 else None


def finish_time_fairness_perf(thr, sf, t, n, cum_iso, N, tol=1e-10):
    iso = (thr * isolated_allocation(thr, sf, N)).sum(axis=1)
    den = np.asarray(cum_iso, float) + np.asarray(n, float) / iso
    lo, hi = 0.0, 1.0
    def need(rho):
        room = rho * den - t
        return None if np.any(room <= 0) else n / room
    while True:
        nd = need(hi)
        if nd is not None and feasible_rates(thr, sf, N, nd) is not None:
            break
        lo, hi = hi, hi * 2.0
    while hi - lo > tol * hi:
        mid = 0.5 * (lo + hi)
        nd = need(mid)
        if nd is not None and feasible_rates(thr, sf, N, nd) is not None:
            hi = mid
        else:
            lo = mid
    return hi, feasible_rates(thr, sf, N, need(hi)), den


def min_total_duration_perf(thr, sf, n, N):
    max_T, min_T, last_max_T, best, bx = 1000000.0, 100.0, 1000000.0, None, None
    while bx is None:
        while 1.05 * min_T < max_T:
            T = (min_T + max_T) / 2.0
            x = feasible_rates(thr, sf, N, np.asarray(n, float) / T)
            if x is not None:
                best, bx, max_T = T, x, T
            else:
                min_T = T
        max_T, min_T, last_max_T = last_max_T * 10.0, last_max_T, last_max_T * 10.0
    return best, bx


def max_sum_throughput(thr, sf, N, costs=None, need=None):
    """max sum_jw thr_jw x_jw / cost_w + base  [+ SLO rows sum_w thr_jw x_jw >= need_j, max_sum_throughput.py:87-93].
    Returns (value, x), or (None, None) when the SLO rows make it infeasible (the reference then re-solves without them)."""
    J, W = thr.shape
    c = np.ones(W) if costs is None else np.asarray(costs, float)
    A, b = _base(J, W, sf, N)
    if need is not None:
        C = np.zeros((J, J * W))
        for j in range(J):
            C[j, j * W:(j + 1) * W] = -thr[j]
        A, b = np.vstack([A, C]), np.concatenate([b, -np.asarray(need, float)])
    res = linprog(-(thr / c[None, :]).reshape(-1), A_ub=A, b_ub=b, bounds=[(0, None)] * (J * W), method="highs")
    if need is not None and res.status != 0:
        return None, None
    assert res.status == 0
    return -float(res.fun), res.x.reshape(J, W)


# ---- which optimal x?  (interior-point selection) ------------------------------------------------------
# The LPs above are degenerate in x.  The reference hands them to interior-point solvers
# (ECOS for the LPs, Gurobi barrier for the finish-time-fairness cone program; utils.py:603-685), whose
# iterates converge to the ANALYTIC CENTRE of the optimal face: the point maximising the sum of the logs of
# the slacks of the constraints that are not active on the whole face.  A simplex vertex (what linprog/HiGHS
# returns) has the same objective but drives the closed loop differently: on the canonical 120-job trace the
# vertex choice moves avg JCT by 6 % for finish_time_fairness / min_total_duration, the analytic centre
# reproduces the golden pickles within 0.8 % (tests/golden/tacc32_policy_pins.json).  Pooled (one live worker
# type) form, x_j scalar:
#     maximise  sum_j [ w_lo log(x_j - lo_j) + w_x log(x_j) + log(1 - x_j) ] + log(N - sum_j sf_j x_j)
# with lo_j the job's requirement at the optimal scalar.  Weights: LP rows (max-min, min-total-duration)
# w_lo = w_x = 1; finish-time fairness w_lo = 2, w_x = 0 (cvxpy states t_j + n_j inv_pos(thr_j x_j) <= rho den_j
# with an auxiliary u_j >= 1/(thr_j x_j) as a rotated second-order cone; maximising log(u y - 1) + log(U - u)
# over u leaves 2 log(x - lo) - log(x), and the -log(x) cancels the x >= 0 row).
def analytic_centre_box(lo, sf, N, w_lo=1.0, w_x=1.0):
    lo = np.minimum(np.asarray(lo, dtype=float), 1.0)
    sf = np.asarray(sf, dtype=float)
    fixed = lo >= 1.0 - 1e-15
    x = np.where(fixed, 1.0, np.maximum(lo, 0.0))
    if N - float((sf * x).sum()) <= 1e-12 * N:
        return x                                      # capacity is active on the whole face: x = lo is the only point
    free = ~fixed
    l0, s = np.maximum(lo[free], 0.0), sf[free]
    Nf = N - float(sf[fixed].sum())

    def xj(lam):
        l, h = l0.copy(), np.ones(len(l0))
        for _ in range(100):
            m = 0.5 * (l + h)
            f = w_x / m - 1.0 / (1.0 - m) + w_lo / np.maximum(m - l0, 1e-300) - lam * s
            l = np.where(f > 0, m, l)
            h = np.where(f > 0, h, m)
        return 0.5 * (l + h)
    a, b = 1e-12, 1e15
    for _ in range(200):
        lam = np.sqrt(a * b)
        sl = Nf - float((s * xj(lam)).sum())
        if sl > 0 and lam * sl > 1.0:
            b = lam
        else:
            a = lam
    x[free] = xj(np.sqrt(a * b))
    return x


def analytic_centre_tied(sf, C, lo=None):
    """max sum_j [ 1{lo_j>0} log(x_j - lo_j) + log x_j + log(1 - x_j) ]  s.t.  sum_j sf_j x_j = C
    (the tied group of a fractional knapsack; lo_j = the job's SLO floor, 0 without one)."""
    sf = np.asarray(sf, dtype=float)
    lo = np.zeros(len(sf)) if lo is None else np.asarray(lo, dtype=float)
    w = (lo > 0).astype(float)

    def xj(nu):
        l, h = lo.copy(), np.ones(len(sf))
        for _ in range(100):
            m = 0.5 * (l + h)
            f = 1.0 / m - 1.0 / (1.0 - m) + w / np.maximum(m - lo, 1e-300) - nu * sf
            l = np.where(f > 0, m, l)
            h = np.where(f > 0, h, m)
        return 0.5 * (l + h)
    a, b = -1e12, 1e12
    for _ in range(300):
        nu = 0.5 * (a + b)
        if float((sf * xj(nu)).sum()) > C:
            a = nu
        else:
            b = nu
    return xj(0.5 * (a + b))


def max_sum_pooled_centre(v, sf, N, lo=None):
    """Pooled max-sum (fractional knapsack by v_j/sf_j, optional SLO floors lo_j <= x_j) with the interior-point
    selection inside the tied group.  Returns (x, value), or (None, None) if the floors alone do not fit."""
    v, sf = np.asarray(v, float), np.asarray(sf, float)
    lo = np.zeros(len(v)) if lo is None else np.asarray(lo, float)
    if np.any(lo > 1.0 + 1e-12) or float((sf * lo).sum()) > N * (1.0 + 1e-12):
        return None, None
    ratio = v / sf
    order = np.argsort(-ratio, kind="stable")
    x = np.minimum(lo, 1.0).copy()
    used, i = float((sf * x).sum()), 0
    while i < len(order):
        i0, r = i, ratio[order[i]]
        while i < len(order) and abs(ratio[order[i]] - r) <= 1e-12 * abs(r):
            i += 1
        grp = order[i0:i]
        need = float((sf[grp] * (1.0 - x[grp])).sum())
        if not r > 0.0:
            break
        if used + need <= N:
            x[grp] = 1.0
            used += need
        else:
            if N - used > 1e-12 * N:
                x[grp] = analytic_centre_tied(sf[grp], (N - used) + float((sf[grp] * x[grp]).sum()), lo[grp])
            break
    return x, float((v * x).sum())


# ---- AlloX (scheduler/policies/allox.py:19-188) -------------------------------------------------------
def allox_q_matrix(p, t, wtype):
    """q[i][k*n + j] = (k+1) * p[i][wtype[j]] + t[i]   (allox.py:108-138), p = steps_remaining / throughput."""
    p, t, wtype = np.asarray(p, float), np.asarray(t, float), np.asarray(wtype, int)
    m, n = p.shape[0], len(wtype)
    q_base = p[:, wtype]                                   # [m, n]
    q = np.concatenate([(k + 1) * q_base for k in range(m)], axis=1)
    return q + t[:, None]


def allox_assignment(p, t, wtype):
    """The reference's solver call (allox.py:144): scipy's Jonker-Volgenant linear_sum_assignment."""
    from scipy.optimize import linear_sum_assignment
    q = allox_q_matrix(p, t, wtype)
    rows, cols = linear_sum_assignment(q)
    return cols[np.argsort(rows)], float(q[rows, cols].sum())


def allox_allocation(job_ids, worker_types, unflattened_throughputs, scale_factors, times_since_start,
                     num_steps_remaining, cluster_spec, prev_allocation, alpha, assign=allox_assignment):
    """Restatement of AlloXPolicy.get_allocation (allox.py:46-188) around a pluggable assignment solver."""
    unalloc, already = [], []
    for job_id in unflattened_throughputs:
        if job_id not in prev_allocation:
            unalloc.append(job_id)
        else:
            tot = 0.0
            for w in worker_types:
                tot += prev_allocation[job_id][w]
            (already if tot == 1.0 else unalloc).append(job_id)
    m = len(unalloc)
    n = 0
    w_of = {}
    for w in worker_types:
        num = cluster_spec[w]
        for j in already:
            if prev_allocation[j][w] == 1.0:
                num -= 1
        for wid in range(n, n + num):
            w_of[wid] = w
            n += 1
    unalloc.sort(key=lambda x: -times_since_start[x])
    unalloc = unalloc[: max(int(alpha * m), n)]
    m = len(unalloc)
    allocation = {j: {w: 0.0 for w in cluster_spec} for j in job_ids}
    for j in job_ids:
        if j in prev_allocation:
            allocation[j] = dict(prev_allocation[j])
    if m > 0 and n > 0:
        wt_index = {w: i for i, w in enumerate(worker_types)}
        p = np.zeros((m, len(worker_types)))
        for i, j in enumerate(unalloc):
            for w in worker_types:
                thr = unflattened_throughputs[j][w]
                p[i, wt_index[w]] = num_steps_remaining[j] / (thr if thr != 0.0 else 1e-10)
        t = np.array([times_since_start[j] for j in unalloc], float)
        wtype = np.array([wt_index[w_of[wid]] for wid in range(n)])
        cols, _ = assign(p, t, wtype)
        per_worker = {i: [] for i in range(n)}
        for row, col in enumerate(cols):
            per_worker[int(col) % n].append((unalloc[row], int(col) // n))
        for wid in range(n):
            lst = [(x[0], len(per_worker[wid]) - 1 - x[1]) for x in per_worker[wid]]
            lst.sort(key=lambda x: x[1])
            if lst:
                allocation[lst[0][0]][w_of[wid]] = 1.0 / scale_factors[lst[0][0]]
    return allocation


def eisenberg_gale(coef, sf, N):
    """The cvxpy program of policies/max_min_fairness_strategy_proof.py:102-123 — maximise geo_mean_j(sum_w coef_jw x_jw)
    s.t. x >= 0, sum_j sf_j x_jw <= N_w, sum_w x_jw <= 1 (policy.py:58-65) — restated as max sum_j log(.) (same argmax)
    on scipy's trust-constr (cvxpy is not installable here).  Per-job utilities are unique at the optimum (strictly
    concave in them); x need not be.  Returns (x [J][W], utilities [J])."""
    import numpy as np
    from scipy.optimize import Bounds, LinearConstraint, minimize
    coef, sf, N = np.asarray(coef, float), np.asarray(sf, float), np.asarray(N, float)
    J, W = coef.shape
    # utilities floored at 1e-9 inside the solver only (finite derivatives when a trial point starves a job); exact
    # block-diagonal Hessian and a start that already uses 90 % of every capacity — with the quasi-Newton default and a
    # timid start trust-constr stops far from the optimum on some instances (capacities 15-60 % used)
    U = lambda z: np.maximum((coef * z.reshape(J, W)).sum(axis=1), 1e-9)
    f = lambda z: -np.log(U(z)).sum()
    g = lambda z: -(coef / U(z)[:, None]).ravel()

    def h(z):
        u = U(z)
        H = np.zeros((J * W, J * W))
        for j in range(J):
            H[j * W:(j + 1) * W, j * W:(j + 1) * W] = np.outer(coef[j], coef[j]) / u[j] ** 2
        return H
    A1 = np.kron(np.eye(J), np.ones((1, W)))
    A2 = np.kron(sf[None, :], np.eye(W)).reshape(W, J * W)
    x0 = np.tile(np.minimum(1.0 / W, N / sf.sum()) * 0.9, J)
    r = minimize(f, x0, jac=g, hess=h, method="trust-constr", bounds=Bounds(0.0, 1.0),
                 constraints=[LinearConstraint(A1, -np.inf, 1.0), LinearConstraint(A2, -np.inf, N)],
                 options=dict(gtol=1e-10, xtol=1e-12, maxiter=5000, barrier_tol=1e-12))
    if r.status not in (1, 2):
        raise RuntimeError(f"trust-constr: {r.message}")
    # polish: the interior-point iterate stays ~1e-3 inside constraints that are exactly tight at the optimum (a job at
    # its full time share); an active-set pass from there lands on them
    cons = [dict(type="ineq", fun=lambda z: 1.0 - A1 @ z, jac=lambda z: -A1),
            dict(type="ineq", fun=lambda z: N - A2 @ z, jac=lambda z: -A2)]
    r2 = minimize(f, np.clip(r.x, 0.0, 1.0), jac=g, bounds=[(0.0, 1.0)] * (J * W), constraints=cons, method="SLSQP",
                  options=dict(maxiter=500, ftol=1e-15))
    if np.isfinite(r2.fun) and r2.fun <= r.fun + 1e-12 and (A1 @ r2.x).max() <= 1 + 1e-9 and np.all(A2 @ r2.x <= N * (1 + 1e-9)):
        r = r2
    x = np.clip(r.x.reshape(J, W), 0.0, 1.0)
    return x, (coef * x).sum(axis=1)


def eisenberg_gale_batch(N, coef, sf, present):
    """present [S][J]: one program per scenario over the jobs present in it; x [S][J][W] (absent jobs: 0)."""
    import numpy as np
    present = np.asarray(present, bool)
    S, J = present.shape
    x = np.zeros((S, J, len(N)))
    for s in range(S):
        idx = np.flatnonzero(present[s])
        if len(idx):
            x[s, idx] = eisenberg_gale(np.asarray(coef)[idx], np.asarray(sf)[idx], N)[0]
    return x
