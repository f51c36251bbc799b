"""CPU ORACLE (test infrastructure, not product code) for the packing policies.

Restates, on scipy's HiGHS, the programs the reference's *WithPacking classes hand to cvxpy — dense matrices and
Python loops exactly as the reference builds them, so that shockwave_b200/packing.py (sparse, vectorised, solved by
swb_lp_solve on the GPU) has an independent yardstick:
  flatten / scale_factors_array / base constraints ..... scheduler/policies/policy.py:68-193
  MaxMinFairnessPolicyWithPacking.get_allocation ....... scheduler/policies/max_min_fairness.py:317-410
  FinishTimeFairnessPolicyWithPacking.get_allocation ... scheduler/policies/finish_time_fairness.py:160-290
  MinTotalDurationPolicyWithPacking .................... scheduler/policies/min_total_duration.py:138-234
  ThroughputNormalizedByCostSumWithPackingSLOs ......... scheduler/policies/max_sum_throughput.py:111-200
  proportional / isolated throughputs .................. proportional.py:14-43, isolated.py:14-53
cvxpy, ECOS and Gurobi are not installed here, so the LP optima are HiGHS's; the array construction
(flatten, scale_factors_array) is pinned bit for bit on the reference's own `policy.py`
(tests/test_oracle_packed.py imports it with a stub `cvxpy`).  Only tests/ may import this module.
"""
import numpy as np
from scipy.optimize import linprog


def flatten(d, cluster_spec, priority_weights=None):
    """policy.py:90-160.  Returns (all_m [n_single, n_comb, W] float32, (job_ids, single_ids, worker_types, relevant))."""
    job_ids = sorted(list(d.keys()))
    if len(job_ids) == 0:
        return None, None
    worker_types = sorted(list(d[job_ids[0]].keys()))
    relevant, single_set, singles = {}, set(), []
    for i, job_id in enumerate(job_ids):
        if not job_id.is_pair():
            single_set.add(job_id)
            singles.append(job_id)
            relevant.setdefault(job_id, []).append(i)
        else:
            for s in job_id.singletons():
                relevant.setdefault(s, []).append(i)
    if len(worker_types) == 0:
        return None, None
    all_m = np.zeros((len(single_set), len(job_ids), len(worker_types)), dtype=np.float32)
    for i, s in enumerate(singles):
        for j in relevant[s]:
            job_id = job_ids[j]
            for k, w in enumerate(worker_types):
                if job_id in single_set:
                    if job_id == s:
                        all_m[i][j][k] = d[job_id][w]
                elif s.overlaps_with(job_id):
                    all_m[i][j][k] = d[job_id][w][job_id.as_tuple().index(s[0])]
        if priority_weights is not None:
            all_m[i] /= priority_weights[s]
    return all_m, (job_ids, singles, worker_types, relevant)


def scale_factors_array(scale_factors, job_ids, m, n):
    """policy.py:72-88: the common scale factor of a combination, 0 when its members disagree."""
    out = np.zeros((m, n))
    for i in range(m):
        sf = None
        for s in job_ids[i].singletons():
            if sf is not None and sf != scale_factors[s]:
                sf = 0
            else:
                sf = scale_factors[s]
        out[i, :] = sf
    return out


def proportional_throughputs(thr, N):
    """proportional.py:14-43."""
    m = thr.shape[0]
    x = np.tile(np.asarray(N, float) / m, (m, 1))
    x = x / x.sum(axis=1).max()
    return (thr * x).sum(axis=1)


def isolated_throughputs(thr, sf, N):
    """isolated.py:14-53."""
    m = thr.shape[0]
    x = np.tile(np.asarray(N, float) / m, (m, 1)) / np.asarray(sf, float)[:, None]
    rs = np.maximum(x.sum(axis=1), 1.0)
    x = x / rs[:, None]
    return (thr * x).sum(axis=1)


class Packed:
    """The arrays every packed program starts from."""
    def __init__(self, thr, scale_factors, cluster_spec, priority_weights=None):
        self.all_m, self.index = flatten(thr, cluster_spec, priority_weights)
        self.job_ids, self.singles, self.worker_types, self.relevant = self.index
        self.N = np.array([cluster_spec[w] for w in self.worker_types], dtype=float)
        self.m, self.n = self.all_m[0].shape
        self.sfa = scale_factors_array(scale_factors, self.job_ids, self.m, self.n)
        self.thr_single = np.array([[thr[s][w] for w in self.worker_types] for s in self.singles], dtype=float)
        self.sf_single = np.array([scale_factors[s] for s in self.singles], dtype=float)

    def base(self):
        """policy.py:172-193 + the explicit x == 0 where the effective scale factor is 0: (A_ub, b_ub, bounds)."""
        m, n, nv = self.m, self.n, self.m * self.n
        A = np.zeros((n + len(self.singles), nv))
        b = np.zeros(n + len(self.singles))
        for w in range(n):
            for c in range(m):
                A[w, c * n + w] = self.sfa[c, w]
            b[w] = self.N[w]
        for i, s in enumerate(self.singles):
            for c in self.relevant[s]:
                A[n + i, c * n:(c + 1) * n] = 1.0
            b[n + i] = 1.0
        bounds = [(0.0, 0.0 if self.sfa[c, w] == 0 else None) for c in range(m) for w in range(n)]
        return A, b, bounds

    def coef(self, i, with_sf=False):
        """Row vector with T_i(x) = coef . x: sum over the relevant combinations of all_m[i] (* scale factor)."""
        v = np.zeros(self.m * self.n)
        for c in self.relevant[self.singles[i]]:
            row = self.all_m[i][c].astype(float)
            if with_sf:
                row = row * self.sfa[c]
            v[c * self.n:(c + 1) * self.n] = row
        return v

    def max_min(self, rows):
        """max z : z <= rows_i . x for all i, base constraints.  Returns (z, x [m, n])."""
        A, b, bounds = self.base()
        nv = self.m * self.n
        Az = np.hstack([-np.asarray(rows), np.ones((len(rows), 1))])
        A2 = np.vstack([np.hstack([A, np.zeros((A.shape[0], 1))]), Az])
        b2 = np.concatenate([b, np.zeros(len(rows))])
        c = np.zeros(nv + 1)
        c[-1] = -1.0
        r = linprog(c, A_ub=A2, b_ub=b2, bounds=bounds + [(None, None)], method="highs")
        assert r.status == 0, r.message
        return -r.fun, r.x[:nv].reshape(self.m, self.n)

    def feasible(self, rows, need):
        """Any x with rows_i . x >= need_i under the base constraints, or None."""
        A, b, bounds = self.base()
        A2 = np.vstack([A, -np.asarray(rows)])
        b2 = np.concatenate([b, -np.asarray(need, float)])
        r = linprog(np.zeros(A.shape[1]), A_ub=A2, b_ub=b2, bounds=bounds, method="highs")
        return r.x.reshape(self.m, self.n) if r.status == 0 else None


def max_min_fairness_packed(thr, scale_factors, priority_weights, cluster_spec):
    """max_min_fairness.py:317-410.  Returns (objective, x [m, n], Packed)."""
    P = Packed(thr, scale_factors, cluster_spec, priority_weights)
    prop = proportional_throughputs(P.thr_single, P.N)
    rows = [P.coef(i, with_sf=True) / prop[i] for i in range(len(P.singles))]
    z, x = P.max_min(rows)
    return z, x, P


def finish_time_fairness_packed(thr, scale_factors, priority_weights, times_since_start, num_steps_remaining,
                                cumulative_isolated_time, cluster_spec, tol=1e-10):
    """finish_time_fairness.py:169-290 with the cumulative isolated times given (the stateful part is host logic):
    minimise max_i (t_i + n_i / T_i(x)) / (cum_i + n_i / iso_i) by bisection on the ratio over LP feasibility.
    Returns (rho, x, Packed)."""
    P = Packed(thr, scale_factors, cluster_spec, priority_weights)
    iso = isolated_throughputs(P.thr_single, P.sf_single, P.N)
    n = np.array([num_steps_remaining[s] for s in P.singles], dtype=float)
    t = np.array([times_since_start[s] for s in P.singles], dtype=float)
    den = np.array([cumulative_isolated_time[s] for s in P.singles], dtype=float) + n / iso
    rows = [P.coef(i) for i in range(len(P.singles))]

    def attempt(rho):
        room = rho * den - t
        if np.any(room <= 0):
            return None
        return P.feasible(rows, n / room)
    lo, hi = 0.0, 1.0
    while attempt(hi) is None:
        lo, hi = hi, hi * 2.0
        assert hi < 1e12
    while hi - lo > tol * hi:
        mid = 0.5 * (lo + hi)
        if attempt(mid) is not None:
            hi = mid
        else:
            lo = mid
    return hi, attempt(hi), P


def min_total_duration_packed(thr, scale_factors, num_steps_remaining, cluster_spec):
    """min_total_duration.py:176-234: the reference's own bisection on T (5 %), every probe an LP feasibility problem.
    Returns (T of the last feasible probe, x, Packed)."""
    P = Packed(thr, scale_factors, cluster_spec)
    n = np.array([num_steps_remaining[s] for s in P.singles], dtype=float)
    rows = [P.coef(i) for i in range(len(P.singles))]
    max_T, min_T, last_max_T = 1000000.0, 100.0, 1000000.0
    last, last_T = None, None
    while last is None:
        while 1.05 * min_T < max_T:
            T = (min_T + max_T) / 2.0
            x = P.feasible(rows, n / T)
            if x is not None:
                last, last_T, max_T = x, T, T
            else:
                min_T = T
        max_T, min_T = last_max_T * 10.0, last_max_T
        last_max_T *= 10
        assert last_max_T < 1e30
    return last_T, last, P


def max_sum_throughput_packed_slos(thr, scale_factors, cluster_spec, instance_costs=None, SLOs={},
                                   num_steps_remaining={}):
    """max_sum_throughput.py:116-200.  Returns (objective, x, used_SLO_rows, Packed)."""
    P = Packed(thr, scale_factors, cluster_spec)
    cost = np.ones(P.n) if instance_costs is None else np.array([instance_costs[w] for w in P.worker_types], float)
    obj = np.zeros(P.m * P.n)
    for i in range(len(P.singles)):
        for c in P.relevant[P.singles[i]]:
            obj[c * P.n:(c + 1) * P.n] += P.all_m[i][c].astype(float) / cost
    A, b, bounds = P.base()
    rows, need = [], []
    for job_id in SLOs:
        i = P.job_ids.index(job_id)
        rows.append(P.coef(i))
        need.append(num_steps_remaining[job_id] / SLOs[job_id])
    used = False
    r = None
    if rows:
        r = linprog(-obj, A_ub=np.vstack([A, -np.asarray(rows)]), b_ub=np.concatenate([b, -np.asarray(need)]),
                    bounds=bounds, method="highs")
        used = r.status == 0
    if not used:
        r = linprog(-obj, A_ub=A, b_ub=b, bounds=bounds, method="highs")
        assert r.status == 0
    return -r.fun, r.x.reshape(P.m, P.n), used, P


def water_filling_packed(thr, scale_factors, priority_weights, cluster_spec, **kw):
    """MaxMinFairnessWaterFillingPolicyWithPacking.get_allocation (max_min_fairness_water_filling.py:569-718): the
    iteration loop of oracle/gavel_waterfill.run_iterations around the two programs restated over the packed columns
    (LP :81-189 on HiGHS, bottleneck MILP :191-305 on scipy.optimize.milp).
    Returns (x [m, n] clipped, normalised effective throughputs, iterations, log, Packed)."""
    from scipy.optimize import Bounds, LinearConstraint, milp
    from oracle import gavel_waterfill as wf
    P = Packed(thr, scale_factors, cluster_spec)
    Ns, nv = len(P.singles), P.m * P.n
    prop = proportional_throughputs(P.thr_single, P.N)
    net = np.array([P.coef(i) / prop[i] for i in range(Ns)])                     # net_i(x) = net[i] . x
    M = max(np.max(P.coef(i, with_sf=True)) / prop[i] for i in range(Ns))        # _get_M :583-602
    A, b, bounds = P.base()

    def lp(_thr, _sf, _N, _prop, lower, mult, add, so_far):
        rows = [np.concatenate([-net[i] * mult[i], [1.0]]) for i in range(Ns)]   # c <= (net_i - so_far_i) mult_i + add_i
        rhs = [add[i] - so_far[i] * mult[i] for i in range(Ns)]
        rows += [np.concatenate([-net[i], [0.0]]) for i in range(Ns)]            # net_i >= lower_i
        rhs += [-lower[i] for i in range(Ns)]
        cost = np.zeros(nv + 1); cost[-1] = -1.0
        r = linprog(cost, A_ub=np.vstack([np.hstack([A, np.zeros((A.shape[0], 1))]), np.array(rows)]),
                    b_ub=np.concatenate([b, rhs]), bounds=bounds + [(None, None)], method="highs")
        if r.status != 0:
            return None, None
        return r.x[:nv].reshape(P.m, P.n), float(r.x[-1])

    def bottleneck(_thr, _sf, _N, _prop, lower, so_far, zmask, M_, slack=wf.SLACK, epsilon=wf.EPSILON):
        cons = [LinearConstraint(np.hstack([A, np.zeros((A.shape[0], Ns))]), -np.inf, b)]
        Z = np.eye(Ns)
        cons.append(LinearConstraint(np.hstack([net, -M_ * Z]), -np.inf, so_far * slack - epsilon))
        cons.append(LinearConstraint(np.hstack([-net, M_ * Z]), -np.inf, M_ - so_far * slack))
        cons.append(LinearConstraint(np.hstack([net, 0 * Z]), lower, np.inf))
        ub = np.concatenate([[np.inf if hi is None else hi for _, hi in bounds], np.where(zmask > 0, 0.0, 1.0)])
        res = milp(np.concatenate([np.zeros(nv), -np.ones(Ns)]), constraints=cons,
                   integrality=np.concatenate([np.zeros(nv), np.ones(Ns)]), bounds=Bounds(np.zeros(nv + Ns), ub))
        if res.status != 0:
            raise RuntimeError("non-optimal allocation in _get_bottleneck_jobs")
        return np.round(res.x[nv:])

    log = []
    x, so_far, final, it = wf.run_iterations(P.singles, None, P.sf_single, P.N, prop,
                                             {s: priority_weights[s] for s in P.singles}, M, lp=lp,
                                             bottleneck=bottleneck, log=log, **kw)
    eff = net @ x.ravel()
    return np.clip(x, 0.0, 1.0), eff, it, log, P


# ---- CPU backend for shockwave_b200/packing.py's `_lp` hook (same signature as Engine.lp_solve) ----
def lp_backend(colp, rowi, val, c, b, max_iter=0):
    """HiGHS stand-in for swb_lp_solve: lets the product's HOST logic (LP construction, multi-section, stateful
    bookkeeping) be tested without a GPU."""
    import scipy.sparse as sp
    val = np.atleast_2d(val); c = np.atleast_2d(c); b = np.atleast_2d(b)
    S, n = c.shape
    m = b.shape[1]
    x = np.zeros((S, n)); obj = np.zeros(S); status = np.zeros(S, dtype=np.int32)
    for s in range(S):
        A = sp.csc_matrix((val[s], rowi, colp), shape=(m, n))
        r = linprog(-c[s], A_ub=A, b_ub=b[s], bounds=(0, None), method="highs")
        status[s] = {0: 0, 2: 1, 3: 2}.get(r.status, 4)
        if r.status == 0:
            x[s], obj[s] = r.x, -r.fun
    return x, obj, status, np.zeros((S, 4), dtype=np.int32)


# ---- the job-TYPE formulation (max_min_fairness.py:122-316) ----
def convert_job_type_allocation(allocation, job_id_to_job_type_key, make_pair):
    """policy.py:195-260: job x job-type allocation -> job x job allocation (x_ij = x_i,type(j) x_j,type(i) / sum)."""
    job_ids = sorted(allocation.keys())
    worker_types = sorted(allocation[job_ids[0]].keys())
    keys = sorted(set(job_id_to_job_type_key[j] for j in job_ids))
    jt = {w: {k: {o: 0.0 for o in [None] + keys} for k in keys} for w in worker_types}
    for w in worker_types:
        for j in allocation:
            k = job_id_to_job_type_key[j]
            for o in allocation[j][w]:
                jt[w][k][o] += allocation[j][w][o]
    out = {}
    for i, j in enumerate(job_ids):
        out[j] = {}
        k = job_id_to_job_type_key[j]
        for w in worker_types:
            out[j][w] = allocation[j][w][None]
        for j2 in job_ids[i + 1:]:
            k2 = job_id_to_job_type_key[j2]
            merged = make_pair(j[0], j2[0])
            out[merged] = {}
            for w in worker_types:
                cur = jt[w][k][k2]
                if cur > 0.0:
                    if k == k2:
                        cur -= allocation[j][w][k]
                    out[merged][w] = allocation[j][w][k2] * allocation[j2][w][k] / cur
                else:
                    out[merged][w] = 0.0
    return out


def max_min_fairness_job_types(thr, job_id_to_job_type_key, scale_factors, priority_weights, cluster_spec):
    """MaxMinFairnessPolicyWithPacking.get_allocation_using_job_type_throughputs (max_min_fairness.py:122-316) before
    the conversion: returns (objective, x [n, (1 + a) m], (job_ids, job_type_keys, worker_types))."""
    job_ids = sorted(job_id_to_job_type_key.keys())
    keys = sorted(thr.keys())
    wts = sorted(cluster_spec.keys())
    N = [cluster_spec[w] for w in wts]
    members = {}
    for i, j in enumerate(job_ids):
        members.setdefault(job_id_to_job_type_key[j], []).append(i)
    n, a, m = len(job_ids), len(keys), len(wts)
    nv1 = 1 + a
    nx = n * nv1 * m
    idx = lambda i, k, j: i * nv1 * m + k * nv1 + j
    sf = np.array([scale_factors[j] for j in job_ids], dtype=float)
    flat = np.zeros((a, nv1 * m), dtype=np.float32)
    for i, key in enumerate(keys):
        for k, w in enumerate(wts):
            for j, other in enumerate([None] + keys):
                flat[i, k * nv1 + j] = 0.0 if (j > 0 and other[1] != key[1]) else thr[key][w][other]
    A_ub, b_ub, A_eq, b_eq = [], [], [], []
    for i in range(n):                                  # sum of allocation values of a job <= 1
        r = np.zeros(nx); r[i * nv1 * m:(i + 1) * nv1 * m] = 1.0
        A_ub.append(r); b_ub.append(1.0)
    for k in range(m):                                  # capacity per worker type, pairs counted half
        r = np.zeros(nx)
        for i in range(n):
            for j in range(nv1):
                r[idx(i, k, j)] = sf[i] * (1.0 if j == 0 else 0.5)
        A_ub.append(r); b_ub.append(N[k])
    for i, k0 in enumerate(keys):                       # type a with type b == type b with type a
        for j, k1 in enumerate(keys):
            if j <= i or k0[1] != k1[1]:
                continue
            for k in range(m):
                r = np.zeros(nx)
                for ji in members.get(k0, []):
                    r[idx(ji, k, 1 + j)] += 1.0
                for ji in members.get(k1, []):
                    r[idx(ji, k, 1 + i)] -= 1.0
                A_eq.append(r); b_eq.append(0.0)
    for i, key in enumerate(keys):                      # i-A variables of the jobs of type A all equal
        for k in range(m):
            mem = members.get(key, [])
            for ja, jb in zip(mem[:-1], mem[1:]):
                r = np.zeros(nx); r[idx(ja, k, 1 + i)] = 1.0; r[idx(jb, k, 1 + i)] = -1.0
                A_eq.append(r); b_eq.append(0.0)
    alone = np.array([[thr[job_id_to_job_type_key[j]][w][None] for w in wts] for j in job_ids], dtype=float)
    prop = proportional_throughputs(alone, N)
    coef = np.zeros((n, nx))
    ub = np.full(nx, np.inf)
    for i, j in enumerate(job_ids):
        key = job_id_to_job_type_key[j]
        ti = keys.index(key)
        if len(members[key]) == 1:
            for k in range(m):
                ub[idx(i, k, 1 + ti)] = 0.0
        coef[i, i * nv1 * m:(i + 1) * nv1 * m] = flat[ti] * sf[i] / (priority_weights[j] * prop[i])
    Az = np.hstack([-coef, np.ones((n, 1))])
    pad = lambda M_: np.hstack([np.asarray(M_).reshape(-1, nx), np.zeros((len(M_), 1))])
    cost = np.zeros(nx + 1); cost[-1] = -1.0
    r = linprog(cost, A_ub=np.vstack([pad(A_ub), Az]), b_ub=np.concatenate([b_ub, np.zeros(n)]),
                A_eq=pad(A_eq) if A_eq else None, b_eq=b_eq if A_eq else None,
                bounds=[(0.0, u if np.isfinite(u) else None) for u in ub] + [(None, None)], method="highs")
    assert r.status == 0, r.message
    return -r.fun, r.x[:nx].reshape(n, nv1 * m), (job_ids, keys, wts)
