"""The packing policies (space sharing: two jobs co-located on one accelerator) behind the reference's signatures.

Mirrors scheduler/policies/policy.py:68-193 (`PolicyWithPacking`) and the four *WithPacking classes that hand an LP
over (job combination x worker type) columns to cvxpy:
  MaxMinFairnessPolicyWithPacking ............... max_min_fairness.py:116-121, 317-410
  FinishTimeFairnessPolicyWithPacking ........... finish_time_fairness.py:160-290   (stateful)
  MinTotalDurationPolicyWithPacking ............. min_total_duration.py:138-234
  ThroughputNormalizedByCostSumWithPackingSLOs .. max_sum_throughput.py:111-200
Keys of `unflattened_throughputs` are the reference's `JobIdPair` objects (duck-typed: is_pair(), singletons(),
as_tuple(), sortable, hashable); a pair's value per worker type is the two members' throughputs when co-located.

Every program is solved on the GPU by swb_lp_solve (lp.cu: revised simplex, one CTA per program) — there is no CPU
path.  The host side only assembles the sparse columns (vectorised; the reference materialises a dense
[n_single x n_comb x W] tensor in Python loops) and, for finish-time fairness, drives a 16-way multi-section over
the scalar ratio with one launch of 16 programs per pass.
"""
from __future__ import annotations

import copy

import numpy as np

from . import policies as _pol
from .policies import Policy, WaterFillingAlgorithm

_SECTIONS = 16      # candidate ratios per launch of the finish-time-fairness search


def _lp(colp, rowi, val, c, b, max_iter=0):
    """All LPs of this module go through here (tests swap in a CPU oracle backend to exercise the host logic)."""
    return _pol._engine().lp_solve(colp, rowi, val, c, b, max_iter)


class _Columns:
    """Sparse column model of one packed program: columns (combination c, worker type w) with capacity and share
    rows (policy.py:172-193), plus per-single-job linear forms T_i(x) = sum over relevant combinations."""

    @classmethod
    def from_matrix(cls, N, a, sf):
        """Single jobs only, float64 coefficients: the (job x worker type) programs of the *_Perf policies."""
        C = cls.__new__(cls)
        J, W = a.shape
        C.job_ids = C.singles = list(range(J))
        C.worker_types = list(range(W))
        C.M, C.W, C.Ns = J, W, J
        C.N = np.asarray(N, dtype=np.float64)
        C.mem = np.stack([np.arange(J), np.full(J, -1)], axis=1)
        C.thr = np.zeros((J, 2, W))
        C.thr[:, 0, :] = a
        C.sfc = np.asarray(sf, dtype=np.float64)
        C.thr_single, C.sf_single = np.asarray(a, dtype=np.float64), C.sfc
        C._live()
        return C

    def __init__(self, d, scale_factors, cluster_spec, priority_weights=None):
        job_ids = sorted(list(d.keys()))
        worker_types = sorted(list(d[job_ids[0]].keys()))
        singles = [j for j in job_ids if not j.is_pair()]
        pos = {s: i for i, s in enumerate(singles)}
        M, W, Ns = len(job_ids), len(worker_types), len(singles)
        self.job_ids, self.worker_types, self.singles = job_ids, worker_types, singles
        self.M, self.W, self.Ns = M, W, Ns
        self.N = np.array([cluster_spec[w] for w in worker_types], dtype=np.float64)
        mem = np.full((M, 2), -1, dtype=np.int64)          # single-job index of each member (-1: none / not a key)
        thr = np.zeros((M, 2, W), dtype=np.float32)        # float32 like the reference's all_m (policy.py:132)
        sfc = np.zeros(M, dtype=np.float64)                # common scale factor of the combination, 0 on mismatch
        for c, jid in enumerate(job_ids):
            row = d[jid]
            if not jid.is_pair():
                mem[c, 0] = pos[jid]
                thr[c, 0] = [row[w] for w in worker_types]
                sfc[c] = scale_factors[jid]
            else:
                tup = jid.as_tuple()
                sf = None
                for s in jid.singletons():
                    # policy.py:75-85
                    sf = 0 if (sf is not None and sf != scale_factors[s]) else scale_factors[s]
                    k = tup.index(s[0])
                    if s in pos:
                        mem[c, k] = pos[s]
                        thr[c, k] = [row[w][k] for w in worker_types]
                sfc[c] = sf
        if priority_weights is not None:                   # policy.py:158-159, in float32 like the reference
            for c in range(M):
                for k in range(2):
                    if mem[c, k] >= 0:
                        thr[c, k] /= priority_weights[singles[mem[c, k]]]
        self.mem, self.thr, self.sfc = mem, thr.astype(np.float64), sfc
        self.thr_single = np.array([[d[s][w] for w in worker_types] for s in singles], dtype=np.float64)
        self.sf_single = np.array([scale_factors[s] for s in singles], dtype=np.float64)
        self._live()

    def _live(self):
        # live columns: effective scale factor != 0 (the reference pins the others to 0, max_min_fairness.py:396-399)
        # and a worker type with capacity (capacity 0 forces the column to 0 through the capacity row)
        live_w = np.flatnonzero(self.N > 0)
        cc, ww = np.meshgrid(np.flatnonzero(self.sfc != 0), live_w, indexing="ij")
        self.col_c, self.col_w = cc.ravel(), ww.ravel()
        self.nv = len(self.col_c)
        self.live_w = live_w
        self.wrow = {int(w): r for r, w in enumerate(live_w)}

    # ---- linear forms ----
    def form(self, with_sf=False):
        """COO triplets (single i, column v, coefficient) of T_i(x) for every single job."""
        ii, vv, aa = [], [], []
        v = np.arange(self.nv)
        for k in range(2):
            i = self.mem[self.col_c, k]
            sel = i >= 0
            a = self.thr[self.col_c, k, self.col_w]
            if with_sf:
                a = a * self.sfc[self.col_c]
            ii.append(i[sel]); vv.append(v[sel]); aa.append(a[sel])
        return np.concatenate(ii), np.concatenate(vv), np.concatenate(aa)

    def base_rows(self):
        """COO triplets of the capacity rows (0 .. Wl-1) and the share rows (Wl .. Wl+Ns-1), and their right sides."""
        Wl = len(self.live_w)
        v = np.arange(self.nv)
        rr = [np.array([self.wrow[int(w)] for w in self.col_w], dtype=np.int64)]
        vv = [v]
        aa = [self.sfc[self.col_c]]
        for k in range(2):
            i = self.mem[self.col_c, k]
            sel = i >= 0
            rr.append(Wl + i[sel]); vv.append(v[sel]); aa.append(np.ones(int(sel.sum())))
        b = np.concatenate([self.N[self.live_w], np.ones(self.Ns)])
        return np.concatenate(rr), np.concatenate(vv), np.concatenate(aa), b

    def expand(self, xv):
        """Column values -> the reference's (n_comb x W) matrix, clipped to [0, 1] like every policy's return."""
        x = np.zeros((self.M, self.W))
        x[self.col_c, self.col_w] = xv
        return np.clip(x, 0.0, 1.0)

    def rates(self, xv):
        """T_i(x) for every single job."""
        i, v, a = self.form()
        return np.bincount(i, weights=a * xv[v], minlength=self.Ns)


def _csc(rows, cols, vals, n):
    """COO -> (colp, rowi, order): `order` sorts any per-entry value array into the CSC layout."""
    order = np.lexsort((rows, cols))
    colp = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(cols, minlength=n), out=colp[1:])
    return colp, rows[order].astype(np.int32), order


def _max_min(C, fi, fv, fa_batch):
    """max z : z <= sum_v fa[i, v] x_v for every single job i, base rows.  fa_batch is [S, nnz_form] (one program per
    row).  Returns x [S, nv], z [S]."""
    fa_batch = np.atleast_2d(fa_batch)
    S = fa_batch.shape[0]
    br, bv, ba, bb = C.base_rows()
    m0 = len(bb)
    nz = C.nv                                   # column index of z
    rows = np.concatenate([br, m0 + fi, m0 + np.arange(C.Ns)])
    cols = np.concatenate([bv, fv, np.full(C.Ns, nz)])
    colp, rowi, order = _csc(rows, cols, None, C.nv + 1)
    val = np.empty((S, len(rows)))
    for s in range(S):
        val[s] = np.concatenate([ba, -fa_batch[s], np.ones(C.Ns)])[order]
    c = np.zeros((S, C.nv + 1))
    c[:, nz] = 1.0
    b = np.tile(np.concatenate([bb, np.zeros(C.Ns)]), (S, 1))
    x, obj, status, stats = _lp(colp, rowi, val, c, b)
    if np.any(status != 0):
        raise RuntimeError(f"packed max-min program: simplex status {status.tolist()}")
    _max_min.last_stats = stats
    return x[:, :C.nv], x[:, nz]


def _ftf_search(C, t, n, den):
    """minimise max_i (t_i + n_i / T_i(x)) / den_i  (finish_time_fairness.py:232-246): rho is attainable iff
    max_x min_i T_i(x) / r_i(rho) >= 1 with r_i = n_i / (rho den_i - t_i).  Every x bounds rho* from above by its own
    ratio, every failed rho from below; 16 candidate ratios per launch.  Returns (rho, column values, launches)."""
    fi, fv, fa = C.form()

    def ratio(xv):
        T = C.rates(xv)
        with np.errstate(divide="ignore", invalid="ignore"):
            return float(np.max((t + np.where(n > 0, n / T, 0.0)) / den))

    lo = float(np.max(t / den))
    xv0, _ = _max_min(C, fi, fv, fa * (den / n)[fi])
    best_x, hi = xv0[0], ratio(xv0[0])
    if not np.isfinite(hi):
        return hi, None, 1
    passes = 1
    while hi - lo > 1e-9 * hi and passes < 16:
        rho = lo + (hi - lo) * (np.arange(1, _SECTIONS + 1) / (_SECTIONS + 1.0))
        r = n[None, :] / (rho[:, None] * den[None, :] - t[None, :])          # [S, Ns] needed throughputs
        xv, z = _max_min(C, fi, fv, fa[None, :] / r[:, fi])
        passes += 1
        for k in range(_SECTIONS):
            rk = ratio(xv[k])
            if rk < hi:
                hi, best_x = rk, xv[k]
            if z[k] < 1.0 - 1e-9:
                lo = max(lo, float(rho[k]))
        lo = min(lo, hi)
    return hi, best_x, passes


def _mtd(C, n):
    """min_total_duration.py:176-234.  T is attainable iff some x has T_i(x) >= n_i / T for all i  <=>
    T >= 1 / max_x min_i T_i(x) / n_i: ONE program instead of the reference's ~14 feasibility probes; the probes are
    then replayed on the verdicts.  Returns (T of the last feasible probe, T*, column values)."""
    fi, fv, fa = C.form()
    Tref = 1.0e4
    xv, z = _max_min(C, fi, fv, fa * (Tref / n)[fi])
    if not z[0] > 0:
        return None, np.inf, xv[0]
    T_star = Tref / float(z[0])
    max_T, min_T, last_max_T, last_T = 1000000.0, 100.0, 1000000.0, None      # min_total_duration.py:206-230
    while last_T is None:
        while 1.05 * min_T < max_T:
            T = (min_T + max_T) / 2.0
            if T >= T_star * (1.0 - 1e-9):
                last_T, max_T = T, T
            else:
                min_T = T
        max_T, min_T = last_max_T * 10.0, last_max_T
        last_max_T *= 10
    return last_T, T_star, xv[0]


def _max_sum(C, cvec, need):
    """max cvec . x under the base rows and  T_i(x) >= need[i]  for the singles in `need` (a dict).
    Returns (column values, objective, status)."""
    fi, fv, fa = C.form()
    br, bv, ba, bb = C.base_rows()
    m0 = len(bb)
    rows, cols, vals, b = [br], [bv], [ba], [bb]
    for k, (i, nd) in enumerate(need.items()):          # -T_i(x) / need_i <= -1
        sel = fi == i
        rows.append(np.full(int(sel.sum()), m0 + k)); cols.append(fv[sel]); vals.append(-fa[sel] / nd)
    b.append(-np.ones(len(need)))
    rows, cols, vals, b = map(np.concatenate, (rows, cols, vals, b))
    colp, rowi, order = _csc(rows, cols, None, C.nv)
    x, obj, status, _ = _lp(colp, rowi, vals[order], cvec, b)
    return x[0], float(obj[0]), int(status[0])


def hetero_lp(mode, N, a, sf, t=None, n=None, den=None):
    """policies._hetero for MORE worker types than hetero.cu's Dantzig-Wolfe master enumerates (W > 4, max-sum W > 3,
    e.g. all six types of tacc_throughputs.json): the same programs (max_min_fairness.py:53-113,
    finish_time_fairness.py:66-157, min_total_duration.py:55-135, max_sum_throughput.py:49-108) as general LPs on
    swb_lp_solve.  a is J x W over worker types WITH capacity.  Returns (x [J, W], objective, rc)."""
    a = np.asarray(a, dtype=np.float64)
    J, W = a.shape
    if W + 2 * J > 2048:
        raise NotImplementedError("more than 4 worker types with capacity: the general LP path holds W + 2 J <= 2048 rows")
    C = _Columns.from_matrix(N, a, sf)
    back = lambda xv: C.expand(xv)
    if mode == _pol.POL_MAXMIN:
        fi, fv, fa = C.form()
        xv, z = _max_min(C, fi, fv, fa)
        return back(xv[0]), float(z[0]), 0
    if mode == _pol.POL_FTF:
        rho, xv, _ = _ftf_search(C, np.asarray(t, float), np.asarray(n, float), np.asarray(den, float))
        if xv is None:
            return np.zeros_like(a), rho, 1
        return back(xv), rho, 0
    if mode == _pol.POL_MTD:
        T, _, xv = _mtd(C, np.asarray(n, float))
        return back(xv), T, (0 if T is not None else 1)
    if mode == _pol.POL_MAXSUM:
        fi, fv, fa = C.form()
        need = {}
        if t is not None:       # SLO floors on the THROUGHPUT a * cost (policies.py: a = thr / cost, den = cost per type)
            C2 = _Columns.from_matrix(N, a * np.asarray(den, float)[None, :], sf)
            need = {int(i): float(v) for i, v in enumerate(np.asarray(t, float)) if v > 0}
            cvec = np.bincount(fv, weights=fa, minlength=C.nv)
            xv, obj, st = _max_sum(C2, cvec, need)
        else:
            xv, obj, st = _max_sum(C, np.bincount(fv, weights=fa, minlength=C.nv), need)
        if st == 1:
            return np.zeros_like(a), 0.0, 1
        if st != 0:
            raise RuntimeError(f"max-sum program: simplex status {st}")
        return back(xv), obj, 0
    raise ValueError(mode)


class PolicyWithPacking(Policy):
    """policy.py:68-193 (host side).  `flatten`/`unflatten` keep the reference's index tuple."""

    def __init__(self, solver="ECOS"):
        Policy.__init__(self, solver)

    def _columns(self, d, scale_factors, cluster_spec, priority_weights=None):
        if len(d) == 0:
            return None
        first = d[next(iter(d))]
        if len(first) == 0:
            return None
        C = _Columns(d, scale_factors, cluster_spec, priority_weights)
        self._num_workers = [cluster_spec[w] for w in C.worker_types]
        return C

    @staticmethod
    def _unflatten(C, x):
        return {jid: dict(zip(C.worker_types, row)) for jid, row in zip(C.job_ids, x.tolist())}


def _proportional(thr, N):
    """proportional.py:14-43: x_iw = (N_w / m) / max row sum -> N_w / sum N."""
    return thr @ (N / N.sum())


def _isolated(thr, sf, N):
    """isolated.py:14-53."""
    m = thr.shape[0]
    x = (N[None, :] / m) / sf[:, None]
    x = x / np.maximum(x.sum(axis=1), 1.0)[:, None]
    return (thr * x).sum(axis=1)


class MaxMinFairnessPolicyWithPacking(PolicyWithPacking):
    def __init__(self, solver):
        PolicyWithPacking.__init__(self, solver)
        self._name = "MaxMinFairness_Packing"

    def get_allocation(self, unflattened_throughputs, scale_factors, unflattened_priority_weights, cluster_spec):
        C = self._columns(unflattened_throughputs, scale_factors, cluster_spec, unflattened_priority_weights)
        if C is None or C.Ns == 0:
            return None
        prop = _proportional(C.thr_single, C.N)                     # max_min_fairness.py:352-360
        fi, fv, fa = C.form(with_sf=True)                           # :367-381
        xv, z = _max_min(C, fi, fv, fa / prop[fi])
        self.last_objective = float(z[0])
        return self._unflatten(C, C.expand(xv[0]))


    def get_allocation_using_job_type_throughputs(self, unflattened_throughputs, job_id_to_job_type_key,
                                                  scale_factors, unflattened_priority_weights, cluster_spec):
        """max_min_fairness.py:122-316: the job x JOB-TYPE formulation (O(n a) instead of O(n^2) variables).
        `unflattened_throughputs[type][worker_type][other_type or None]`, job types are (name, scale_factor) keys.
        The reference's equality rows are folded into the columns: the "i paired with its own type" variables of one
        type are ONE column (the reference constrains them all equal, :256-271), variables the reference pins to 0
        (types with a single job, :295-298) or whose throughput it zeroes (other scale factor, :171-172) are dropped;
        "a with b == b with a" (:226-254) stays as a pair of inequality rows."""
        job_ids = sorted(job_id_to_job_type_key.keys())
        if len(job_ids) == 0:
            return None
        keys = sorted(unflattened_throughputs.keys())
        wts = sorted(cluster_spec.keys())
        N = np.array([cluster_spec[w] for w in wts], dtype=np.float64)
        n, a = len(job_ids), len(keys)
        kidx = {k: i for i, k in enumerate(keys)}
        tof = np.array([kidx[job_id_to_job_type_key[j]] for j in job_ids])          # type of job i
        members = [np.flatnonzero(tof == A) for A in range(a)]
        sf = np.array([scale_factors[j] for j in job_ids], dtype=np.float64)
        pw = np.array([unflattened_priority_weights[j] for j in job_ids], dtype=np.float64)
        flat = np.zeros((a, len(wts), 1 + a), dtype=np.float32)                     # :165-180, float32 like the reference
        for A, key in enumerate(keys):
            for k, w in enumerate(wts):
                for j, other in enumerate([None] + keys):
                    flat[A, k, j] = 0.0 if (j > 0 and other[1] != key[1]) else unflattened_throughputs[key][w][other]
        flat = flat.astype(np.float64)
        alone = np.array([[unflattened_throughputs[keys[A]][w][None] for w in wts] for A in range(a)], dtype=np.float64)
        prop = _proportional(alone[tof], N)                                         # :273-284 (float64, not all_m's float32)
        scale = sf / (pw * prop)                                                    # :286-301
        live = np.flatnonzero(N > 0)
        Wl = len(live)
        # ---- columns: (kind, i or A, k, B) ----
        cols = []            # tuples (kind, owner, k, other): kind 0 isolated, 1 cross (job i with type B), 2 same (type A)
        for i in range(n):
            for k in live:
                cols.append((0, i, k, -1))
        for i in range(n):
            A = tof[i]
            for B in range(a):
                if B != A and keys[B][1] == keys[A][1]:
                    for k in live:
                        cols.append((1, i, k, B))
        for A in range(a):
            if len(members[A]) >= 2:
                for k in live:
                    cols.append((2, A, k, A))
        nv = len(cols)
        wrow = {int(k): r for r, k in enumerate(live)}
        R_cap, R_share, R_epi = 0, Wl, Wl + n
        rows, cidx, vals = [], [], []

        def put(r, v, x):
            rows.append(r); cidx.append(v); vals.append(x)
        pair_rows = {}
        R_eq = Wl + 2 * n
        for A in range(a):
            for B in range(A + 1, a):
                if keys[A][1] == keys[B][1]:
                    for k in live:
                        pair_rows[(A, B, int(k))] = R_eq
                        R_eq += 2
        for v, (kind, o, k, B) in enumerate(cols):
            if kind == 0:
                put(R_cap + wrow[int(k)], v, sf[o]); put(R_share + o, v, 1.0)
                put(R_epi + o, v, -flat[tof[o], k, 0] * scale[o])
            elif kind == 1:
                A = tof[o]
                put(R_cap + wrow[int(k)], v, 0.5 * sf[o]); put(R_share + o, v, 1.0)
                put(R_epi + o, v, -flat[A, k, 1 + B] * scale[o])
                r = pair_rows[(min(A, B), max(A, B), int(k))]
                sgn = 1.0 if A < B else -1.0
                put(r, v, sgn); put(r + 1, v, -sgn)
            else:
                put(R_cap + wrow[int(k)], v, 0.5 * sf[members[o]].sum())
                for i in members[o]:
                    put(R_share + i, v, 1.0)
                    put(R_epi + i, v, -flat[o, k, 1 + o] * scale[i])
        for i in range(n):
            put(R_epi + i, nv, 1.0)
        m = R_eq
        b = np.zeros(m)
        b[:Wl] = N[live]
        b[R_share:R_share + n] = 1.0
        rows, cidx, vals = np.array(rows), np.array(cidx), np.array(vals)
        keep = vals != 0.0
        colp, rowi, order = _csc(rows[keep], cidx[keep], None, nv + 1)
        cost = np.zeros(nv + 1)
        cost[nv] = 1.0
        x, obj, status, stats = _lp(colp, rowi, vals[keep][order], cost, b)
        if status[0] != 0:
            raise RuntimeError(f"MaxMinFairness_Packing (job types): simplex status {int(status[0])}")
        self.last_objective = float(obj[0])
        xv = np.clip(x[0, :nv], 0.0, 1.0)
        # ---- back to the reference's nested dict, then policy.py:195-260 ----
        alloc = {j: {w: {o: 0.0 for o in [None] + keys} for w in wts} for j in job_ids}
        for v, (kind, o, k, B) in enumerate(cols):
            if kind == 0:
                alloc[job_ids[o]][wts[k]][None] = float(xv[v])
            elif kind == 1:
                alloc[job_ids[o]][wts[k]][keys[B]] = float(xv[v])
            else:
                for i in members[o]:
                    alloc[job_ids[i]][wts[k]][keys[o]] = float(xv[v])
        self.last_job_type_allocation = alloc
        return self.convert_job_type_allocation(alloc, job_id_to_job_type_key)

    @staticmethod
    def convert_job_type_allocation(allocation, job_id_to_job_type_key):
        """policy.py:195-260: x_{i,j} = x_{i,type(j)} x_{j,type(i)} / sum_k x_{k,type(j)} over the jobs k of type(i)."""
        job_ids = sorted(allocation.keys())
        wts = sorted(allocation[job_ids[0]].keys())
        keys = sorted(set(job_id_to_job_type_key[j] for j in job_ids))
        make_pair = type(job_ids[0])
        tot = {w: {k: {o: 0.0 for o in keys} for k in keys} for w in wts}
        for j in job_ids:
            k = job_id_to_job_type_key[j]
            for w in wts:
                row = allocation[j][w]
                for o in keys:
                    tot[w][k][o] += row.get(o, 0.0)
        out = {}
        for i, j in enumerate(job_ids):
            k = job_id_to_job_type_key[j]
            out[j] = {w: allocation[j][w][None] for w in wts}
            for j2 in job_ids[i + 1:]:
                k2 = job_id_to_job_type_key[j2]
                d = {}
                for w in wts:
                    cur = tot[w][k][k2]
                    if cur > 0.0:
                        if k == k2:
                            cur -= allocation[j][w][k]
                        d[w] = allocation[j][w][k2] * allocation[j2][w][k] / cur
                    else:
                        d[w] = 0.0
                out[make_pair(j[0], j2[0])] = d
        return out


class FinishTimeFairnessPolicyWithPacking(PolicyWithPacking):
    def __init__(self, solver):
        PolicyWithPacking.__init__(self, solver)
        self._name = "FinishTimeFairness_Packing"
        self._cumulative_isolated_time = {}
        self._isolated_throughputs_prev_iteration = {}
        self._num_steps_remaining_prev_iteration = {}

    def get_allocation(self, unflattened_throughputs, scale_factors, unflattened_priority_weights,
                       times_since_start, num_steps_remaining, cluster_spec):
        C = self._columns(unflattened_throughputs, scale_factors, cluster_spec, unflattened_priority_weights)
        if C is None or C.Ns == 0:
            self._isolated_throughputs_prev_iteration = {}
            self._num_steps_remaining_prev_iteration = {}
            return None
        iso = _isolated(C.thr_single, C.sf_single, C.N)             # finish_time_fairness.py:199-210
        for s in C.singles:                                          # :214-230
            if s not in self._cumulative_isolated_time:
                self._cumulative_isolated_time[s] = 0
            if s in self._num_steps_remaining_prev_iteration:
                self._cumulative_isolated_time[s] += (
                    self._num_steps_remaining_prev_iteration[s] - num_steps_remaining[s]
                ) / self._isolated_throughputs_prev_iteration[s]
        n = np.array([num_steps_remaining[s] for s in C.singles], dtype=np.float64)
        t = np.array([times_since_start[s] for s in C.singles], dtype=np.float64)
        den = np.array([self._cumulative_isolated_time[s] for s in C.singles], dtype=np.float64) + n / iso
        hi, best_x, passes = _ftf_search(C, t, n, den)
        if best_x is None:
            raise RuntimeError("FinishTimeFairness_Packing: a job has no column with positive throughput")
        self.last_objective, self.last_passes = hi, passes
        self._num_steps_remaining_prev_iteration = copy.copy(num_steps_remaining)
        self._isolated_throughputs_prev_iteration = {s: iso[i] for i, s in enumerate(C.singles)}
        return self._unflatten(C, C.expand(best_x))


class MinTotalDurationPolicyWithPacking(PolicyWithPacking):
    def __init__(self, solver):
        PolicyWithPacking.__init__(self, solver)
        self._name = "MinTotalDuration_Packing"

    def get_allocation(self, unflattened_throughputs, scale_factors, num_steps_remaining, cluster_spec):
        C = self._columns(unflattened_throughputs, scale_factors, cluster_spec)
        if C is None or C.Ns == 0:
            return None
        n = np.array([num_steps_remaining[s] for s in C.singles], dtype=np.float64)
        last_T, T_star, xv = _mtd(C, n)
        if last_T is None:
            raise RuntimeError("MinTotalDuration_Packing: a job cannot make progress on this cluster")
        self.last_objective, self.last_T_star = last_T, T_star
        return self._unflatten(C, C.expand(xv))


class ThroughputNormalizedByCostSumWithPackingSLOs(PolicyWithPacking):
    def __init__(self, solver):
        Policy.__init__(self, solver)
        self._name = "ThroughputNormalizedByCostSum_PackingSLOs"

    def get_allocation(self, unflattened_throughputs, scale_factors, cluster_spec, instance_costs=None, SLOs={},
                       num_steps_remaining={}):
        C = self._columns(unflattened_throughputs, scale_factors, cluster_spec)
        if C is None or C.Ns == 0:
            return None
        cost = np.ones(C.W)
        if instance_costs is not None:
            cost = np.array([instance_costs[w] for w in C.worker_types], dtype=np.float64)
        fi, fv, fa = C.form()
        cvec = np.bincount(fv, weights=fa / cost[C.col_w[fv]], minlength=C.nv)    # max_sum_throughput.py:146-163
        need = {}
        for job_id in SLOs:                                                       # :170-183
            i = C.job_ids.index(job_id)
            assert job_id in num_steps_remaining
            need[i] = num_steps_remaining[job_id] / SLOs[job_id]
        xv, obj, status = _max_sum(C, cvec, need)
        self.used_SLOs = bool(need) and status == 0
        if need and status == 1:          # "x.value is None": the reference warns and solves again without the SLO rows
            print("WARNING: No allocation possible with provided SLOs!")
            xv, obj, status = _max_sum(C, cvec, {})
        if status != 0:
            raise RuntimeError(f"ThroughputNormalizedByCostSum_PackingSLOs: simplex status {status}")
        self.last_objective = obj
        return self._unflatten(C, C.expand(xv))


class MaxMinFairnessWaterFillingPolicyWithPacking(PolicyWithPacking, WaterFillingAlgorithm):
    """max_min_fairness_water_filling.py:569-718.  The iteration loop, the entity re-weighting and the MILP's
    infeasibility quirk are WaterFillingAlgorithm's (policies.py); the two programs of an iteration run on swb_lp_solve
    over the (combination, worker type) columns:
      _get_allocation (:81-189)        max c :  c / mult_i - net_i(x) <= -so_far_i (active i),  net_i >= lower_i,  c <= M
      _get_bottleneck_jobs (:191-305)  max sum z :  net_i >= so_far_i (1 + (slack - 1) z_i),  0 <= z_i <= 1 (active i)
    — the second is the tight relaxation of the reference's big-M MILP (z_i < 0.5 reads as "job i is a bottleneck"),
    the same reading swb_policy_waterfill_step uses for the unpacked classes."""

    def __init__(self, priority_reweighting_policies=None):
        WaterFillingAlgorithm.__init__(self, priority_reweighting_policies)
        PolicyWithPacking.__init__(self, solver=None)
        self._name = "MaxMinFairnessWaterFilling_Packing"

    def _waterfill_step(self, N, thr, sf, prop, lower, mult, M, slack):
        C = self._C
        fi, fv, fa = C.form()
        fa = fa / prop[fi]                                  # net_i(x) = T_i(x) / prop_i
        br, bv, ba, bb = C.base_rows()
        m0, Ns = len(bb), C.Ns
        active = mult > 0
        # ---- LP: columns x (nv) and c ----
        nc = C.nv
        rows = np.concatenate([br, m0 + fi, m0 + np.flatnonzero(active), [m0 + Ns]])
        cols = np.concatenate([bv, fv, np.full(int(active.sum()), nc), [nc]])
        vals = np.concatenate([ba, -fa, 1.0 / mult[active], [1.0]])
        b = np.concatenate([bb, -lower, [M]])               # lower_i == so_far_i for the active jobs
        colp, rowi, order = _csc(rows, cols, None, C.nv + 1)
        cost = np.zeros(C.nv + 1)
        cost[nc] = 1.0
        x, obj, status, _ = _lp(colp, rowi, vals[order], cost, b)
        if status[0] != 0:
            return None, None, None
        xs, c = x[0, :C.nv], float(x[0, nc])
        so_far = lower + np.where(active, c / np.where(active, mult, 1.0), 0.0)
        # ---- bottleneck program: columns x (nv) and z_i for the active jobs ----
        act = np.flatnonzero(active)
        na = len(act)
        zcol = C.nv + np.arange(na)
        rows = np.concatenate([br, m0 + fi, m0 + act, m0 + Ns + np.arange(na)])
        cols = np.concatenate([bv, fv, zcol, zcol])
        vals = np.concatenate([ba, -fa, so_far[act] * (slack - 1.0), np.ones(na)])
        b = np.concatenate([bb, -so_far, np.ones(na)])
        colp, rowi, order = _csc(rows, cols, None, C.nv + na)
        cost = np.zeros(C.nv + na)
        cost[zcol] = 1.0
        x2, _, status2, _ = _lp(colp, rowi, vals[order], cost, b)
        z = np.zeros(Ns)
        if status2[0] == 0:
            z[act] = x2[0, zcol]
        return xs, c, z

    def get_allocation(self, unflattened_throughputs, scale_factors, unflattened_priority_weights, cluster_spec,
                       entity_weights=None, entity_to_job_mapping=None, verbose=False,
                       return_effective_throughputs=False):
        C = self._columns(unflattened_throughputs, scale_factors, cluster_spec)
        if C is None or C.Ns == 0:
            return None
        self._C = C
        prop = _proportional(C.thr_single, C.N)                                   # :640-652
        fi, fv, fa = C.form(with_sf=True)
        self._M = float(np.max(fa / prop[fi]))                                    # _get_M :583-602
        xv = self._run_get_allocation_iterations(C.singles, None, C.sf_single, C.N, prop, self._M, entity_weights,
                                                 unflattened_priority_weights, entity_to_job_mapping, verbose)
        x = C.expand(xv)
        if return_effective_throughputs:
            return C.rates(xv) / prop, C.singles
        return self._unflatten(C, x)
