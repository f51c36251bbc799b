"""CPU ORACLE backend (test infrastructure, not product code) for shockwave_b200/policies.py.

`policies.py` hands every LP to the CUDA library through two module-level functions, `_pooled` and `_hetero`.
This module restates both on scipy's HiGHS (oracle/gavel_lp.py) with the same signatures, so that the
product's HOST logic (flatten / pooling / stateful FinishTimeFairness bookkeeping / AlloX queueing) can be
driven closed-loop by the UNMODIFIED reference simulator on a machine without a GPU, and the LP oracle can be
pinned end-to-end against the reference's golden pickles (tests/golden/make_policy_pins.py).
Only tests/ and tests/golden/ scripts may import this.
"""
import contextlib

import numpy as np

from oracle import gavel_lp as gl

POL_MAXMIN, POL_FTF, POL_MTD, POL_MAXSUM, POL_ISOLATED = 1, 2, 3, 4, 5


def _ftf(thr, sf, t, n, den, N, tol=1e-10):
    """finish_time_fairness.py:66-157 with the denominators given: bisection on rho over LP feasibility."""
    def need(rho):
        room = rho * den - t
        return None if np.any(room <= 0) else n / room
    lo, hi = 0.0, 1.0
    for _ in range(200):
        nd = need(hi)
        if nd is not None and gl.feasible_rates(thr, sf, N, nd) is not None:
            break
        lo, hi = hi, hi * 2.0
    else:
        return None, hi
    while hi - lo > tol * hi:
        mid = 0.5 * (lo + hi)
        nd = need(mid)
        if nd is not None and gl.feasible_rates(thr, sf, N, nd) is not None:
            hi = mid
        else:
            lo = mid
    return gl.feasible_rates(thr, sf, N, need(hi)), hi


def hetero_cpu(mode, N, a, sf, t=None, n=None, den=None):
    a = np.asarray(a, dtype=float)
    sf = np.asarray(sf, dtype=float)
    N = np.asarray(N, dtype=float)
    if mode == POL_MAXMIN:
        z, x = gl.max_min(a, sf, N)
        return x, z, 0
    if mode == POL_FTF:
        x, rho = _ftf(a, sf, np.asarray(t, float), np.asarray(n, float), np.asarray(den, float), N)
        if x is None:
            return np.zeros_like(a), rho, 1
        return x, rho, 0
    if mode == POL_MTD:
        T, x = gl.min_total_duration_perf(a, sf, np.asarray(n, float), N)
        return x, T, 0
    if mode == POL_MAXSUM:
        # SLO floors: pooled callers (W = 1, den None) pass t = lo_j (a time fraction); heterogeneous callers pass
        # t = needed throughput and den = instance cost per type (a = thr / cost)
        need, thr_rows = None, None
        if t is not None and den is None:
            need = np.asarray(t, float) * a[:, 0]
        elif t is not None:
            need, thr_rows = np.asarray(t, float), a * np.asarray(den, float)[None, :]
        if thr_rows is None:
            v, x = gl.max_sum_throughput(a, sf, N, need=need)
        else:
            v, x = gl.max_sum_throughput(thr_rows, sf, N, costs=np.asarray(den, float), need=need)
        if v is None:
            return np.zeros_like(a), 0.0, 1
        return x, v, 0
    raise ValueError(mode)


def pooled_cpu(mode, N, coef, sf, t=None, n=None, den=None, select="centre"):
    """select="vertex": the x HiGHS returns; "centre": the interior-point selection (gavel_lp.analytic_centre_box)
    the reference's ECOS / Gurobi-barrier iterates converge to — same objective."""
    coef = np.asarray(coef, dtype=float)
    sf = np.asarray(sf, dtype=float)
    N = float(N)
    if mode == POL_ISOLATED:        # isolated.py:46-53 / proportional.py / gandiva_fair_proportional.py, pooled
        v = (N / len(coef)) / coef
        return np.minimum(v, 1.0), 0.0, 0
    x, obj, rc = hetero_cpu(mode, [N], coef[:, None], sf, t, n, den)
    x = np.clip(x[:, 0], 0.0, 1.0)
    if select == "vertex" or rc != 0:
        return x, obj, rc
    if mode == POL_MAXMIN:
        x = gl.analytic_centre_box(obj / coef, sf, N)
    elif mode == POL_FTF:
        lo = np.asarray(n, float) / (coef * (obj * np.asarray(den, float) - np.asarray(t, float)))
        x = gl.analytic_centre_box(lo, sf, N, w_lo=2.0, w_x=0.0)
    elif mode == POL_MTD:
        x = gl.analytic_centre_box(np.asarray(n, float) / (obj * coef), sf, N)
    elif mode == POL_MAXSUM:
        x, _ = gl.max_sum_pooled_centre(coef, sf, N, lo=t)
    return x, obj, rc


def waterfill_step_cpu(N, thr, sf, prop, lower, mult, M, slack=1.0001):
    """policies._waterfill_step on the HiGHS restatement (oracle/gavel_waterfill.py): the LP and the bottleneck MILP of
    one water-filling iteration; returns (x, c, z) or (None, None, None)."""
    from oracle import gavel_waterfill as wf
    thr, sf, prop = np.asarray(thr, float), np.asarray(sf, float), np.asarray(prop, float)
    lower, mult, N = np.asarray(lower, float), np.asarray(mult, float), np.asarray(N, float)
    add = np.where(mult > 0, 0.0, M)
    x, c = wf.lp_step(thr, sf, N, prop, lower, mult, add, lower)      # objective terms (net - lower) * mult
    if x is None:
        return None, None, None
    so_far = lower + np.where(mult > 0, c / np.where(mult > 0, mult, 1.0), 0.0)
    try:
        z = wf.bottleneck_milp(thr, sf, N, prop, so_far, so_far, (mult <= 0).astype(float), M, slack, wf.EPSILON)
    except RuntimeError:
        z = np.zeros(len(sf))
    return x, c, z


@contextlib.contextmanager
def cpu_backend():
    """Route shockwave_b200.policies through the HiGHS oracle for the duration of the block."""
    from shockwave_b200 import policies as P
    saved = (P._pooled, P._hetero)
    saved_wf = P._waterfill_step
    P._waterfill_step = waterfill_step_cpu
    saved_eg = P._eisenberg_gale
    P._eisenberg_gale = lambda N, coef, sf, present, iters=0: gl.eisenberg_gale_batch(N, coef, sf, present)
    def _het(mode, N, a, sf, t=None, n=None, den=None):
        out = hetero_cpu(mode, N, a, sf, t, n, den)
        _het.last_stats = (0, 0)
        return out
    class _Eng:       # AlloXPolicy calls the engine directly (swb_allox_assign)
        @staticmethod
        def allox_assign(p, t, wtype):
            return gl.allox_assignment(p, t, wtype)
    saved_engine = P._engine
    P._pooled, P._hetero, P._engine = pooled_cpu, _het, (lambda: _Eng)
    try:
        yield P
    finally:
        P._pooled, P._hetero = saved
        P._waterfill_step = saved_wf
        P._eisenberg_gale = saved_eg
        P._engine = saved_engine
