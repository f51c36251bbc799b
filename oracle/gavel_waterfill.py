"""CPU ORACLE (test infrastructure, not product code) for the water-filling max-min policies.

Restates on scipy's HiGHS the two programs of WaterFillingAlgorithm and its iteration loop
  scheduler/policies/max_min_fairness_water_filling.py:81-189   _get_allocation      (LP, ECOS in the reference)
  scheduler/policies/max_min_fairness_water_filling.py:191-305  _get_bottleneck_jobs (MILP, GLPK_MI in the reference)
  scheduler/policies/max_min_fairness_water_filling.py:307-413  _run_get_allocation_iterations
  scheduler/policies/max_min_fairness_water_filling.py:16-79    _compute_priority_weights (entity re-weighting)
for MaxMinFairnessWaterFillingPolicy[WithPerf] (:416-576; no packing).  Plain arrays in, plain arrays out.
Pinning: the reference ships no known-answer test or golden pickle for these policies ("parity unpinned" at the value
level); the restatement follows the reference statement by statement, including its quirks: the objective is capped
by M through the additive terms of the saturated jobs, a job with z = 0 must stay BELOW so_far * slack - epsilon (which
makes the MILP infeasible for so_far < 0.1 — the reference's `except` then freezes every active job), and the
returned x is the LAST successful LP's.  Only tests/ may import this.
"""
import numpy as np
from scipy.optimize import Bounds, LinearConstraint, linprog, milp

from oracle import gavel_lp as gl

SLACK, EPSILON = 1.0001, 1e-5


def lp_step(thr, sf, N, prop, lower, mult, add, so_far):
    """_get_allocation (:81-189): max min_j [(net_j - so_far_j) mult_j + add_j], net_j = thr_j.x_j / prop_j,
    base constraints, net >= lower.  Returns (x, c) or (None, None)."""
    J, W = thr.shape
    A, b = gl._base(J, W, sf, N)
    A = np.hstack([A, np.zeros((A.shape[0], 1))])
    rows, rhs = [], []
    for j in range(J):                      # c <= (net_j - so_far_j) mult_j + add_j
        r = np.zeros(J * W + 1)
        r[j * W:(j + 1) * W] = -thr[j] / prop[j] * mult[j]
        r[-1] = 1.0
        rows.append(r); rhs.append(add[j] - so_far[j] * mult[j])
    for j in range(J):                      # net_j >= lower_j
        r = np.zeros(J * W + 1)
        r[j * W:(j + 1) * W] = -thr[j] / prop[j]
        rows.append(r); rhs.append(-lower[j])
    cost = np.zeros(J * W + 1); cost[-1] = -1.0
    res = linprog(cost, A_ub=np.vstack([A] + [np.array(rows)]), b_ub=np.concatenate([b, np.array(rhs)]),
                  bounds=[(0, None)] * (J * W) + [(None, None)], method="highs")
    if res.status != 0:
        return None, None
    return res.x[:-1].reshape(J, W), float(res.x[-1])


def bottleneck_milp(thr, sf, N, prop, lower, so_far, zmask, M, slack=SLACK, epsilon=EPSILON):
    """_get_bottleneck_jobs (:191-305): max sum z;  M z >= net - so_far slack + eps;  M (1 - z) >= so_far slack - net;
    net >= lower;  z_j = 0 where zmask_j.  Returns z or raises (the reference raises on a non-optimal status)."""
    J, W = thr.shape
    nx = J * W
    A, b = gl._base(J, W, sf, N)
    cons = [LinearConstraint(np.hstack([A, np.zeros((A.shape[0], J))]), -np.inf, b)]
    R1 = np.zeros((J, nx + J)); R2 = np.zeros((J, nx + J)); R3 = np.zeros((J, nx + J))
    for j in range(J):
        net = thr[j] / prop[j]
        R1[j, j * W:(j + 1) * W] = net; R1[j, nx + j] = -M          # net - M z <= so_far slack - eps
        R2[j, j * W:(j + 1) * W] = -net; R2[j, nx + j] = M          # -net + M z <= M - so_far slack
        R3[j, j * W:(j + 1) * W] = net                               # net >= lower
    cons.append(LinearConstraint(R1, -np.inf, so_far * slack - epsilon))
    cons.append(LinearConstraint(R2, -np.inf, M - so_far * slack))
    cons.append(LinearConstraint(R3, lower, np.inf))
    ub = np.concatenate([np.full(nx, np.inf), np.where(zmask > 0, 0.0, 1.0)])
    res = milp(np.concatenate([np.zeros(nx), -np.ones(J)]), constraints=cons,
               integrality=np.concatenate([np.zeros(nx), np.ones(J)]), bounds=Bounds(np.zeros(nx + J), ub))
    if res.status != 0:
        raise RuntimeError("non-optimal allocation in _get_bottleneck_jobs")
    return np.round(res.x[nx:])


def compute_priority_weights(entity_weights, priority_weights, entity_to_job_mapping, final, job_ids, policies):
    """_compute_priority_weights (:16-79) on dicts keyed like the reference's."""
    if policies is None:
        return priority_weights
    if entity_to_job_mapping is None:
        raise ValueError("entity_to_job_mapping cannot be None when priority_reweighting_policies is not None!")
    out = {}
    for ent in entity_to_job_mapping:
        pol = policies[ent]
        ew = entity_weights[ent]
        if pol == "fairness":
            tot = 0.0
            for j in entity_to_job_mapping[ent]:
                if j in final:
                    continue
                tot += float(priority_weights[j])
            for j in entity_to_job_mapping[ent]:
                out[j] = 0.0 if j in final else ew * (float(priority_weights[j]) / tot)
        elif pol == "fifo":
            entity_to_job_mapping[ent].sort()
            done = False
            for j in entity_to_job_mapping[ent]:
                if j in final:
                    out[j] = 0.0
                elif not done:
                    out[j] = ew
                    done = True
                else:
                    out[j] = 0.0
        else:
            raise ValueError("Unknown priority reweighting policy!")
    return out


def run_iterations(job_ids, thr, sf, N, prop, unflattened_priority_weights, M, entity_weights=None,
                   entity_to_job_mapping=None, policies=None, lp=lp_step, bottleneck=bottleneck_milp, log=None):
    """_run_get_allocation_iterations (:307-413) around pluggable LP / bottleneck solvers (the GPU test plugs the
    kernels in here to compare iteration by iteration).  Returns (x, so_far, final dict, iterations)."""
    J = len(job_ids)
    final = {}
    so_far = np.zeros(J)
    done = False
    it = 0
    c, x, mask = 0, None, None
    while not done:
        pw_d = compute_priority_weights(entity_weights, unflattened_priority_weights, entity_to_job_mapping, final,
                                        job_ids, policies)
        pw = np.array([1.0 / pw_d[j] if pw_d[j] > 0 else 0.0 for j in job_ids])
        old = (x, c, mask)
        mult = np.zeros(J); mask = np.zeros(J); add = np.zeros(J)
        for i, j in enumerate(job_ids):
            if j not in final:
                if pw[i] > 0.0:
                    mult[i] = pw[i] * sf[i]
                    mask[i] = 1.0 / mult[i]
                else:
                    add[i] = M
            else:
                add[i] = M
        lower = np.array([final[j] if j in final else so_far[i] for i, j in enumerate(job_ids)])
        try:
            x, c = lp(thr, sf, N, prop, lower, mult, add, so_far)
            if x is None:
                x, c, mask = old
                done = True
            else:
                so_far = so_far + mask * c
        except Exception:
            x, c, mask = old
            done = True
        lower = np.array([final[j] if j in final else so_far[i] for i, j in enumerate(job_ids)])
        zmask = np.array([1.0 if (j in final or pw[i] == 0.0) else 0.0 for i, j in enumerate(job_ids)])
        try:
            z = bottleneck(thr, sf, N, prop, lower, so_far, zmask, M)
        except Exception:
            z = np.zeros(J)
        before = len(final)
        for i, j in enumerate(job_ids):
            if j not in final and (z is None or not z[i]) and pw[i] > 0.0:
                final[j] = so_far[i]
        if log is not None:
            log.append(dict(c=c, so_far=so_far.copy(), z=None if z is None else np.array(z, float), nfinal=len(final)))
        if before == len(final):
            done = True
        it += 1
        if len(final) == J:
            done = True
    return x, so_far, final, it


def water_filling_perf(thr, sf, priority, N, **kw):
    """MaxMinFairnessWaterFillingPolicyWithPerf.get_allocation (:476-576) on arrays: returns (x clipped to [0, 1],
    normalised effective throughputs, iterations)."""
    thr = np.asarray(thr, float); sf = np.asarray(sf, float); N = np.asarray(N, float)
    J = thr.shape[0]
    prop = gl.proportional_throughputs(thr, N)
    M = float(np.max(thr / prop[:, None] * sf[:, None]))
    ids = list(range(J))
    x, so_far, final, it = run_iterations(ids, thr, sf, N, prop, {j: priority[j] for j in ids}, M, **kw)
    x = np.clip(x, 0.0, 1.0)
    return x, (thr * x).sum(axis=1) / prop, it
