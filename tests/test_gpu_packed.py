"""GPU parity of swb_lp_solve (lp.cu) and of the packing policies built on it (shockwave_b200/packing.py) against
the HiGHS oracle (oracle/gavel_packed.py: the reference's *WithPacking programs restated densely).

LP optima are not unique in x: parity is on the objective (1e-6 relative, the bar of the other Gavel policies; the
simplex is exact to roundoff, so most cases agree to 1e-9), on the reference's base constraints (policy.py:172-193)
and, for the bisection policies, on the probe sequence."""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.optimize import linprog

from oracle import gavel_packed as gp
from shockwave_b200 import policies as P
from shockwave_b200 import packing as pk
from tests.packing_fixtures import instance
from tests.test_oracle_packed import COSTS, SPEC, _base_ok, random_lp

pytestmark = pytest.mark.gpu


def test_lp_solve_random_programs():
    eng = P._engine()
    rng = np.random.default_rng(0)
    seen = set()
    for trial in range(90):
        A, b, c = random_lp(rng, trial % 3)
        x, obj, st, stats = eng.lp_solve(A.indptr, A.indices, A.data, c, b)
        r = linprog(-c, A_ub=A, b_ub=b, bounds=(0, None), method="highs")
        if r.status == 4:
            continue
        assert int(st[0]) == {0: 0, 2: 1, 3: 2}[r.status], (trial, st, r.status)
        seen.add(int(st[0]))
        if r.status == 0:
            assert abs(obj[0] + r.fun) <= 1e-7 * (1 + abs(r.fun))
            assert (A @ x[0] - b).max() <= 1e-7 and x[0].min() >= -1e-9
    assert seen == {0, 1, 2}


def test_lp_solve_batch_shared_pattern():
    """S programs in one launch (one CTA each): same pattern, different values / costs / right-hand sides."""
    eng = P._engine()
    rng = np.random.default_rng(1)
    m, n, S = 60, 400, 48
    A = sp.vstack([sp.random(m - 1, n, density=0.05, random_state=3, data_rvs=lambda k: rng.uniform(0.1, 2, k)),
                   sp.csr_matrix(np.ones((1, n)))]).tocsc()       # the last row bounds every column
    A.sort_indices()
    val = A.data[None, :] * rng.uniform(0.5, 1.5, (S, A.nnz))
    c = rng.uniform(0, 1, (S, n))
    b = rng.uniform(1, 5, (S, m))
    x, obj, st, stats = eng.lp_solve(A.indptr, A.indices, val, c, b)
    assert np.all(st == 0)
    for s in range(0, S, 5):
        As = sp.csc_matrix((val[s], A.indices, A.indptr), shape=(m, n))
        r = linprog(-c[s], A_ub=As, b_ub=b[s], bounds=(0, None), method="highs")
        assert r.status == 0 and abs(obj[s] + r.fun) <= 1e-8 * (1 + abs(r.fun))
        assert (As @ x[s] - b[s]).max() <= 1e-8


def test_lp_solve_rejects_bad_arguments():
    eng = P._engine()
    with pytest.raises(RuntimeError):
        eng.lp_solve(np.array([0, 1, 1]), np.array([7]), np.array([1.0]), np.ones(2), np.ones(3))   # row 7 >= m
    with pytest.raises(RuntimeError):
        eng.lp_solve(np.array([0, 1, 0]), np.array([0]), np.array([1.0]), np.ones(2), np.ones(3))   # colp decreasing


@pytest.mark.parametrize("ns,pf", [(6, 1.0), (24, 1.0), (40, 0.5), (64, 1.0)])
def test_packing_policies(ns, pf):
    thr, sf, prio, t0, steps, spec, singles = instance(ns, SPEC, seed=100 + ns, pair_fraction=pf)
    z, _, _ = gp.max_min_fairness_packed(thr, sf, prio, spec)
    pol = P.get_policy("max_min_fairness_packed", solver="ECOS")
    assert pol.name == "MaxMinFairness_Packing"
    _base_ok(pol.get_allocation(thr, sf, prio, spec), thr, sf, spec)
    assert abs(pol.last_objective - z) <= 1e-6 * z

    T, _, _ = gp.min_total_duration_packed(thr, sf, steps, spec)
    pol = P.get_policy("min_total_duration_packed", solver="ECOS")
    x = _base_ok(pol.get_allocation(thr, sf, steps, spec), thr, sf, spec)
    assert pol.last_objective == T
    Pk = gp.Packed(thr, sf, spec)
    for i, s in enumerate(singles):
        assert Pk.coef(i) @ x.ravel() >= steps[s] / T * (1 - 1e-7)

    pol = P.get_policy("finish_time_fairness_packed", solver="ECOS")
    cum = {s: 0.0 for s in singles}
    steps_now = dict(steps)
    for rnd in range(2):
        if rnd == 1:
            Pp = gp.Packed(thr, sf, spec, prio)
            iso = gp.isolated_throughputs(Pp.thr_single, Pp.sf_single, Pp.N)
            steps_next = {s: steps_now[s] * 0.9 for s in singles}
            for i, s in enumerate(singles):
                cum[s] += (steps_now[s] - steps_next[s]) / iso[i]
            steps_now = steps_next
        rho, _, _ = gp.finish_time_fairness_packed(thr, sf, prio, t0, steps_now, cum, spec)
        _base_ok(pol.get_allocation(thr, sf, prio, t0, steps_now, spec), thr, sf, spec)
        assert abs(pol.last_objective - rho) <= 1e-6 * rho, (rnd, pol.last_objective, rho)
        assert pol.last_passes <= 12

    slo = {singles[i]: steps[singles[i]] / (thr[singles[i]]["k80"] * 0.1) for i in range(0, ns, 5)}
    o, _, used, _ = gp.max_sum_throughput_packed_slos(thr, sf, spec, COSTS, slo, steps)
    pol = P.get_policy("max_sum_throughput_normalized_by_cost_packed_SLOs", solver="ECOS")
    _base_ok(pol.get_allocation(thr, sf, spec, COSTS, slo, steps), thr, sf, spec)
    assert pol.used_SLOs == used and abs(pol.last_objective - o) <= 1e-6 * o
    slo_bad = {singles[0]: 1e-3}
    o2, _, used2, _ = gp.max_sum_throughput_packed_slos(thr, sf, spec, COSTS, slo_bad, steps)
    pol.get_allocation(thr, sf, spec, COSTS, slo_bad, steps)
    assert not used2 and not pol.used_SLOs and abs(pol.last_objective - o2) <= 1e-6 * o2


def test_packing_degenerate_and_large():
    """Identical jobs (maximally degenerate, Bland's rule engages) and 100 jobs with all 4950 pairs (15 150 columns)."""
    thr, sf, prio, _, _, spec, singles = instance(36, SPEC, seed=7, pair_fraction=1.0)
    for k in thr:
        for w in thr[k]:
            thr[k][w] = [0.6, 0.6] if k.is_pair() else 1.0
    sf = {s: 1 for s in singles}
    prio = {s: 1.0 for s in singles}
    z, _, _ = gp.max_min_fairness_packed(thr, sf, prio, spec)
    pol = pk.MaxMinFairnessPolicyWithPacking("ECOS")
    _base_ok(pol.get_allocation(thr, sf, prio, spec), thr, sf, spec)
    assert abs(pol.last_objective - z) <= 1e-9 * z
    big = {"v100": 24, "p100": 16, "k80": 12}
    thr, sf, prio, _, _, spec, singles = instance(100, big, seed=100, pair_fraction=1.0)
    z, _, _ = gp.max_min_fairness_packed(thr, sf, prio, spec)
    pol.get_allocation(thr, sf, prio, spec)
    assert abs(pol.last_objective - z) <= 1e-8 * z


@pytest.mark.parametrize("J,seed", [(40, 11), (300, 12)])
def test_six_worker_types(J, seed):
    """All six worker types of tacc_throughputs.json live: policies._hetero routes to the general LP (W > 4)."""
    from tests.test_oracle_packed import check_six_types
    check_six_types(P, J, seed)


@pytest.mark.parametrize("ns,pf,seed", [(6, 1.0, 1), (10, 1.0, 2), (20, 0.4, 4), (12, 0.7, 6)])
def test_water_filling_packed(ns, pf, seed):
    """max_min_fairness_water_filling_packed: iteration by iteration against the reference's loop with its LP on HiGHS
    and its bottleneck MILP on scipy.optimize.milp (same number of iterations, same water levels)."""
    thr, sf, prio, _, _, spec, singles = instance(ns, SPEC, seed=seed, pair_fraction=pf)
    x, eff, it, log, _ = gp.water_filling_packed(thr, sf, prio, spec)
    pol = P.get_policy("max_min_fairness_water_filling_packed")
    assert pol.name == "MaxMinFairnessWaterFilling_Packing"
    e2, ids = pol.get_allocation(thr, sf, prio, spec, return_effective_throughputs=True)
    assert ids == singles and pol.last_iterations == it
    assert np.max(np.abs(eff - e2) / np.maximum(eff, 1e-9)) <= 1e-6
    _base_ok(pol.get_allocation(thr, sf, prio, spec), thr, sf, spec)


@pytest.mark.parametrize("n,nt,seed", [(6, 2, 1), (30, 5, 3), (120, 8, 5)])
def test_job_type_formulation(n, nt, seed):
    """max_min_fairness.py:122-316 (job x job-type columns) on swb_lp_solve vs the HiGHS restatement."""
    from tests.packing_fixtures import JobId, job_type_instance
    thr, j2k, sf, prio = job_type_instance(n, nt, SPEC, seed)
    z, _, (job_ids, keys, wts) = gp.max_min_fairness_job_types(thr, j2k, sf, prio, SPEC)
    pol = pk.MaxMinFairnessPolicyWithPacking("ECOS")
    out = pol.get_allocation_using_job_type_throughputs(thr, j2k, sf, prio, SPEC)
    assert abs(pol.last_objective - z) <= 1e-7 * max(1.0, abs(z))
    al = pol.last_job_type_allocation
    for w in wts:
        used = sum(sf[j] * (al[j][w][None] + 0.5 * sum(al[j][w][o] for o in keys)) for j in job_ids)
        assert used <= SPEC[w] + 1e-6
    ref = gp.convert_job_type_allocation(al, j2k, JobId)
    byrepr = {repr(k): v for k, v in out.items()}
    assert set(map(repr, ref)) == set(byrepr)
    for k, row in ref.items():
        for w in wts:
            assert abs(row[w] - byrepr[repr(k)][w]) <= 1e-12
