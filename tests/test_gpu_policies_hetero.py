"""GPU parity of the HETEROGENEOUS Gavel *_Perf policies (swb_policy_hetero, hetero.cu) against the HiGHS LP oracle.

Worker types k80 / p100 / v100 with different per-job throughputs (the general case of
scheduler/policies/{max_min_fairness,finish_time_fairness,min_total_duration,max_sum_throughput}.py).
Parity: objective within 1e-6 relative of the LP optimum, base constraints (policy.py:58-65)
to 1e-9, every job's requirement met."""
import numpy as np
import pytest

from oracle import gavel_lp as gl
from shockwave_b200 import policies as P

pytestmark = pytest.mark.gpu
WT = ["k80", "p100", "v100"]


def _instance(J, spec, seed, kinds=0, zero=False):
    rng = np.random.default_rng(seed)
    if kinds:       # few job types (the reference's throughput table has 21 templates): massive ties
        tab = np.sort(rng.uniform(0.5, 20.0, size=(kinds, 3)), axis=1)
        m = tab[rng.integers(0, kinds, J)]
    else:
        m = rng.uniform(0.5, 20.0, size=(J, 1)) * np.sort(rng.uniform(0.1, 1.0, size=(J, 3)), axis=1)
    if zero:        # some jobs cannot run on the slowest type at all
        m[rng.integers(0, J, max(1, J // 8)), 0] = 0.0
    thr = {j: {w: float(m[j, i]) for i, w in enumerate(WT)} for j in range(J)}
    sf = {j: int(rng.choice([1, 2, 4, 8], p=[0.6, 0.3, 0.09, 0.01])) for j in range(J)}
    return thr, sf, dict(spec), rng, m


def _mat(d, J):
    return np.array([[d[j][w] for w in WT] for j in range(J)])


def _check_base(x, sf, spec, J):
    N = np.array([spec[w] for w in WT], dtype=float)
    s = np.array([sf[j] for j in range(J)], dtype=float)
    assert x.min() >= -1e-12 and x.max() <= 1 + 1e-12
    assert np.all(x.sum(axis=1) <= 1 + 1e-9)
    assert np.all((x * s[:, None]).sum(axis=0) <= N + 1e-9 * np.maximum(1, N))
    assert np.all(x[:, N == 0] == 0)


SPECS = [{"v100": 16, "p100": 8, "k80": 8}, {"v100": 4, "p100": 0, "k80": 12}, {"v100": 36, "p100": 36, "k80": 36},
         {"v100": 128, "p100": 128, "k80": 256}]
CASES = [(7, SPECS[0], 0, False), (40, SPECS[0], 0, True), (60, SPECS[1], 0, False), (96, SPECS[2], 4, False),
         (150, SPECS[0], 3, True), (400, SPECS[0], 0, False), (400, SPECS[0], 5, True), (2048, SPECS[3], 0, False),
         (2048, SPECS[3], 21, False)]


@pytest.mark.parametrize("J,spec,kinds,zero", CASES)
def test_max_min_fairness_hetero(J, spec, kinds, zero):
    thr, sf, spec, rng, m = _instance(J, spec, 1000 + J, kinds, zero)
    prio = {j: float(rng.choice([1.0, 2.0, 5.0])) for j in range(J)}
    pol = P.MaxMinFairnessPolicyWithPerf(solver="ECOS")
    x = _mat(pol.get_allocation(thr, sf, prio, spec), J)
    _check_base(x, sf, spec, J)
    N = np.array([spec[w] for w in WT], float)
    s = np.array([sf[j] for j in range(J)], float)
    pr = np.array([prio[j] for j in range(J)])
    live = N > 0
    z, _ = gl.max_min_fairness_perf(m[:, live], s, pr, N[live])
    # the reference computes the proportional throughputs over ALL types (zero-capacity ones contribute 0)
    pw = (1.0 / pr) / gl.proportional_throughputs(m[:, live], N[live])
    assert abs(pol.last_objective - z) <= 1e-6 * abs(z), (pol.last_objective, z, P._hetero.last_stats)
    eff = (m * x).sum(axis=1) * pw * s
    assert eff.min() >= z * (1 - 1e-6)
    import time
    t0 = time.perf_counter(); pol.get_allocation(thr, sf, prio, spec); dt = time.perf_counter() - t0
    t0 = time.perf_counter(); gl.max_min_fairness_perf(m[:, live], s, pr, N[live]); dto = time.perf_counter() - t0
    print("max-min J", J, "objective", pol.last_objective, "lp", z, "pricing passes / checks", P._hetero.last_stats,
          "get_allocation %.2f ms, HiGHS LP %.1f ms" % (dt * 1e3, dto * 1e3))


@pytest.mark.parametrize("J,spec,kinds,zero", CASES[:7])
def test_finish_time_fairness_hetero(J, spec, kinds, zero):
    thr, sf, spec, rng, m = _instance(J, spec, 2000 + J, kinds, zero)
    prio = {j: 1.0 for j in range(J)}
    pol = P.FinishTimeFairnessPolicyWithPerf(solver="GUROBI")
    steps = {j: float(rng.uniform(1e4, 1e6)) for j in range(J)}
    t = {j: float(rng.uniform(0, 5e3)) for j in range(J)}
    N = np.array([spec[w] for w in WT], float)
    live = N > 0
    s = np.array([sf[j] for j in range(J)], float)
    cum = np.zeros(J)
    prev_steps = prev_iso = None
    for it in range(2):
        x = _mat(pol.get_allocation(thr, sf, prio, t, steps, spec), J)
        _check_base(x, sf, spec, J)
        st = np.array([steps[j] for j in range(J)]); tt = np.array([t[j] for j in range(J)])
        if prev_steps is not None:
            cum += (prev_steps - st) / prev_iso
        rho, _, den = gl.finish_time_fairness_perf(m[:, live], s, tt, st, cum, N[live], tol=1e-9)
        assert abs(pol.last_objective - rho) <= 1e-6 * rho, (pol.last_objective, rho)
        eff = (m * x).sum(axis=1)
        assert ((tt + st / eff) / den).max() <= rho * (1 + 1e-6)
        prev_steps = st
        prev_iso = (m[:, live] * gl.isolated_allocation(m[:, live], s, N[live])).sum(axis=1)
        for j in range(J):
            steps[j] *= 0.8
            t[j] += 360.0


@pytest.mark.parametrize("J,spec,kinds,zero", CASES[:7])
def test_min_total_duration_and_max_sum_hetero(J, spec, kinds, zero):
    thr, sf, spec, rng, m = _instance(J, spec, 3000 + J, kinds, zero)
    steps = {j: float(rng.uniform(1e3, 1e6)) for j in range(J)}
    N = np.array([spec[w] for w in WT], float)
    live = N > 0
    s = np.array([sf[j] for j in range(J)], float)
    st = np.array([steps[j] for j in range(J)])
    pol = P.MinTotalDurationPolicyWithPerf(solver="ECOS")
    x = _mat(pol.get_allocation(thr, sf, steps, spec), J)
    _check_base(x, sf, spec, J)
    T, _ = gl.min_total_duration_perf(m[:, live], s, st, N[live])
    assert pol.last_objective == T                       # same bisection sequence as the reference
    assert np.all((m * x).sum(axis=1) >= st / T * (1 - 1e-9))
    # max-sum, with and without per-type instance costs (max_sum_throughput.py:72-86)
    for costs in (None, {"k80": 0.9, "p100": 1.46, "v100": 2.48}):
        pol = P.ThroughputNormalizedByCostSumWithPerf(solver="ECOS") if costs else P.ThroughputSumWithPerf(solver="ECOS")
        alloc = pol.get_allocation(thr, sf, spec, costs) if costs else pol.get_allocation(thr, sf, spec)
        x = _mat(alloc, J)
        _check_base(x, sf, spec, J)
        c = np.array([costs[w] for w in WT]) if costs else np.ones(3)
        v, _ = gl.max_sum_throughput(m[:, live], s, N[live], costs=c[live])
        got = ((m / c[None, :]) * x).sum()
        assert abs(got - v) <= 1e-6 * v, (got, v, P._hetero.last_stats)
        print("max-sum J", J, "costs", bool(costs), "value", got, "lp", v, "passes / checks", P._hetero.last_stats)


def test_random_sweep_hetero():
    """80 random small instances straight through swb_policy_hetero (W = 2 and 3, capacities from scarce to ample,
    duplicated jobs, zero throughputs, wide gangs): max-min and max-sum objectives against the HiGHS LPs."""
    rng = np.random.default_rng(77)
    worst = 0.0
    for it in range(80):
        W = int(rng.integers(2, 4))
        J = int(rng.integers(2, 70))
        a = rng.uniform(0.2, 10.0, size=(J, 1)) * rng.uniform(0.05, 1.0, size=(J, W))
        if it % 3 == 0:
            a = a[rng.integers(0, max(1, J // 6), J)]           # few distinct rows: ties everywhere
        if it % 4 == 1:
            a[rng.integers(0, J, max(1, J // 5)), rng.integers(0, W)] = 0.0
        a[a.max(axis=1) == 0.0, 0] = 1.0
        sf = rng.choice([1.0, 2.0, 4.0, 8.0], J)
        N = rng.integers(1, 4 * J + 2, W).astype(float) * rng.choice([0.25, 1.0, 4.0])
        N = np.maximum(np.round(N), 1.0)
        x, obj, rc = P._hetero(P.POL_MAXMIN, N, a, sf)
        z, _ = gl.max_min(a, sf, N)
        assert rc == 0 and abs(obj - z) <= 1e-6 * abs(z), (it, J, W, obj, z)
        assert x.min() >= -1e-12 and np.all(x.sum(axis=1) <= 1 + 1e-9)
        assert np.all((x * sf[:, None]).sum(axis=0) <= N * (1 + 1e-9))
        assert (a * x).sum(axis=1).min() >= z * (1 - 1e-6)
        worst = max(worst, abs(obj - z) / abs(z))
        x, obj, rc = P._hetero(P.POL_MAXSUM, N, a, sf)
        v, _ = gl.max_sum_throughput(a, sf, N)
        assert rc == 0 and abs(obj - v) <= 1e-6 * abs(v), (it, J, W, obj, v)
        assert x.min() >= -1e-12 and np.all(x.sum(axis=1) <= 1 + 1e-9)
        assert np.all((x * sf[:, None]).sum(axis=0) <= N * (1 + 1e-9))
        assert abs((a * x).sum() - v) <= 1e-6 * v
        worst = max(worst, abs(obj - v) / abs(v))
    print("hetero random sweep: worst relative objective difference vs HiGHS", worst)


def test_four_worker_types():
    """W = 4 (one more type than Gavel's k80 / p100 / v100) for the three programs with a per-job requirement."""
    rng = np.random.default_rng(91)
    for it in range(12):
        J = int(rng.integers(3, 60))
        a = rng.uniform(0.2, 10.0, size=(J, 1)) * rng.uniform(0.05, 1.0, size=(J, 4))
        sf = rng.choice([1.0, 2.0, 4.0], J)
        N = np.maximum(np.round(rng.integers(1, 2 * J + 2, 4) * rng.choice([0.25, 1.0])), 1.0)
        x, obj, rc = P._hetero(P.POL_MAXMIN, N, a, sf)
        z, _ = gl.max_min(a, sf, N)
        assert rc == 0 and abs(obj - z) <= 1e-6 * abs(z), (it, obj, z)
        assert np.all(x.sum(axis=1) <= 1 + 1e-9) and np.all((x * sf[:, None]).sum(axis=0) <= N * (1 + 1e-9))
        n = rng.uniform(1e3, 1e5, J)
        x, T, rc = P._hetero(P.POL_MTD, N, a, sf, n=n)
        To, _ = gl.min_total_duration_perf(a, sf, n, N)
        assert rc == 0 and T == To
        assert np.all((a * x).sum(axis=1) >= n / T * (1 - 1e-9))


@pytest.mark.parametrize("J,spec,kinds", [(12, SPECS[0], 0), (60, SPECS[0], 4), (90, SPECS[1], 0), (300, SPECS[3], 0)])
def test_max_sum_with_slos_hetero(J, spec, kinds):
    """SLO rows (max_sum_throughput.py:87-93) with different per-type throughputs and per-type instance costs."""
    thr, sf, spec, rng, m = _instance(J, spec, 5000 + J, kinds, False)
    N = np.array([spec[w] for w in WT], float)
    live = N > 0
    s = np.array([sf[j] for j in range(J)], float)
    costs = {"k80": 0.9, "p100": 1.46, "v100": 2.48}
    c = np.array([costs[w] for w in WT])
    steps = {j: float(rng.uniform(1e3, 1e5)) for j in range(J)}
    slo_jobs = [j for j in range(J) if rng.random() < 0.25] or [0]
    need = np.zeros(J)
    for j in slo_jobs:
        need[j] = m[j, live].max() * float(rng.uniform(0.05, 0.6))
    # keep the floors inside ~half of the cluster so that they fit
    while (s * need / m[:, live].max(axis=1)).sum() > 0.5 * N.sum():
        need *= 0.7
    slos = {j: steps[j] / need[j] for j in slo_jobs}
    pol = P.ThroughputNormalizedByCostSumWithPerfSLOs(solver="ECOS")
    x = _mat(pol.get_allocation(thr, sf, spec, instance_costs=costs, SLOs=slos, num_steps_remaining=steps), J)
    _check_base(x, sf, spec, J)
    assert np.all((m * x).sum(axis=1) >= need * (1 - 1e-9))
    v, _ = gl.max_sum_throughput(m[:, live], s, N[live], costs=c[live], need=need)
    got = ((m / c[None, :]) * x).sum()
    assert v is not None and abs(got - v) <= 1e-6 * v, (got, v, P._hetero.last_stats)
    assert abs(pol.last_objective - v) <= 1e-6 * v
    # floors that cannot be met: fall back to the SLO-free program, like the reference
    hard = {j: 1e-9 for j in slo_jobs}
    x2 = _mat(pol.get_allocation(thr, sf, spec, instance_costs=costs, SLOs=hard, num_steps_remaining=steps), J)
    v0, _ = gl.max_sum_throughput(m[:, live], s, N[live], costs=c[live])
    assert abs(((m / c[None, :]) * x2).sum() - v0) <= 1e-6 * v0
