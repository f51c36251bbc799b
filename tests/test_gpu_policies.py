"""GPU parity of the Gavel policies' get_allocation() against the HiGHS LP oracle (oracle/gavel_lp.py).

LP optima are degenerate in x: parity is on the objective (1e-6 relative), on the base
constraints (policy.py:58-65, to 1e-9) and on the effective throughput of the bottleneck jobs."""
import numpy as np
import pytest

from oracle import gavel_lp as gl
from shockwave_b200 import policies as P

pytestmark = pytest.mark.gpu
WT = ["k80", "p100", "v100"]          # sorted worker-type order of Policy.flatten


def _instance(J, spec, seed, equal_columns=True):
    rng = np.random.default_rng(seed)
    col = rng.uniform(0.5, 20.0, size=J)
    thr = {j: {w: (col[j] if equal_columns else col[j] * rng.uniform(0.3, 1.0)) for w in WT} for j in range(J)}
    sf = {j: int(rng.choice([1, 2, 4, 8], p=[0.6, 0.3, 0.09, 0.01])) for j in range(J)}
    return thr, sf, dict(spec), rng


def _mat(d, J):
    return np.array([[d[j][w] for w in WT] for j in range(J)])


def _check_base(x, sf, spec, J):
    N = np.array([spec[w] for w in WT], dtype=float)
    s = np.array([sf[j] for j in range(J)], dtype=float)
    assert x.min() >= -1e-12 and x.max() <= 1 + 1e-12
    assert np.all(x.sum(axis=1) <= 1 + 1e-9)
    assert np.all((x * s[:, None]).sum(axis=0) <= N + 1e-9 * np.maximum(1, N))


SPECS = [{"v100": 32, "p100": 0, "k80": 0}, {"v100": 16, "p100": 8, "k80": 8}, {"v100": 512, "p100": 0, "k80": 0}]


@pytest.mark.parametrize("J,spec", [(24, SPECS[0]), (120, SPECS[0]), (96, SPECS[1]), (2048, SPECS[2])])
def test_max_min_fairness(J, spec):
    thr, sf, spec, rng = _instance(J, spec, seed=J)
    prio = {j: float(rng.choice([1.0, 2.0, 5.0])) for j in range(J)}
    pol = P.MaxMinFairnessPolicyWithPerf(solver="ECOS")
    alloc = pol.get_allocation(thr, sf, prio, spec)
    x = _mat(alloc, J)
    _check_base(x, sf, spec, J)
    N = [spec[w] for w in WT]
    z, _ = gl.max_min_fairness_perf(_mat(thr, J), np.array([sf[j] for j in range(J)], float),
                                    np.array([prio[j] for j in range(J)]), N)
    assert abs(pol.last_objective - z) <= 1e-6 * abs(z)
    # every job reaches the max-min level
    pw = (1.0 / np.array([prio[j] for j in range(J)])) / gl.proportional_throughputs(_mat(thr, J), N)
    eff = (_mat(thr, J) * x).sum(axis=1) * pw * np.array([sf[j] for j in range(J)], float)
    assert eff.min() >= z * (1 - 1e-6)
    # the non-Perf class ignores throughputs (max_min_fairness.py:33-37)
    a2 = P.MaxMinFairnessPolicy(solver="ECOS").get_allocation(thr, sf, prio, spec)
    _check_base(_mat(a2, J), sf, spec, J)


@pytest.mark.parametrize("J,spec", [(24, SPECS[0]), (120, SPECS[0]), (96, SPECS[1]), (1024, SPECS[2])])
def test_finish_time_fairness_stateful(J, spec):
    thr, sf, spec, rng = _instance(J, spec, seed=100 + J)
    prio = {j: 1.0 for j in range(J)}
    pol = P.FinishTimeFairnessPolicyWithPerf(solver="GUROBI")
    steps = {j: float(rng.uniform(1e4, 1e6)) for j in range(J)}
    t = {j: float(rng.uniform(0, 5e3)) for j in range(J)}
    cum = np.zeros(J)
    N = [spec[w] for w in WT]
    s = np.array([sf[j] for j in range(J)], float)
    prev_steps, prev_iso = None, None
    for it in range(3):        # three allocation rounds: cumulative isolated time evolves
        alloc = pol.get_allocation(thr, sf, prio, t, steps, spec)
        x = _mat(alloc, J)
        _check_base(x, sf, spec, J)
        if prev_steps is not None:
            cum += (prev_steps - np.array([steps[j] for j in range(J)])) / prev_iso
        rho, _, den = gl.finish_time_fairness_perf(_mat(thr, J), s, np.array([t[j] for j in range(J)]),
                                                   np.array([steps[j] for j in range(J)]), cum, N)
        assert abs(pol.last_objective - rho) <= 1e-6 * rho
        eff = (_mat(thr, J) * x).sum(axis=1)
        ratio = (np.array([t[j] for j in range(J)]) + np.array([steps[j] for j in range(J)]) / eff) / den
        assert ratio.max() <= rho * (1 + 1e-6)
        prev_steps = np.array([steps[j] for j in range(J)])
        prev_iso = (_mat(thr, J) * gl.isolated_allocation(_mat(thr, J), s, N)).sum(axis=1)
        for j in range(J):
            steps[j] *= 0.8
            t[j] += 360.0
    assert P.FinishTimeFairnessPolicy(solver="GUROBI").get_allocation({}, {}, {}, {}, {}, spec) is None


@pytest.mark.parametrize("J,spec", [(24, SPECS[0]), (120, SPECS[0]), (96, SPECS[1])])
def test_min_total_duration_and_max_sum(J, spec):
    thr, sf, spec, rng = _instance(J, spec, seed=200 + J)
    steps = {j: float(rng.uniform(1e3, 1e6)) for j in range(J)}
    N = [spec[w] for w in WT]
    s = np.array([sf[j] for j in range(J)], float)
    pol = P.MinTotalDurationPolicyWithPerf(solver="ECOS")
    x = _mat(pol.get_allocation(thr, sf, steps, spec), J)
    _check_base(x, sf, spec, J)
    T, _ = gl.min_total_duration_perf(_mat(thr, J), s, np.array([steps[j] for j in range(J)]), N)
    assert pol.last_objective == T                       # same bisection sequence as the reference
    assert np.all((_mat(thr, J) * x).sum(axis=1) >= np.array([steps[j] for j in range(J)]) / T * (1 - 1e-9))
    pol = P.ThroughputSumWithPerf(solver="ECOS")
    x = _mat(pol.get_allocation(thr, sf, spec), J)
    _check_base(x, sf, spec, J)
    v, _ = gl.max_sum_throughput(_mat(thr, J), s, N)
    assert abs((_mat(thr, J) * x).sum() - v) <= 1e-6 * v


def test_closed_forms_and_edge_cases():
    thr, sf, spec, rng = _instance(10, SPECS[1], seed=3)
    N = np.array([spec[w] for w in WT], float)
    s = np.array([sf[j] for j in range(10)], float)
    iso = _mat(P.IsolatedPolicy().get_allocation(thr, sf, spec), 10)
    assert np.allclose(iso, gl.isolated_allocation(_mat(thr, 10), s, N), rtol=1e-12)
    prop = _mat(P.ProportionalPolicy().get_allocation(thr, spec), 10)
    assert np.allclose(prop, np.tile(N / N.sum(), (10, 1)))
    gf = _mat(P.GandivProportionalPolicy().get_allocation(thr, sf, spec), 10)
    want = np.tile(N / 10, (10, 1)); want = want / np.maximum(want.sum(axis=1), 1.0)[:, None]
    assert np.allclose(gf, want, rtol=1e-12)
    ip = _mat(P.get_policy("isolated_plus").get_allocation(thr, sf, spec), 10)     # isolated_plus.py:36-55: same closed form
    assert np.allclose(ip, want, rtol=1e-12) and P.get_policy("isolated_plus").name == "Isolated_plus"
    assert P.MaxMinFairnessPolicy(solver=None).get_allocation({}, {}, {}, spec) is None
    assert P.get_policy("max_min_fairness").name == "MaxMinFairness"
    assert P.get_policy("finish_time_fairness").name.startswith("FinishTimeFairness")
    assert P.get_policy("shockwave").name == "shockwave"


def test_allox_assignment_and_policy(engine):
    """AlloX: the GPU shortest-augmenting-path solver reaches the same optimal cost as the reference's
    scipy.optimize.linear_sum_assignment on the implicit q matrix (allox.py:108-144), with a valid assignment;
    the policy (stateful _prev_allocation) returns a valid allocation round after round."""
    rng = np.random.default_rng(11)
    same_alloc = total = 0
    for trial in range(12):
        m = int(rng.integers(1, 60)); n = int(rng.integers(1, 40)); W = int(rng.integers(1, 4))
        p = rng.uniform(10.0, 1e5, size=(m, W)); t = rng.uniform(0.0, 1e4, size=m)
        wtype = np.sort(rng.integers(0, W, size=n)).astype(np.int32)
        cols, cost = engine.allox_assign(p, t, wtype)
        ref_cols, ref_cost = gl.allox_assignment(p, t, wtype)
        assert len(set(cols.tolist())) == m and cols.min() >= 0 and cols.max() < m * n
        q = gl.allox_q_matrix(p, t, wtype)
        assert abs(q[np.arange(m), cols].sum() - cost) <= 1e-9 * cost
        assert abs(cost - ref_cost) <= 1e-9 * ref_cost, (trial, m, n, W, cost, ref_cost)
    # policy level, three consecutive calls with jobs finishing in between
    J, spec = 40, {"v100": 6, "p100": 4, "k80": 3}
    thr = {j: {"v100": float(rng.uniform(5, 20)), "p100": float(rng.uniform(2, 10)), "k80": float(rng.uniform(0.5, 5))}
           for j in range(J)}
    sf = {j: 1 for j in range(J)}
    tss = {j: float(rng.uniform(0, 5000)) for j in range(J)}
    steps = {j: float(rng.uniform(1e3, 1e6)) for j in range(J)}
    pol = P.AlloXPolicy(alpha=0.2)
    prev = {}
    live = list(range(J))
    for it in range(3):
        sub = {j: thr[j] for j in live}
        alloc = pol.get_allocation(sub, sf, tss, steps, [], spec)
        want = gl.allox_allocation(sorted(sub), WT, sub, sf, tss, steps, spec, prev, 0.2)
        prev = {j: dict(v) for j, v in want.items()}
        for w in WT:
            assert sum(alloc[j][w] for j in live) <= spec[w] + 1e-9
        assert all(sum(alloc[j].values()) <= 1 + 1e-9 for j in live)
        total += 1
        same_alloc += int(all(alloc[j] == want[j] for j in live))
        pol._prev_allocation = {j: dict(v) for j, v in want.items()}     # keep both histories identical
        live = [j for j in live if rng.random() > 0.2]
    print("AlloX allocations identical to the scipy-backed restatement:", same_alloc, "/", total)
    assert P.get_policy("allox").name == "AlloX_Perf"


@pytest.mark.parametrize("J,N,kinds", [(12, 32.0, 0), (120, 32.0, 0), (60, 48.0, 3), (700, 512.0, 0), (2048, 512.0, 21), (3000, 300.0, 6)])
def test_pooled_selection_is_the_analytic_centre(J, N, kinds):
    """Which optimal x: the interior-point selection of the reference's solvers (oracle/gavel_lp.py, pinned closed-loop
    in tests/golden/tacc32_policy_pins.json).  policy.cu against the numpy restatement, every mode, to 1e-6; the
    objective against the HiGHS LP to 1e-7."""
    from oracle import gavel_backend as gb
    rng = np.random.default_rng(J)
    thr = rng.uniform(0.5, 20.0, J) if not kinds else rng.uniform(0.5, 20.0, kinds)[rng.integers(0, kinds, J)]
    sf = rng.choice([1.0, 2.0, 4.0, 8.0], J, p=[0.6, 0.3, 0.09, 0.01])
    pw = rng.choice([1.0, 2.0, 5.0], J)
    n = rng.uniform(1e4, 1e6, J)
    t = rng.uniform(0, 5e3, J)
    iso = thr * np.minimum(1.0, (N / J) / sf)
    den = rng.uniform(0, 2e3, J) + n / iso
    cases = [(P.POL_MAXMIN, dict(coef=sf / pw)), (P.POL_FTF, dict(coef=thr, t=t, n=n, den=den)),
             (P.POL_MTD, dict(coef=thr, n=n)), (P.POL_MAXSUM, dict(coef=thr))]
    for mode, kw in cases:
        coef = kw.pop("coef")
        x, obj, rc = P._pooled(mode, N, coef, sf, **kw)
        xo, oo, rco = gb.pooled_cpu(mode, N, coef, sf, **kw)
        assert rc == rco == 0
        assert abs(obj - oo) <= 1e-7 * abs(oo), (mode, obj, oo)      # HiGHS' own feasibility tolerance is 1e-7
        assert np.abs(x - xo).max() <= 1e-6, (mode, np.abs(x - xo).max())
        assert x.min() >= 0 and x.max() <= 1 + 1e-12 and (sf * x).sum() <= N * (1 + 1e-9)


def test_pooled_random_sweep_against_oracle_centre():
    """36 random small pooled instances, capacity from scarce to ample (everything fits), J from 1: objective and the
    selected x against the oracle (HiGHS objective + analytic-centre selection), every mode."""
    from oracle import gavel_backend as gb
    rng = np.random.default_rng(5)
    worst = 0.0
    for it in range(36):
        J = int(rng.integers(1, 60))
        thr = rng.uniform(0.5, 20.0, J) if it % 3 else rng.uniform(0.5, 20.0, 3)[rng.integers(0, 3, J)]
        sf = rng.choice([1.0, 2.0, 4.0, 8.0], J)
        N = float(max(1, round(sf.sum() * rng.choice([0.1, 0.4, 0.9, 1.0, 1.5, 4.0]))))
        pw = rng.choice([1.0, 2.0, 5.0], J)
        n = rng.uniform(1e4, 1e6, J)
        t = rng.uniform(0, 5e3, J)
        den = rng.uniform(0, 2e3, J) + n / (thr * np.minimum(1.0, (N / J) / sf))
        for mode, kw in ((P.POL_MAXMIN, dict(coef=sf / pw)), (P.POL_FTF, dict(coef=thr, t=t, n=n, den=den)),
                         (P.POL_MTD, dict(coef=thr, n=n)), (P.POL_MAXSUM, dict(coef=thr))):
            coef = kw.pop("coef")
            x, obj, rc = P._pooled(mode, N, coef, sf, **kw)
            xo, oo, rco = gb.pooled_cpu(mode, N, coef, sf, **kw)
            assert rc == rco == 0, (it, mode)
            assert abs(obj - oo) <= 1e-7 * abs(oo), (it, mode, obj, oo)
            assert np.abs(x - xo).max() <= 2e-6, (it, mode, J, N, np.abs(x - xo).max())
            assert x.min() >= 0 and x.max() <= 1 + 1e-12 and (sf * x).sum() <= N * (1 + 1e-9)
            worst = max(worst, np.abs(x - xo).max())
    print("pooled random sweep: worst |x - x_oracle|", worst)


@pytest.mark.parametrize("J,spec,kinds", [(20, SPECS[0], 0), (120, SPECS[0], 3), (96, SPECS[1], 0), (600, SPECS[2], 5)])
def test_max_sum_with_slos(J, spec, kinds):
    """ThroughputNormalizedByCostSum_PerfSLOs (max_sum_throughput.py:49-108): SLO rows as floors on x, the reference's
    fall-back to the SLO-free program when they do not fit, objective vs the HiGHS LP, x vs the oracle selection."""
    from oracle import gavel_backend as gb
    rng = np.random.default_rng(4000 + J)
    col = rng.uniform(0.5, 20.0, J) if not kinds else rng.uniform(0.5, 20.0, kinds)[rng.integers(0, kinds, J)]
    thr = {j: {w: float(col[j]) for w in WT} for j in range(J)}
    sf = {j: int(rng.choice([1, 2, 4, 8], p=[0.6, 0.3, 0.09, 0.01])) for j in range(J)}
    steps = {j: float(rng.uniform(1e3, 1e5)) for j in range(J)}
    N = np.array([spec[w] for w in WT], float)
    s = np.array([sf[j] for j in range(J)], float)
    costs = {"k80": 1.5, "p100": 1.5, "v100": 1.5}
    # SLOs that fit: a fifth of the jobs must finish within 2-20x their full-speed time, scaled so the floors use < 60 % of the cluster
    slo_jobs = [j for j in range(J) if rng.random() < 0.2] or [0]
    slos = {j: steps[j] / col[j] * float(rng.uniform(2.0, 20.0)) for j in slo_jobs}
    lo = np.zeros(J)
    for j in slo_jobs:
        lo[j] = steps[j] / slos[j] / col[j]
    scale = max(1.0, (s * lo).sum() / (0.6 * N.sum()))
    slos = {j: v * scale for j, v in slos.items()}
    lo = lo / scale
    pol = P.ThroughputNormalizedByCostSumWithPerfSLOs(solver="ECOS")
    x = _mat(pol.get_allocation(thr, sf, dict(spec), instance_costs=costs, SLOs=slos, num_steps_remaining=steps), J)
    _check_base(x, sf, spec, J)
    xj = x.sum(axis=1)
    assert np.all(xj >= lo * (1 - 1e-9))
    v, _ = gl.max_sum_throughput((col / 1.5)[:, None], s, [N.sum()], need=lo * col / 1.5)
    assert abs(pol.last_objective - v) <= 1e-7 * v
    xo, _, rco = gb.pooled_cpu(P.POL_MAXSUM, N.sum(), col / 1.5, s, t=lo)
    assert rco == 0 and np.abs(xj - xo).max() <= 1e-6, np.abs(xj - xo).max()
    # SLOs that cannot be met: the reference re-solves without them
    hard = {j: slos[j] * 1e-6 for j in slo_jobs}
    x2 = _mat(pol.get_allocation(thr, sf, dict(spec), instance_costs=costs, SLOs=hard, num_steps_remaining=steps), J)
    x3 = _mat(P.ThroughputNormalizedByCostSumWithPerf(solver="ECOS").get_allocation(thr, sf, dict(spec), costs), J)
    assert np.allclose(x2, x3)
