"""CPU tests of the Eisenberg-Gale oracle (oracle/gavel_lp.py::eisenberg_gale) and of the strategy-proof policy's host
logic (shockwave_b200/policies.py routed to that oracle): KKT conditions of the program of
max_min_fairness_strategy_proof.py:102-123, the closed form of the single-type case, discount factors in (0, 1]."""
import numpy as np

from oracle import gavel_backend as gb
from oracle import gavel_lp as gl


def test_oracle_satisfies_kkt():
    rng = np.random.default_rng(0)
    J, W = 9, 3
    c = rng.uniform(0.5, 5.0, (J, 1)) * rng.uniform(0.2, 1.0, (J, W))
    sf = rng.choice([1.0, 2.0, 4.0], J)
    N = np.array([4.0, 3.0, 2.0])
    x, u = gl.eisenberg_gale(c, sf, N)
    assert x.min() >= 0 and np.all(x.sum(axis=1) <= 1 + 1e-8) and np.all((sf[:, None] * x).sum(axis=0) <= N * (1 + 1e-8))
    # stationarity: c_jw / u_j <= sf_j p_w + lam_j with equality where x_jw > 0; recover (p, lam) by least squares on the
    # active entries and check the inequalities on the others
    act = x > 1e-6
    rows, rhs = [], []
    for j in range(J):
        for w in range(W):
            if act[j, w]:
                r = np.zeros(W + J); r[w] = sf[j]; r[W + j] = 1.0 if x[j].sum() > 1 - 1e-6 else 0.0
                rows.append(r); rhs.append(c[j, w] / u[j])
    sol = np.linalg.lstsq(np.array(rows), np.array(rhs), rcond=None)[0]
    p, lam = sol[:W], sol[W:] * (x.sum(axis=1) > 1 - 1e-6)
    assert p.min() > -1e-6 and lam.min() > -1e-6
    assert np.all(c / u[:, None] <= sf[:, None] * p[None] + lam[:, None] + 1e-5)


def test_single_type_closed_form():
    """One worker type, nobody time-limited: equal budgets -> x_j = N / (J sf_j)."""
    J, N = 8, np.array([4.0])
    sf = np.array([1.0, 2.0, 4.0, 1.0, 1.0, 2.0, 1.0, 4.0])
    c = np.random.default_rng(1).uniform(1, 3, (J, 1))
    x, u = gl.eisenberg_gale(c, sf, N)
    assert np.allclose(x[:, 0], N[0] / (J * sf), rtol=1e-6)


def test_policy_host_logic_on_the_oracle_backend():
    rng = np.random.default_rng(3)
    J = 6
    thr = {j: {"k80": float(a * 0.3), "p100": float(a * 0.6), "v100": float(a)} for j, a in enumerate(rng.uniform(1, 5, J))}
    thr[2]["k80"] *= 2.0                                        # one job that likes the slow type
    sf = {j: int(v) for j, v in enumerate(rng.choice([1, 2, 4], J))}
    prio = {j: 1.0 for j in range(J)}
    spec = {"k80": 3, "p100": 2, "v100": 2}
    with gb.cpu_backend() as P:
        pol = P.MaxMinFairnessStrategyProofPolicyWithPerf(solver="ECOS")
        alloc, disc = pol.get_allocation(thr, sf, prio, spec)
        thr_only = pol.get_allocation(thr, sf, prio, spec, recurse_deeper=False)
    assert pol.name == "MaxMinFairness_Perf" and set(alloc) == set(thr) and len(disc) == J
    assert np.all(disc > 0) and np.all(disc <= 1 + 1e-6)        # removing a competitor never hurts the others
    a = np.array([[alloc[j][w] for w in sorted(spec)] for j in range(J)])
    assert a.min() >= 0 and np.all(a.sum(axis=1) <= 1 + 1e-6)
    N = np.array([spec[w] for w in sorted(spec)], float)
    assert np.all((np.array([sf[j] for j in range(J)])[:, None] * a).sum(axis=0) <= N * (1 + 1e-6))
    full = np.array([[thr[j][w] for w in sorted(spec)] for j in range(J)])
    assert np.allclose([thr_only[j] for j in range(J)], (full * (a / disc[:, None])).sum(axis=1), rtol=1e-5)
