"""GPU parity of MaxMinFairnessStrategyProofPolicyWithPerf (scheduler/policies/max_min_fairness_strategy_proof.py:47-155):
its J + 1 Eisenberg-Gale programs solved as ONE batch of scenarios by the dense price-response kernel (market.cu, log
utility) against a scipy trust-constr restatement of the cvxpy program (oracle/gavel_lp.py::eisenberg_gale).  The
reference ships no golden or known-answer test for this policy and cvxpy cannot be installed: parity unpinned at the
value level; compared are the quantities that are UNIQUE at the optimum — per-job utilities / throughputs and the
discount factors built from them — plus the base constraints (policy.py:58-65) of the returned allocation."""
import numpy as np
import pytest

from oracle import gavel_backend as gb
from oracle import gavel_lp as gl
from shockwave_b200 import policies as P

pytestmark = pytest.mark.gpu
WT = ["k80", "p100", "v100"]


def _instance(J, N, seed, mode="random"):
    rng = np.random.default_rng(seed)
    base = rng.uniform(0.5, 20.0, size=(J, 1))
    if mode == "pooled":
        thr = np.repeat(base, 3, axis=1)
    elif mode == "proportional":                       # every job has the same relative speeds: all types tie
        thr = base * np.array([[0.3, 0.6, 1.0]])
    else:
        thr = base * np.sort(rng.uniform(0.1, 1.0, size=(J, 3)), axis=1)
    sf = rng.choice([1, 2, 4], J, p=[0.6, 0.3, 0.1])
    prio = rng.choice([1.0, 2.0, 0.5], J)
    return thr, sf, prio, np.asarray(N, float)


@pytest.mark.parametrize("J,N,seed,mode", [(6, [3, 2, 2], 1, "random"), (12, [6, 4, 2], 2, "random"),
                                           (24, [8, 8, 4], 3, "random"), (16, [4, 3, 6], 4, "proportional"),
                                           (10, [3, 3, 3], 5, "pooled")])
def test_eisenberg_gale_batch_matches_oracle(engine, J, N, seed, mode):
    """The kernel on S = J + 1 scenarios vs the oracle, scenario by scenario: utilities to 5e-4 relative (the objective
    is flat at the optimum: 5e-4 on a utility is ~1e-7 on sum_j log u_j)."""
    P._shared_engine = engine
    thr, sf, prio, N = _instance(J, N, seed, mode)
    prop = thr @ (N / N.sum())
    coef = thr * (1.0 / prio / prop * sf)[:, None]
    present = np.ones((J + 1, J), bool)
    present[np.arange(1, J + 1), np.arange(J)] = False
    x = P._eisenberg_gale(N, coef, sf.astype(float), present)
    assert x.min() >= 0 and np.all(x.sum(axis=2) <= 1 + 1e-5)
    assert np.all((sf[None, :, None] * x).sum(axis=1) <= N[None] * (1 + 1e-4))
    worst = 0.0
    for s in (0, J // 2 + 1):
        idx = np.flatnonzero(present[s])
        _, u = gl.eisenberg_gale(coef[idx], sf[idx].astype(float), N)
        ug = (coef[idx] * x[s, idx]).sum(axis=1)
        worst = max(worst, float(np.abs(ug / u - 1).max()))
        assert np.all(x[s, ~present[s]] == 0)
    print("J", J, mode, "worst relative utility deviation", worst)
    assert worst < 5e-4


@pytest.mark.parametrize("J,N,seed,mode", [(8, [3, 2, 2], 11, "random"), (12, [4, 4, 4], 13, "proportional")])
def test_policy_call_matches_oracle_backend(engine, J, N, seed, mode):
    P._shared_engine = engine
    thr, sf, prio, N = _instance(J, N, seed, mode)
    tdict = {j: dict(zip(WT, thr[j].tolist())) for j in range(J)}
    sfd = {j: int(sf[j]) for j in range(J)}
    pd = {j: float(prio[j]) for j in range(J)}
    spec = dict(zip(WT, [int(v) for v in N]))
    pol = P.MaxMinFairnessStrategyProofPolicyWithPerf(solver="ECOS")
    alloc, disc = pol.get_allocation(tdict, sfd, pd, spec)
    thr_only = pol.get_allocation(tdict, sfd, pd, spec, recurse_deeper=False)
    with gb.cpu_backend() as PC:
        polc = PC.MaxMinFairnessStrategyProofPolicyWithPerf(solver="ECOS")
        alloc_c, disc_c = polc.get_allocation(tdict, sfd, pd, spec)
        thr_c = polc.get_allocation(tdict, sfd, pd, spec, recurse_deeper=False)
    assert np.allclose(disc, disc_c, rtol=5e-3), np.abs(disc / disc_c - 1).max()     # a product of J - 1 ratios
    assert np.allclose([thr_only[j] for j in range(J)], [thr_c[j] for j in range(J)], rtol=5e-4)
    a = np.array([[alloc[j][w] for w in WT] for j in range(J)])
    ac = np.array([[alloc_c[j][w] for w in WT] for j in range(J)])
    assert a.min() >= 0 and np.all(a.sum(axis=1) <= 1 + 1e-6)
    assert np.all((sf[:, None] * a).sum(axis=0) <= N * (1 + 1e-4))
    # discounted effective throughput per job (what the allocation is worth): unique, unlike x itself
    assert np.allclose((thr * a).sum(axis=1), (thr * ac).sum(axis=1), rtol=4e-3)
    assert P.get_policy("max_min_fairness_strategy_proof_perf").name == "MaxMinFairness_Perf"
