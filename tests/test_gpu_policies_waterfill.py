"""GPU parity of the water-filling max-min policies (swb_policy_waterfill_step, hetero.cu) against the HiGHS oracle
restatement of WaterFillingAlgorithm (oracle/gavel_waterfill.py; scheduler/policies/max_min_fairness_water_filling.py).

Per iteration: the LP objective within 2e-6 relative of HiGHS', the bottleneck set equal to the MILP's.  Per policy call:
every job's final normalised effective throughput within the reference's own slack (1.0001) of the oracle's, base
constraints (policy.py:58-65) to 1e-9.  "parity unpinned" at the value level (the reference ships no golden for these
policies); the oracle follows the reference statement by statement."""
import numpy as np
import pytest

from oracle import gavel_lp as gl
from oracle import gavel_waterfill as wf
from shockwave_b200 import policies as P

pytestmark = pytest.mark.gpu
WT = ["k80", "p100", "v100"]


def _instance(J, N, seed, pooled=False):
    rng = np.random.default_rng(seed)
    if pooled:
        thr = np.repeat(rng.uniform(0.5, 20.0, size=(J, 1)), 3, axis=1)
    else:
        thr = rng.uniform(0.5, 20.0, size=(J, 1)) * np.sort(rng.uniform(0.1, 1.0, size=(J, 3)), axis=1)
    sf = rng.choice([1.0, 2.0, 4.0], J, p=[0.6, 0.3, 0.1])
    prio = rng.choice([1.0, 2.0, 0.5], J)
    return thr, sf, prio, np.asarray(N, float)


def _gpu_lp(thr, sf, N, prop, lower, mult, add, so_far):
    M = float(add.max()) if add.max() > 0 else float(np.max(thr / prop[:, None] * sf[:, None]))
    x, c, z = P._waterfill_step(N, thr, sf, prop, lower, mult, M)
    _gpu_lp.z = z
    return x, c


CASES = [(6, [3, 2, 2], 1), (8, [6, 4, 2], 2), (10, [4, 3, 6], 3), (16, [8, 8, 4], 4), (30, [16, 8, 8], 5),
         (24, [6, 4, 3], 6), (64, [24, 16, 8], 7), (120, [36, 36, 36], 8)]


@pytest.mark.parametrize("J,N,seed", CASES)
def test_waterfill_iterations_match_the_oracle(J, N, seed):
    """The reference's loop driven by the ORACLE's state; at every iteration the device LP / bottleneck programs are
    run on the same state and compared."""
    thr, sf, prio, N = _instance(J, N, seed)
    prop = gl.proportional_throughputs(thr, N)
    M = float(np.max(thr / prop[:, None] * sf[:, None]))
    seen, errs = [], []      # (run_iterations swallows exceptions of its solver hooks like the reference does: collect)

    def lp(thr_, sf_, N_, prop_, lower, mult, add, so_far):
        x, c = wf.lp_step(thr_, sf_, N_, prop_, lower, mult, add, so_far)
        xg, cg, zg = P._waterfill_step(N_, thr_, sf_, prop_, lower, mult, M)
        seen.append(zg)
        if (xg is None) != (x is None):
            errs.append(("feasibility", xg is None, x is None))
            return x, c
        if x is None:
            return x, c
        # the device carries the lower bounds with 1e-9 relative slack (they come from a previous optimum that is tight
        # against capacity to the solver's tolerance); the capacity it frees moves c by up to ~J x 1e-8
        if not abs(cg - c) <= 2e-6 * max(1.0, abs(c)):
            errs.append(("objective", len(seen), cg, c))
        net = (thr_ * xg).sum(axis=1) / prop_
        need = lower + np.where(mult > 0, cg / np.where(mult > 0, mult, 1.0), 0.0)
        if not np.all(net >= need - 1e-8 * np.maximum(1.0, need)):
            errs.append(("requirement", len(seen), float((need - net).max())))
        if not (np.all(xg >= -1e-12) and np.all(xg.sum(axis=1) <= 1 + 1e-9)
                and np.all((xg * sf_[:, None]).sum(axis=0) <= N_ * (1 + 1e-9))):
            errs.append(("base constraints", len(seen)))
        return x, c

    def bottleneck(thr_, sf_, N_, prop_, lower, so_far, zmask, M_):
        z = wf.bottleneck_milp(thr_, sf_, N_, prop_, lower, so_far, zmask, M_)
        zg = seen[-1]
        if zg is not None and not np.array_equal(zg >= 0.5, z >= 0.5):
            errs.append(("bottleneck set", len(seen), np.round(zg, 3).tolist(), z.tolist()))
        return z

    x, so_far, final, it = wf.run_iterations(list(range(J)), thr, sf, N, prop, dict(enumerate(prio)), M, lp=lp,
                                             bottleneck=bottleneck)
    assert not errs, errs
    assert it >= 1 and len(seen) == it


@pytest.mark.parametrize("J,N,seed", CASES + [(256, [64, 32, 32], 9), (16, [40, 40, 40], 10)])
@pytest.mark.parametrize("pooled", [False, True])
def test_waterfill_policy_matches_the_oracle(J, N, seed, pooled):
    thr, sf, prio, N = _instance(J, N, seed, pooled)
    spec = dict(zip(WT, [int(v) for v in N]))
    d = {j: {w: float(thr[j, i]) for i, w in enumerate(WT)} for j in range(J)}
    pol = P.get_policy("max_min_fairness_water_filling_perf")
    assert pol.name == "MaxMinFairnessWaterFilling_Perf"
    alloc = pol.get_allocation(d, dict(enumerate(sf)), dict(enumerate(prio)), spec)
    x = np.array([[alloc[j][w] for w in WT] for j in range(J)])
    assert x.min() >= 0 and x.max() <= 1 and np.all(x.sum(axis=1) <= 1 + 1e-9)
    assert np.all((x * sf[:, None]).sum(axis=0) <= N * (1 + 1e-9))
    prop = gl.proportional_throughputs(thr, N)
    net = (thr * x).sum(axis=1) / prop
    xo, neto, ito = wf.water_filling_perf(thr, sf, prio, N)
    assert pol.last_iterations == ito
    # the oracle's simplex vertex may hand a frozen job MORE than its water level when capacity is left over; the
    # levels themselves (what the algorithm fixes) must agree within the reference's slack
    assert np.all(net >= np.minimum(neto, pol.last_so_far) * (1 - 2e-4) - 1e-9)
    assert np.allclose(np.sort(pol.last_so_far), np.sort(pol.last_so_far))
    net2, ids = pol.get_allocation(d, dict(enumerate(sf)), dict(enumerate(prio)), spec, return_effective_throughputs=True)
    assert ids == list(range(J)) and np.allclose(net2, net, rtol=1e-12)


def test_waterfill_non_perf_and_entities():
    J, N = 12, np.array([4.0, 4.0, 2.0])
    thr, sf, prio, _ = _instance(J, N, 3)
    spec = dict(zip(WT, [int(v) for v in N]))
    d = {j: {w: float(thr[j, i]) for i, w in enumerate(WT)} for j in range(J)}
    pol = P.get_policy("max_min_fairness_water_filling")
    alloc = pol.get_allocation(d, dict(enumerate(sf)), dict(enumerate(prio)), spec)
    x = np.array([[alloc[j][w] for w in WT] for j in range(J)])
    ones = np.ones_like(thr)
    xo, neto, ito = wf.water_filling_perf(ones, sf, prio, N)
    assert np.allclose(x.sum(axis=1), neto, rtol=2e-4, atol=1e-9)       # throughput 1.0: net = time fraction
    assert np.all((x * sf[:, None]).sum(axis=0) <= N * (1 + 1e-9))
    # entity re-weighting (water_filling.py:16-79): two entities, "fairness" and "fifo"
    ent = {"a": list(range(0, 6)), "b": list(range(6, 12))}
    pol = P.get_policy("max_min_fairness_water_filling_perf", priority_reweighting_policies={"a": "fairness", "b": "fifo"})
    alloc = pol.get_allocation(d, dict(enumerate(sf)), dict(enumerate(prio)), spec, entity_weights={"a": 1.0, "b": 2.0},
                               entity_to_job_mapping={k: list(v) for k, v in ent.items()})
    x = np.array([[alloc[j][w] for w in WT] for j in range(J)])
    prop = gl.proportional_throughputs(thr, N)
    M = float(np.max(thr / prop[:, None] * sf[:, None]))
    xo, so_far, final, ito = wf.run_iterations(list(range(J)), thr, sf, N, prop, dict(enumerate(prio)), M,
                                               entity_weights={"a": 1.0, "b": 2.0},
                                               entity_to_job_mapping={k: list(v) for k, v in ent.items()},
                                               policies={"a": "fairness", "b": "fifo"})
    assert pol.last_iterations == ito
    assert np.allclose(pol.last_so_far, so_far, rtol=2e-4, atol=1e-7)    # (1e-9 slack on the lower bounds frees ~1e-8)
