"""CPU tests of the water-filling oracle (oracle/gavel_waterfill.py): consistency with the max-min LP oracle and the
defining properties of a water-filling (lexicographic max-min) allocation.  The reference ships no golden for these
policies ("parity unpinned" at the value level)."""
import numpy as np

from oracle import gavel_lp as gl
from oracle import gavel_waterfill as wf


def _inst(J, seed):
    rng = np.random.default_rng(seed)
    thr = rng.uniform(0.5, 10.0, (J, 1)) * rng.uniform(0.1, 1.0, (J, 3))
    return thr, rng.choice([1.0, 2.0, 4.0], J, p=[.6, .3, .1]), rng.choice([1.0, 2.0, 0.5], J)


def test_first_iteration_is_the_max_min_lp():
    for J, N, seed in ((6, [3, 2, 2], 1), (16, [8, 8, 4], 4), (24, [6, 4, 3], 6)):
        thr, sf, prio = _inst(J, seed)
        log = []
        wf.water_filling_perf(thr, sf, prio, N, log=log)
        z, _ = gl.max_min_fairness_perf(thr, sf, prio, N)
        assert abs(log[0]["c"] - z) <= 1e-9 * z


def test_water_levels_are_lexicographically_max_min():
    for J, N, seed in ((8, [6, 4, 2], 2), (10, [4, 3, 6], 3), (30, [16, 8, 8], 5)):
        thr, sf, prio = _inst(J, seed)
        log = []
        x, net, it = wf.water_filling_perf(thr, sf, prio, N, log=log)
        N = np.asarray(N, float)
        assert np.all(x.sum(axis=1) <= 1 + 1e-9) and np.all((x * sf[:, None]).sum(axis=0) <= N + 1e-9)
        assert it == len(log) and it > 1
        nf = [l["nfinal"] for l in log]
        assert all(b > a for a, b in zip(nf, nf[1:])) and nf[-1] == J      # every iteration saturates somebody
        # weighted levels never decrease from one iteration to the next and no job ends below the plain max-min level
        z, _ = gl.max_min_fairness_perf(thr, sf, prio, N)
        w = (1.0 / prio) * sf
        assert np.all(net * w >= z * (1 - 1e-6))
        assert all(l["c"] >= -1e-12 for l in log)


def test_entity_reweighting_follows_the_reference_rules():
    final = {2: 0.5}
    pw = wf.compute_priority_weights({"a": 3.0, "b": 2.0}, {0: 1.0, 1: 3.0, 2: 5.0, 3: 1.0, 4: 1.0},
                                     {"a": [0, 1, 2], "b": [4, 3]}, final, [0, 1, 2, 3, 4], {"a": "fairness", "b": "fifo"})
    assert pw == {0: 3.0 * 0.25, 1: 3.0 * 0.75, 2: 0.0, 3: 2.0, 4: 0.0}


def test_policy_host_loop_with_the_oracle_backend():
    """shockwave_b200.policies' water-filling classes (host loop, pooling, entity re-weighting, return values) driven by
    the HiGHS backend instead of the CUDA library: same water levels and iteration counts as the restatement of the
    reference's own loop."""
    from oracle import gavel_backend as gb
    WT = ["k80", "p100", "v100"]
    for J, N, seed in ((8, [6, 4, 2], 2), (10, [4, 3, 6], 3), (24, [6, 4, 3], 6)):
        thr, sf, prio = _inst(J, seed)
        N = np.asarray(N, float)
        d = {j: {w: float(thr[j, i]) for i, w in enumerate(WT)} for j in range(J)}
        spec = dict(zip(WT, [int(v) for v in N]))
        with gb.cpu_backend() as P:
            pol = P.get_policy("max_min_fairness_water_filling_perf")
            alloc = pol.get_allocation(d, dict(enumerate(sf)), dict(enumerate(prio)), spec)
            net2, ids = pol.get_allocation(d, dict(enumerate(sf)), dict(enumerate(prio)), spec,
                                           return_effective_throughputs=True)
        x = np.array([[alloc[j][w] for w in WT] for j in range(J)])
        xo, neto, ito = wf.water_filling_perf(thr, sf, prio, N)
        assert pol.last_iterations == ito
        prop = gl.proportional_throughputs(thr, N)
        assert np.allclose((thr * x).sum(axis=1) / prop, neto, rtol=1e-6, atol=1e-9)
        assert ids == list(range(J)) and np.allclose(net2, neto, rtol=1e-6, atol=1e-9)
    # pooled: every live type gives the job the same throughput -> one pooled type, split by capacity
    thr, sf, prio = _inst(12, 9)
    thr = np.repeat(thr[:, :1], 3, axis=1)
    N = np.array([4.0, 2.0, 2.0])
    d = {j: {w: float(thr[j, i]) for i, w in enumerate(WT)} for j in range(12)}
    with gb.cpu_backend() as P:
        pol = P.get_policy("max_min_fairness_water_filling")          # non-Perf: throughputs replaced by 1.0
        alloc = pol.get_allocation(d, dict(enumerate(sf)), dict(enumerate(prio)), dict(zip(WT, [4, 2, 2])))
    x = np.array([[alloc[j][w] for w in WT] for j in range(12)])
    xo, neto, ito = wf.water_filling_perf(np.ones_like(thr), sf, prio, N)
    assert np.allclose(x.sum(axis=1), neto, rtol=1e-6, atol=1e-9)
    assert np.all((x * sf[:, None]).sum(axis=0) <= N * (1 + 1e-9))
