"""B200: the device round loop on several worker types (swb_sim_set_worker_types) replays the schedules the UNMODIFIED
reference recorded on mixed v100 / p100 / k80 clusters (tests/golden/sim_hetero_pins.json) — completion times, makespan
and rounds bit-identical — in one launch and step by step; random per-scenario schedules equal the restatement
oracle/sim_loop.py round by round; PolicyEnsemble on mixed clusters against the reference loop driving the same kernels.
B200 record: profiles/gputests_sim_hetero_r02.txt, profiles/sim_ensemble_hetero_r02.json."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
PINS = json.load(open(os.path.join(HERE, "golden", "sim_hetero_pins.json")))


def _inputs(rec):
    from oracle import sim_loop
    tr = sim_loop.trace_arrays(rec)
    return tr, np.asarray(rec["throughput_w"], np.float64), np.asarray(rec["ngpus_w"], np.int32)


@pytest.mark.parametrize("key", sorted(PINS))
def test_replay_of_the_reference_records_on_mixed_clusters(key):
    from shockwave_b200.simulate import DeviceSim
    rec = PINS[key]
    tr, thr_w, cap = _inputs(rec)
    J, R, S = len(tr["arrival"]), len(rec["per_round_schedule"]), 3
    sched = np.zeros((R, J), np.uint8)
    for r, rnd in enumerate(rec["per_round_schedule"]):
        for j, w in rnd:
            sched[r, j] = 1 + w
    sim = DeviceSim(tr, S, int(cap.sum()), 120.0, 120.0)
    sim.set_worker_types(thr_w, cap)
    scn = sim.replay(sched)                              # all rounds in one launch
    res = sim.results()
    for s in range(S):
        assert scn["err"][s] == 0 and scn["done"][s] == 1
        assert scn["rounds"][s] == rec["rounds"] and scn["now"][s] == rec["makespan"]
        for j in range(J):
            assert res["jct"][s, j] == rec["jct"][str(j)]
    sim.close()
    sim = DeviceSim(tr, 1, int(cap.sum()), 120.0, 120.0)  # one round per call
    sim.set_worker_types(thr_w, cap)
    scn = sim.begin()
    for r in range(R):
        scn = sim.step(sched[r][None, :])
        assert scn["err"][0] == 0
    assert scn["done"][0] == 1 and scn["now"][0] == rec["makespan"]
    assert np.array_equal(sim.results()["jct"][0], res["jct"][0], equal_nan=True)
    sim.close()


def test_random_schedules_per_scenario_equal_the_restatement():
    from oracle import sim_loop
    from shockwave_b200.simulate import DeviceSim
    from tests import sim_fixtures as sf_
    J, S, W = 150, 6, 3
    tr = sf_.random_trace(J, 11)
    rng = np.random.default_rng(5)
    thr_w = tr["throughput"][:, None] * rng.uniform(0.2, 1.0, (J, W))
    cap = np.array([6, 5, 9], np.int32)
    sf = np.asarray(tr["scale_factor"])

    def policy(seed):
        def select(c, now, active):
            r = np.random.default_rng(seed * 7919 + c)
            left = cap.copy()
            out = {}
            for i in r.permutation(len(active)):
                j = active[i]
                w = int(r.integers(W))
                if sf[j] <= left[w]:
                    out[j] = w
                    left[w] -= sf[j]
            return out
        return select
    want = [sim_loop.run(tr, policy(s), throughput_w=thr_w) for s in range(S)]
    sim = DeviceSim(tr, S, int(cap.sum()), 120.0, 120.0)
    sim.set_worker_types(thr_w, cap)
    scn = sim.begin()
    sel = [policy(s) for s in range(S)]
    c = 0
    while not scn["done"].all():
        ch = np.zeros((S, J), np.uint8)
        for s in range(S):
            if not scn["done"][s]:
                for j, w in sel[s](c, float(scn["now"][s]), np.flatnonzero(sim.status[s] == 1).tolist()).items():
                    ch[s, j] = 1 + w
        scn = sim.step(ch)
        assert (scn["err"] == 0).all()
        c += 1
    res = sim.results()
    for s in range(S):
        assert scn["now"][s] == want[s]["makespan"] and scn["rounds"][s] == want[s]["rounds"]
        assert np.array_equal(res["jct"][s], np.asarray(want[s]["jct"]), equal_nan=True)
        assert np.array_equal(res["steps_run"][s], np.asarray(want[s]["steps_run"]))
    sim.close()


def test_worker_types_argument_checks_and_error_flags():
    from shockwave_b200.simulate import DeviceSim
    from tests import sim_fixtures as sf_
    tr = sf_.random_trace(12, 3)
    tr["scale_factor"][:] = 1
    thr_w = np.stack([tr["throughput"], 0.5 * tr["throughput"]], axis=1)
    thr_w[0, 1] = 0.0
    sim = DeviceSim(tr, 2, 4, 120.0, 120.0)
    with pytest.raises(RuntimeError):
        sim.set_worker_types(-thr_w, [2, 2])
    with pytest.raises(RuntimeError):
        sim.set_worker_types(np.ones((12, 9)), [1] * 9)
    sim.set_worker_types(thr_w, [2, 2])
    sim.begin()
    ch = np.zeros((2, 12), np.uint8)
    ch[0, [0, 1, 2]] = 1                                  # three gangs on two workers of type 0
    ch[1, 0] = 2                                          # job 0 cannot run on type 1
    scn = sim.step(ch)
    assert scn["err"][0] & 1 and scn["err"][1] & 8
    sim.close()
    dyn_tr = dict(tr, adaptation_mode=np.ones(12, np.int32))
    sim = DeviceSim(dyn_tr, 1, 4, 120.0, 120.0)
    with pytest.raises(RuntimeError):                     # several worker types run static jobs only
        sim.set_worker_types(thr_w, [2, 2])
    sim.close()


@pytest.mark.reference
@pytest.mark.parametrize("policy,keep,cluster", [("max_min_fairness_perf", 40, "4:3:2"), ("finish_time_fairness_perf", 36, "2:4:4"),
                                                  ("max_min_fairness", 40, "4:3:2")])
def test_policy_ensemble_on_a_mixed_cluster_against_the_reference_loop(policy, keep, cluster):
    """(a) the unmodified reference loop on a v100 + p100 + k80 cluster with the product's policies (hetero.cu) and round
    step (gavel.cu, GavelRoundMixin); (b) PolicyEnsemble(worker_types=...): the same device calls, the loop on
    swb_sim_step and the per-type bookkeeping off the reference's dicts.  Same inputs into the same kernels: the schedules
    must coincide (1 % on the end-to-end metrics as the assertion, exactness recorded in gpurun_out/sim_ensemble.json).
    The host logic is held to exact equality on the CPU (tests/test_policy_ensemble_hetero_host.py)."""
    from oracle import ref_harness as rh
    from oracle import sim_loop
    if not rh.reference_available():
        pytest.skip("needs the (staged) reference simulator")
    from shockwave_b200 import policies as P
    from shockwave_b200.simulate import PolicyEnsemble
    from tests.golden import make_sim_hetero_pins as gen
    rec = gen.record(policy, keep, cluster, product=True)
    tr = sim_loop.trace_arrays(rec)
    wt = dict(names=rec["worker_types"], throughput=np.asarray(rec["throughput_w"]), ngpus=rec["ngpus_w"])
    ens = PolicyEnsemble(tr, [P.get_policy(policy, solver="ECOS", seed=0) for _ in range(2)], None, worker_types=wt, seed=0)
    out = ens.run()
    want = [{j: tuple(ws) for j, ws in rnd} for rnd in rec["per_round_workers"]]
    want_jct = np.array([rec["jct"][str(j)] for j in range(keep)])
    row = dict(trace=f"first {keep} jobs of the canonical trace (static), cluster {cluster}, {policy} (PolicyEnsemble, "
                     f"several worker types)", makespan_ref=rec["makespan"], makespan=out["makespan"].tolist(),
               rounds_ref=rec["rounds"], rounds=out["rounds"].tolist(), allocations=out["allocations"].tolist(),
               schedule_identical=bool(out["per_round_schedule"][0] == want),
               jct_identical=bool(np.array_equal(out["jct"][0], want_jct)))
    root = os.path.dirname(HERE)
    path = os.path.join(root, "gpurun_out", "sim_ensemble.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    rows = json.load(open(path)) if os.path.exists(path) else []
    json.dump([r for r in rows if r.get("trace") != row["trace"]] + [row], open(path, "w"), indent=1)
    assert np.isfinite(out["jct"]).all()
    assert abs(out["makespan"][0] - rec["makespan"]) <= 0.01 * rec["makespan"]
    assert abs(np.mean(out["jct"][0]) - rec["avg_jct"]) <= 0.01 * rec["avg_jct"]
