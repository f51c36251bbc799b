"""swb_sim_* on the B200: the device round loop against (1) the records of the unmodified reference
loop (tests/golden/sim_static_pins.json: completion times, makespan, rounds — bit for bit), (2) the pinned restatement
oracle/sim_loop.py round by round on random traces, one policy per scenario, and (3) the reference's own loop driving the
product's ShockwaveScheduler on the static 120-job trace vs ShockwaveEnsemble (same kernels behind round_schedule(),
device loop instead of the reference's Python bookkeeping)."""
import json
import os
import tempfile
import time

import numpy as np
import pytest

from oracle import ref_harness as rh
from oracle import sim_loop
from tests import sim_fixtures as sf_

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PINS = os.path.join(ROOT, "tests", "golden", "sim_static_pins.json")


def _mask(schedule, J):
    m = np.zeros((len(schedule), J), np.uint8)
    for r, ids in enumerate(schedule):
        m[r, ids] = 1
    return m


@pytest.mark.parametrize("name", ["fifo_32", "max_min_fairness_32", "max_min_fairness_12"])
def test_replay_of_the_reference_records(name):
    from shockwave_b200.simulate import DeviceSim
    rec = json.load(open(PINS))[name]
    tr = sim_loop.trace_arrays(rec)
    J = len(rec["arrival"])
    S = 5
    sim = DeviceSim(tr, S, rec["ngpus"], rec["time_per_iteration"])
    sched = _mask(rec["per_round_schedule"], J)
    pad = np.zeros((7, J), np.uint8)                       # rounds after the end are ignored
    scn = sim.replay(np.concatenate([sched, pad]))
    res = sim.results()
    want = np.array([rec["jct"][str(j)] for j in range(J)])
    for s in range(S):
        assert scn["err"][s] == 0 and scn["done"][s] == 1
        assert scn["rounds"][s] == rec["rounds"] and scn["now"][s] == rec["makespan"]
        assert np.array_equal(res["jct"][s], want)
    # step by step == one launch
    sim2 = DeviceSim(tr, 2, rec["ngpus"], rec["time_per_iteration"])
    z = sim2.begin()
    for r in range(len(sched)):
        z = sim2.step(np.stack([sched[r], sched[r]]))
    assert z["done"].all() and np.array_equal(sim2.results()["jct"][1], want)
    last = rec["timeline"]
    for j in range(J):
        ns, end = sf_.timeline_summary([tuple(e) for e in last[str(j)]], 120.0)
        assert sim2.tl_ns[0, j] == ns and sim2.tl_end[0, j] == end
    sim.close(); sim2.close()


@pytest.mark.parametrize("J,ngpus,gap", [(150, 16, False), (30, 4, True), (700, 64, False)])
def test_step_by_step_against_the_restatement(J, ngpus, gap):
    """One random policy per scenario; every round's outputs (status, epoch progress, timeline sums) and the final
    completion times equal oracle/sim_loop.py."""
    from shockwave_b200.simulate import DeviceSim
    tr = sf_.random_trace(J, 40 + J, gap)
    S = 4
    sels = [sf_.random_policy(tr, ngpus, 7 + s) for s in range(S)]
    oras = [sim_loop.run(tr, sels[s], tpi=120.0) for s in range(S)]
    sim = DeviceSim(tr, S, ngpus, 120.0)
    scn = sim.begin().copy()
    status = sim.status.copy()
    c = 0
    while not scn["done"].all():
        chosen = np.zeros((S, J), np.uint8)
        for s in range(S):
            if scn["done"][s]:
                continue
            active = np.flatnonzero(status[s] == 1).tolist()
            ids = [j for j in sels[s](c, scn["now"][s], active) if status[s, j] == 1]
            assert sorted(ids) == oras[s]["per_round_schedule"][c]
            chosen[s, ids] = 1
        scn = sim.step(chosen).copy()
        status = sim.status.copy()
        for s in range(S):
            for j in np.flatnonzero(chosen[s]):
                upto = [e for e in oras[s]["timeline"][j] if e[0] <= c + 1]
                ns, end = sf_.timeline_summary(upto, 120.0)
                assert sim.tl_ns[s, j] == ns and sim.tl_end[s, j] == end
        c += 1
        assert c < 100000
    res = sim.results()
    for s in range(S):
        assert scn["err"][s] == 0
        assert scn["rounds"][s] == oras[s]["rounds"] and scn["now"][s] == oras[s]["makespan"]
        assert np.array_equal(res["jct"][s], np.array(oras[s]["jct"]), equal_nan=True)
        assert np.array_equal(res["steps_run"][s], np.array(oras[s]["steps_run"]))
    sim.close()


def test_argument_checks():
    from shockwave_b200.simulate import DeviceSim
    tr = sf_.random_trace(8, 1)
    dynamic = DeviceSim(dict(tr, adaptation_mode=np.array([0, 0, 1, 0, 0, 0, 0, 0])), 1, 4)
    with pytest.raises(RuntimeError, match="swb_sim_set_dynamic"):
        dynamic.begin()                               # accordion job, no tables: refused, never run as a static job
    dynamic.close()
    with pytest.raises(RuntimeError, match="adaptation_mode"):
        DeviceSim(dict(tr, adaptation_mode=np.full(8, 7)), 1, 4)
    with pytest.raises(RuntimeError, match="non-decreasing"):
        DeviceSim(dict(tr, arrival=tr["arrival"][::-1].copy()), 1, 4)
    sim = DeviceSim(tr, 1, 1)
    sim.begin()
    z = sim.step(np.ones((1, 8), np.uint8))           # everything at once on one GPU: flagged, not silently accepted
    assert z["err"][0] & 1
    sim.close()


@pytest.mark.parametrize("name", ["fifo_32", "max_min_fairness_32", "max_min_fairness_12"])
def test_replay_of_the_reference_records_dynamic_trace(name):
    """The canonical trace AS SHIPPED (59 accordion + 57 gns + 4 static jobs): batch-size rescaling from tables, the
    micro-task failure branch; completion times, makespan, rounds and the timeline sums equal the reference's records."""
    from shockwave_b200.simulate import DeviceSim
    d = json.load(open(os.path.join(ROOT, "tests", "golden", "sim_dynamic_pins.json")))
    rec, dyn = d[name], d["fifo_32"]["dyn"]
    tr = dict(sim_loop.trace_arrays(rec), adaptation_mode=np.asarray(dyn["mode"]))
    J = len(rec["arrival"])
    sim = DeviceSim(tr, 3, rec["ngpus"], rec["time_per_iteration"])
    sim.set_dynamic(dyn)
    sched = _mask(rec["per_round_schedule"], J)
    scn = sim.replay(sched)
    res = sim.results()
    want = np.array([rec["jct"][str(j)] for j in range(J)])
    for s in range(3):
        assert scn["err"][s] == 0 and scn["done"][s] == 1
        assert scn["rounds"][s] == rec["rounds"] and scn["now"][s] == rec["makespan"]
        assert np.array_equal(res["jct"][s], want)
    sim2 = DeviceSim(tr, 1, rec["ngpus"], rec["time_per_iteration"])
    sim2.set_dynamic(dyn)
    sim2.begin()
    for r in range(len(sched)):
        z = sim2.step(sched[r][None, :])
    assert z["done"].all() and np.array_equal(sim2.results()["jct"][0], want)
    for j in range(J):
        ns, end = sf_.timeline_summary([tuple(e) for e in rec["timeline"][str(j)]], 120.0)
        assert sim2.tl_ns[0, j] == ns and sim2.tl_end[0, j] == end
    sim.close(); sim2.close()


@pytest.mark.reference
@pytest.mark.skipif(not rh.reference_available(), reason="needs the (staged) reference simulator")
@pytest.mark.parametrize("static", [True, False])
def test_shockwave_ensemble_against_the_reference_loop(static):
    """The 120-job trace on 32 GPUs, with every job static and as shipped (accordion / gns jobs).  (a) the UNMODIFIED reference loop drives the product's ShockwaveScheduler;
    (b) ShockwaveEnsemble runs the same scenario + two what-ifs on the device loop.  Same kernels, same inputs: the
    scenario equal to (a) must land on the same end-to-end metrics (tolerance 1 % for solver-level tie-breaks; the
    exact-equality flags are recorded in gpurun_out/sim_ensemble.json)."""
    from shockwave_b200 import ShockwaveScheduler
    from shockwave_b200.simulate import ShockwaveEnsemble
    from tests.golden import make_sim_pins as pins
    scratch = tempfile.mkdtemp(prefix="swens_")
    dst = pins.stage_static_trace(scratch, static=static)

    def extract(sched, jobs, arrival_times):
        rec = (pins.extract if static else pins.extract_dynamic)(sched, jobs, arrival_times)
        rec["profiles"] = [dict(p) for p in sched._profiles[:len(jobs)]]
        return rec
    t0 = time.perf_counter()
    ref = rh.simulate("shockwave", shockwave_scheduler_cls=ShockwaveScheduler, trace=pins.REL, scratch=scratch,
                      cluster="32:0:0", extract=extract)
    t_ref = time.perf_counter() - t0
    rec = ref["extra"]
    cfg = json.load(open(os.path.join(dst, "configurations/tacc_32gpus.json")))
    tr = {k: np.asarray(rec[k]) for k in ("arrival", "total_steps", "scale_factor", "throughput", "duration", "batch_size",
                                          "dataset_len")}
    scen = [{}, {"k": cfg["k"] * 10}, {"lambda": cfg["lambda"] * 2}, {"future_rounds": cfg["future_rounds"] + 10}]
    ens = ShockwaveEnsemble(tr, rec["profiles"], cfg, scen, ngpus=32, time_per_iteration=120,
                            dynamic=None if static else rec["dyn"])
    t0 = time.perf_counter()
    out = ens.run()
    t_ens = time.perf_counter() - t0
    J = len(rec["arrival"])
    want_jct = np.array([rec["jct"][str(j)] for j in range(J)])
    want_sched = [sorted(int(k) for k in rnd.keys()) for rnd in ref["per_round_schedule"]]
    row = dict(reference_loop_s=round(t_ref, 2), ensemble_s=round(t_ens, 2), scenarios=len(scen),
               makespan_ref=ref["makespan"], makespan=out["makespan"].tolist(), avg_jct_ref=ref["avg_jct"],
               avg_jct=out["avg_jct"].tolist(), rounds_ref=rec["rounds"], rounds=out["rounds"].tolist(),
               resolves=out["resolves"].tolist(),
               schedule_identical=bool(out["per_round_schedule"][0] == want_sched),
               jct_identical=bool(np.array_equal(out["jct"][0], want_jct)))
    row["trace"] = "canonical 120-job trace, every job static" if static else "canonical 120-job trace as shipped (dynamic)"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "sim_ensemble.json")
    rows = json.load(open(path)) if os.path.exists(path) else []
    rows = [r for r in (rows if isinstance(rows, list) else []) if r.get("trace") != row["trace"]] + [row]
    json.dump(rows, open(path, "w"), indent=1)
    if not static:
        # P5 on the device loop: the reference's golden pickle of this very trace (reproduce/pickles/tacc_32gpus), inside
        # the 3 % spread of its own three Shockwave pickles — no reference loop in between
        import glob
        import pickle
        gold = glob.glob(os.path.join(rh.GOLDEN_DIR, "shockwave_*"))
        if gold:
            g = pickle.load(open(gold[0], "rb"))
            d = ens.result_dicts()[0]
            row.update(golden_makespan=g["makespan"], golden_avg_jct=g["avg_jct"],
                       golden_worst_ftf=float(np.max(g["finish_time_fairness_list"])),
                       worst_ftf=float(np.max(d["finish_time_fairness_list"])))
            json.dump([r for r in rows if r.get("trace") != row["trace"]] + [row], open(path, "w"), indent=1)
            assert abs(out["makespan"][0] - g["makespan"]) <= 0.03 * g["makespan"]
            assert abs(out["avg_jct"][0] - g["avg_jct"]) <= 0.03 * g["avg_jct"]
            assert abs(len(out["per_round_schedule"][0]) - len(g["per_round_schedule"])) <= 6
    assert np.isfinite(out["jct"]).all()
    assert abs(out["makespan"][0] - ref["makespan"]) <= 0.01 * ref["makespan"]
    assert abs(out["avg_jct"][0] - ref["avg_jct"]) <= 0.01 * ref["avg_jct"]
    assert abs(int(out["rounds"][0]) - rec["rounds"]) <= 3


@pytest.mark.reference
@pytest.mark.skipif(not rh.reference_available(), reason="needs the (staged) reference simulator")
@pytest.mark.parametrize("policy", ["max_min_fairness", "finish_time_fairness"])
def test_policy_ensemble_against_the_reference_loop(policy):
    """Gavel policies: (a) the unmodified reference loop with the product's policies and round step (GavelRoundMixin) on
    the canonical trace as shipped; (b) PolicyEnsemble: same device calls, the loop and the mechanism's bookkeeping off
    the reference's dicts.  Same inputs into the same kernels: the schedules must coincide (1 % tolerance on the
    end-to-end metrics as the assertion, exactness recorded)."""
    from shockwave_b200 import policies as P
    from shockwave_b200.placement import GavelRoundMixin
    from shockwave_b200.simulate import PolicyEnsemble
    from tests.golden import make_sim_pins as pins
    scratch = tempfile.mkdtemp(prefix="swpol_")
    rh.prepare_tree(scratch)
    t0 = time.perf_counter()
    ref = rh.simulate(policy, policy_obj=P.get_policy(policy, solver="ECOS", seed=0), scratch=scratch, cluster="32:0:0",
                      extract=pins.extract_dynamic, scheduler_mixin=GavelRoundMixin)
    t_ref = time.perf_counter() - t0
    rec = ref["extra"]
    tr = {k: np.asarray(rec[k]) for k in ("arrival", "total_steps", "scale_factor", "throughput", "duration", "batch_size",
                                          "dataset_len")}
    ens = PolicyEnsemble(tr, [P.get_policy(policy, solver="ECOS", seed=0) for _ in range(2)], 32, dynamic=rec["dyn"])
    t0 = time.perf_counter()
    out = ens.run()
    t_ens = time.perf_counter() - t0
    want = [{int(k): tuple(v) for k, v in rnd.items()} for rnd in ref["per_round_schedule"]]
    J = len(rec["arrival"])
    want_jct = np.array([rec["jct"][str(j)] for j in range(J)])
    row = dict(trace=f"canonical trace as shipped, {policy} (PolicyEnsemble)", reference_loop_s=round(t_ref, 2),
               ensemble_s=round(t_ens, 2), scenarios=2, makespan_ref=ref["makespan"], makespan=out["makespan"].tolist(),
               rounds_ref=rec["rounds"], rounds=out["rounds"].tolist(), allocations=out["allocations"].tolist(),
               schedule_identical=bool(out["per_round_schedule"][0] == want),
               jct_identical=bool(np.array_equal(out["jct"][0], want_jct)))
    path = os.path.join(ROOT, "gpurun_out", "sim_ensemble.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    rows = json.load(open(path)) if os.path.exists(path) else []
    json.dump([r for r in rows if r.get("trace") != row["trace"]] + [row], open(path, "w"), indent=1)
    assert np.isfinite(out["jct"]).all()
    assert abs(out["makespan"][0] - ref["makespan"]) <= 0.01 * ref["makespan"]
    assert abs(np.mean(out["jct"][0]) - ref["avg_jct"]) <= 0.01 * ref["avg_jct"]
