"""oracle/sim_loop.py (restatement of the reference simulator's round loop for static jobs) replays the schedules the
UNMODIFIED reference recorded (tests/golden/make_sim_pins.py) and must reproduce the reference's own bookkeeping:
completion time of every job, makespan, number of rounds, measured-throughput timeline — bit for bit."""
import json
import os

import numpy as np
import pytest

from oracle import sim_loop

PINS = os.path.join(os.path.dirname(__file__), "golden", "sim_static_pins.json")


def load():
    with open(PINS) as f:
        return json.load(f)


@pytest.mark.parametrize("name", ["fifo_32", "max_min_fairness_32", "max_min_fairness_12"])
def test_replay_of_recorded_schedules_reproduces_the_reference(name):
    rec = load()[name]
    sched = rec["per_round_schedule"]
    out = sim_loop.run(rec, lambda c, now, active: sched[c], tpi=rec["time_per_iteration"])
    assert out["rounds"] == rec["rounds"] == len(sched)
    assert out["makespan"] == rec["makespan"]
    assert out["per_round_schedule"] == sched
    for j in range(len(rec["arrival"])):
        assert out["jct"][j] == rec["jct"][str(j)], j
        want = [(r, t, b) for r, t, b in rec["timeline"][str(j)]]
        assert out["timeline"][j] == want, j
    assert np.mean(out["jct"]) == pytest.approx(rec["avg_jct"], rel=1e-12)


# ---- the device round loop's source on the host (tests/native/sim_host.cpp) against the pinned restatement ----
from tests import sim_fixtures as sf_  # noqa: E402


@pytest.fixture(scope="module")
def host_lib():
    lib = sf_.host_sim_lib()
    if lib is None:
        pytest.skip("g++ not available")
    return lib


def run_host(lib, tr, select, ngpus, tpi, grd):
    sim = sf_.HostSim(lib, tr, ngpus, tpi, grd)
    scn = sim.begin()
    arrival = np.asarray(tr["arrival"])
    status = (arrival <= scn.now).astype(np.uint8)          # begin() admitted these (checked below through step)
    per_round, epochs, tls = [], [], []
    c = 0
    while not scn.done:
        active = [int(j) for j in np.flatnonzero(status == 1)]
        chosen = [j for j in select(c, scn.now, active) if status[j] == 1]
        per_round.append(sorted(chosen))
        scn = sim.step(chosen)
        status = sim.status.copy()
        epochs.append(sim.epoch.copy())
        tls.append((sim.tl_ns.copy(), sim.tl_end.copy(), sim.thr_meas.copy()))
        c += 1
    jct, steps_run, run_time = sim.results()
    return dict(makespan=scn.now, rounds=scn.rounds, err=scn.err, jct=jct, steps_run=steps_run,
                per_round_schedule=per_round, epochs=epochs, tls=tls, remaining=scn.remaining)


def compare(host, ora, tr, grd):
    assert host["err"] == 0
    assert host["rounds"] == ora["rounds"] and host["makespan"] == ora["makespan"]
    assert host["per_round_schedule"] == ora["per_round_schedule"]
    J = len(tr["arrival"])
    for j in range(J):
        a, b = host["jct"][j], ora["jct"][j]
        assert (np.isnan(a) and np.isnan(b)) or a == b, j
        assert host["steps_run"][j] == ora["steps_run"][j]
    # timeline summary (running sum on the device) == JobMetaData's walk over the full timeline, after every round
    for c, (ns, end, tm) in enumerate(host["tls"]):
        for j in host["per_round_schedule"][c]:
            upto = [e for e in ora["timeline"][j] if e[0] <= c + 1]
            want_ns, want_end = sf_.timeline_summary(upto, grd)
            assert ns[j] == want_ns and end[j] == want_end, (c, j)
            assert tm[j] == upto[-1][1]


@pytest.mark.parametrize("name", ["fifo_32", "max_min_fairness_12"])
def test_device_loop_source_replays_the_reference_records(host_lib, name):
    rec = load()[name]
    sched = rec["per_round_schedule"]
    tr = sim_loop.trace_arrays(rec)
    host = run_host(host_lib, tr, lambda c, now, active: sched[c], rec["ngpus"], rec["time_per_iteration"], 120.0)
    assert host["err"] == 0 and host["rounds"] == rec["rounds"] and host["makespan"] == rec["makespan"]
    for j in range(len(rec["arrival"])):
        assert host["jct"][j] == rec["jct"][str(j)]
    ora = sim_loop.run(rec, lambda c, now, active: sched[c], tpi=rec["time_per_iteration"])
    compare(host, ora, tr, 120.0)


@pytest.mark.parametrize("J,ngpus,seed,gap", [(40, 8, 1, False), (150, 16, 2, False), (30, 4, 3, True), (300, 64, 4, False),
                                              (12, 2, 5, True)])
def test_device_loop_source_on_random_traces(host_lib, J, ngpus, seed, gap):
    """Random static traces (idle gaps, over-deadline retirements, gangs) under a random policy: host build of the
    device loop == restatement, every round; epoch progress as _update_shockwave_scheduler computes it."""
    tr = sf_.random_trace(J, seed, gap)
    sel = sf_.random_policy(tr, ngpus, seed)
    seen = {}

    def on_round(c, now, info):
        seen[c] = info["epoch"]
    ora = sim_loop.run(tr, sel, tpi=120.0, on_round=on_round)
    host = run_host(host_lib, tr, sel, ngpus, 120.0, 120.0)
    compare(host, ora, tr, 120.0)
    for c, ep in seen.items():                      # on_round(c) reports the round c - 1 that just ended
        for j, e in ep.items():
            got = host["epochs"][c - 1][j]
            assert got == (-1 if e is None else e), (c, j)
    if gap:
        assert np.isnan(host["jct"]).any() or host["remaining"] == 0      # the reference stops when the cluster drains


# ---- dynamic adaptation (accordion / gns): the canonical trace as shipped ----
DPINS = os.path.join(os.path.dirname(__file__), "golden", "sim_dynamic_pins.json")


def load_dynamic():
    with open(DPINS) as f:
        d = json.load(f)
    dyn = d["fifo_32"]["dyn"]
    return d, dyn


@pytest.mark.parametrize("name", ["fifo_32", "max_min_fairness_32", "max_min_fairness_12"])
def test_dynamic_replay_reproduces_the_reference(name):
    d, dyn = load_dynamic()
    rec = d[name]
    sched = rec["per_round_schedule"]
    out = sim_loop.run(rec, lambda c, now, active: sched[c], tpi=rec["time_per_iteration"], dyn=dyn)
    assert out["errs"] == []
    assert out["rounds"] == rec["rounds"] == len(sched)
    assert out["makespan"] == rec["makespan"]
    for j in range(len(rec["arrival"])):
        assert out["jct"][j] == rec["jct"][str(j)], j
        assert out["timeline"][j] == [(r, t, b) for r, t, b in rec["timeline"][str(j)]], j


def run_host_dyn(lib, rec, dyn, select):
    tr = sim_loop.trace_arrays(rec)
    sim = sf_.HostSim(lib, tr, rec["ngpus"], rec["time_per_iteration"], 120.0)
    sim.set_dynamic(dyn)
    scn = sim.begin()
    status = (tr["arrival"] <= scn.now).astype(np.uint8)
    per_round, eps = [], []
    c = 0
    while not scn.done:
        active = [int(j) for j in np.flatnonzero(status == 1)]
        chosen = [j for j in select(c, scn.now, active) if status[j] == 1]
        per_round.append(sorted(chosen))
        scn = sim.step(chosen)
        status = sim.status.copy()
        eps.append(sim.epoch.copy())
        c += 1
    jct, steps_run, run_time = sim.results()
    return dict(makespan=scn.now, rounds=scn.rounds, err=scn.err, jct=jct, steps_run=steps_run, per_round_schedule=per_round,
                tl_ns=sim.tl_ns.copy(), tl_end=sim.tl_end.copy(), epochs=eps)


@pytest.mark.parametrize("name", ["fifo_32", "max_min_fairness_32", "max_min_fairness_12"])
def test_device_loop_source_on_the_canonical_dynamic_trace(host_lib, name):
    """Host build of sim_core.cuh with the dynamic tables == the reference's records on the trace as shipped (59
    accordion + 57 gns jobs; a ResNet-50 job ends through the micro-task failure branch)."""
    d, dyn = load_dynamic()
    rec = d[name]
    sched = rec["per_round_schedule"]
    host = run_host_dyn(host_lib, rec, dyn, lambda c, now, active: sched[c])
    assert host["err"] == 0 and host["rounds"] == rec["rounds"] and host["makespan"] == rec["makespan"]
    for j in range(len(rec["arrival"])):
        assert host["jct"][j] == rec["jct"][str(j)], j
        tl = [tuple(e) for e in rec["timeline"][str(j)]]
        ns, end = sf_.timeline_summary(tl, 120.0)
        assert host["tl_ns"][j] == ns and host["tl_end"][j] == end, j


def test_device_loop_source_dynamic_random_policy(host_lib):
    """Another schedule than the recorded ones (random gang selection on 16 GPUs): host build == restatement, including
    the epoch progress handed to the shockwave hook after every round."""
    d, dyn = load_dynamic()
    rec = dict(d["fifo_32"], ngpus=16)
    sel = sf_.random_policy(rec, 16, 3)
    seen = {}
    ora = sim_loop.run(rec, sel, tpi=120.0, dyn=dyn, on_round=lambda c, now, info: seen.__setitem__(c, info["epoch"]))
    host = run_host_dyn(host_lib, rec, dyn, sel)
    assert host["err"] == 0 and ora["errs"] == []
    assert host["rounds"] == ora["rounds"] and host["makespan"] == ora["makespan"]
    assert host["per_round_schedule"] == ora["per_round_schedule"]
    assert np.array_equal(host["jct"], np.array(ora["jct"]), equal_nan=True)
    assert np.array_equal(host["steps_run"], np.array(ora["steps_run"]))
    for c, ep in seen.items():
        for j, e in ep.items():
            assert host["epochs"][c - 1][j] == (-1 if e is None else e), (c, j)
