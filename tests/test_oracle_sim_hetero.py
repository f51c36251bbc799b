"""Simulator round loop on SEVERAL worker types (static jobs; the heterogeneity-aware Gavel policies on mixed clusters):
the restatement oracle/sim_loop.py and the HOST build of the device loop (sim_core.cuh through tests/native/sim_host.cpp)
replay the schedules the UNMODIFIED reference recorded on v100 + p100 + k80 clusters
(tests/golden/make_sim_hetero_pins.py -> sim_hetero_pins.json) and must reproduce every completion time, the makespan and
the round count to the last bit.  The same records are replayed on the B200 in tests/test_zz_gpu_sim_hetero.py."""
import json
import os

import numpy as np
import pytest

from oracle import ref_harness as rh
from oracle import sim_loop
from tests import sim_fixtures as sf_

HERE = os.path.dirname(os.path.abspath(__file__))
PINS = json.load(open(os.path.join(HERE, "golden", "sim_hetero_pins.json")))


def hetero_trace(rec):
    tr = sim_loop.trace_arrays(rec)
    thr_w = np.asarray(rec["throughput_w"], np.float64)
    return tr, thr_w, np.asarray(rec["ngpus_w"], np.int32)


@pytest.mark.parametrize("key", sorted(PINS))
def test_restatement_replays_the_reference_on_mixed_clusters(key):
    rec = PINS[key]
    tr, thr_w, cap = hetero_trace(rec)
    sched = [{j: w for j, w in rnd} for rnd in rec["per_round_schedule"]]
    out = sim_loop.run(tr, lambda c, now, active: sched[c], throughput_w=thr_w)
    assert out["rounds"] == rec["rounds"] and out["makespan"] == rec["makespan"]
    for j in range(len(tr["arrival"])):
        assert out["jct"][j] == rec["jct"][str(j)], (j, out["jct"][j], rec["jct"][str(j)])
    # the recorded schedules use more than one worker type and respect the per-type capacities
    used = {w for rnd in sched for w in rnd.values()}
    assert len(used) == len(cap)
    for rnd in sched:
        for w in range(len(cap)):
            assert sum(int(tr["scale_factor"][j]) for j, ww in rnd.items() if ww == w) <= cap[w]


@pytest.mark.parametrize("key", sorted(PINS))
def test_host_build_of_the_device_loop_replays_the_reference_on_mixed_clusters(key):
    lib = sf_.host_sim_lib()
    if lib is None:
        pytest.skip("g++ not available")
    rec = PINS[key]
    tr, thr_w, cap = hetero_trace(rec)
    m = sf_.HostSim(lib, tr, int(cap.sum()), 120.0, 120.0)
    m.set_worker_types(thr_w, cap)
    z = m.begin()
    for rnd in rec["per_round_schedule"]:
        assert not z.done
        z = m.step({j: w for j, w in rnd})
        assert z.err == 0
    assert z.done and z.rounds == rec["rounds"] and z.now == rec["makespan"]
    jct = m.results()[0]
    for j in range(m.J):
        assert jct[j] == rec["jct"][str(j)]


def test_host_build_flags_what_the_reference_raises_on():
    lib = sf_.host_sim_lib()
    if lib is None:
        pytest.skip("g++ not available")
    tr = sf_.random_trace(12, 3)
    tr["scale_factor"][:] = 1
    thr_w = np.stack([tr["throughput"], 0.5 * tr["throughput"]], axis=1)
    thr_w[0, 1] = 0.0                                   # job 0 cannot run on type 1
    m = sf_.HostSim(lib, tr, 4, 120.0, 120.0)
    m.set_worker_types(thr_w, [2, 2])
    m.begin()
    assert m.step({0: 0, 1: 0, 2: 0}).err & 1            # three gangs on two workers of type 0
    m = sf_.HostSim(lib, tr, 4, 120.0, 120.0)
    m.set_worker_types(thr_w, [2, 2])
    m.begin()
    assert m.step({0: 1}).err & 8                        # scheduler.py:1494-1504 raises


@pytest.mark.skipif(not rh.reference_available(), reason="staged reference not present")
def test_pins_are_what_the_reference_produces_now():
    from tests.golden import make_sim_hetero_pins as gen
    policy, keep, cluster = gen.RUNS[0]
    live = gen.record(policy, keep, cluster)
    rec = PINS[f"{policy}_{keep}_{cluster}"]
    assert live["makespan"] == rec["makespan"] and live["rounds"] == rec["rounds"]
    assert live["per_round_schedule"] == rec["per_round_schedule"] and live["jct"] == rec["jct"]
