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
