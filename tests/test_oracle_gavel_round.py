"""Pins oracle/gavel_round.py (the CPU restatement of Gavel's per-round priority -> selection -> worker-assignment step,
scheduler/scheduler.py:3669-3724, :1166-1258, :1306-1378, :1049-1110) on the reference itself: the UNMODIFIED reference
simulator runs twice on the canonical trace with the same policy backend — once as shipped, once with
shockwave_b200.placement.GavelRoundMixin in front of its Scheduler class and the oracle as the mixin's backend.
`per_round_schedule` must be IDENTICAL round by round: same jobs, same worker ids, same dict insertion order.  Against
the shipped golden pickle the number of rounds, the makespan and the average JCT must match (which of several
equal-priority jobs runs in a round depends on the last bits of the ECOS allocation the pickle was made with; HiGHS
gives the same allocation up to those bits).  Also records (state -> decision) pairs as the fixture of the GPU test."""
import glob
import os
import pickle

import numpy as np
import pytest

from oracle import gavel_backend as gb
from oracle import gavel_round as gr
from oracle import ref_harness as rh
from oracle.gavel_round_backend import OracleBackend
from shockwave_b200.placement import GavelRoundMixin

HERE = os.path.dirname(os.path.abspath(__file__))


def _mixin(backend):
    return type("OracleRoundMixin", (GavelRoundMixin,), {"_swb_backend": backend})


@pytest.mark.parametrize("policy", ["max_min_fairness", "gandiva_fair"])
def test_restated_round_step_reproduces_the_golden_schedule(policy):
    if not rh.reference_available():
        pytest.skip("reference tree not present")
    rec = []
    before = GavelRoundMixin.swb_round_calls
    with gb.cpu_backend() as P:
        out = rh.simulate(policy, policy_obj=P.get_policy(policy, solver="ECOS", seed=0),
                          scheduler_mixin=_mixin(OracleBackend(rec)))
    assert GavelRoundMixin.swb_round_calls - before == len(out["per_round_schedule"]) > 100
    with gb.cpu_backend() as P:
        ref = rh.simulate(policy, policy_obj=P.get_policy(policy, solver="ECOS", seed=0))
    assert len(out["per_round_schedule"]) == len(ref["per_round_schedule"])
    for r, (a, b) in enumerate(zip(out["per_round_schedule"], ref["per_round_schedule"])):
        assert list(a.items()) == list(b.items()), r          # same jobs, same workers, same insertion order
    assert out["makespan"] == ref["makespan"] and out["jct_list"] == ref["jct_list"]
    gold = pickle.load(open(glob.glob(os.path.join(rh.GOLDEN_DIR, policy + "_120_*"))[0], "rb"))
    assert len(out["per_round_schedule"]) == len(gold["per_round_schedule"])
    assert abs(out["makespan"] - gold["makespan"]) <= 1e-9 * gold["makespan"]
    assert abs(out["avg_jct"] - gold["avg_jct"]) <= 1e-9 * gold["avg_jct"]
    if policy == "max_min_fairness" and os.environ.get("SWB_WRITE_FIXTURES"):
        keep = rec[::3]
        with open(os.path.join(HERE, "golden", "gavel_round_states.pkl"), "wb") as f:
            pickle.dump(keep, f, protocol=4)


def test_oracle_selection_edge_cases():
    # Isolated_plus stops the whole walk at the first job that does not fit while workers are left; with none left
    # the reference only `continue`s (scheduler.py:1213-1214, :1243-1250)
    prio = np.array([[5.0], [4.0], [3.0], [2.0]]); z = np.zeros((4, 1)); one = np.ones((4, 1), dtype=bool)
    sf = np.array([2, 4, 1, 1])
    assert gr.select_jobs(prio, z, z, one, sf, [4], [0]) == {0: [0, 2, 3]}
    assert gr.select_jobs(prio, z, z, one, sf, [4], [0], isolated_plus=True) == {0: [0]}
    assert gr.select_jobs(prio, z, z, one, np.array([2, 2, 1, 1]), [4], [0], isolated_plus=True) == {0: [0, 1]}
    # FIFO skips non-positive priorities; zero throughput never runs
    thr = np.array([[True], [False], [True], [True]])
    assert gr.select_jobs(np.array([[1.0], [9.0], [0.0], [2.0]]), z, z, thr, np.ones(4, int), [4], [0], fifo=True) == {0: [3, 0]}
