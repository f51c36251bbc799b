"""gavel.cu (swb_gavel_round: priorities -> sorted queues -> greedy selection -> worker assignment in one launch) against
oracle/gavel_round.py, which is pinned on the unmodified reference's own round loop (tests/test_oracle_gavel_round.py).
Integer / comparison work: everything must be bit-for-bit — priorities as float64 bit patterns, selections in order,
assignments with worker ids in the reference's OrderedDict insertion order."""
import os
import pickle

import numpy as np
import pytest

from oracle import gavel_round as gr

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _check(engine, s):
    prio, sel, asg = engine.gavel_round(s["alloc"], s["job_time"], s["worker_time"], s["thr"], s["deficit"], s["sf"],
                                        s["capacity"], s["type_order"], s["worker_lists"], s["prev"],
                                        isolated_plus=s["isolated_plus"], fifo=s["fifo"])
    assert np.array_equal(prio.view(np.uint64), np.asarray(s["prio"], dtype=np.float64).view(np.uint64))
    for t in s["type_order"]:
        assert sel[t] == list(s["sel"][t]), (t, sel[t][:10], list(s["sel"][t])[:10])
    assert asg == [(j, tuple(w)) for j, w in s["asg"]]


def test_recorded_states_of_the_canonical_run(engine):
    """every third round of the reference's max_min_fairness run on the canonical 120-job trace (93 states)"""
    states = pickle.load(open(os.path.join(HERE, "golden", "gavel_round_states.pkl"), "rb"))
    assert len(states) > 50
    for s in states:
        _check(engine, s)


def _random_state(rng, J, W, cap, isolated_plus=False, fifo=False):
    order = list(rng.permutation(W))
    alloc = rng.uniform(0, 1, (J, W)) * (rng.random((J, W)) < 0.8)
    alloc[rng.random(J) < 0.1] = np.nan                      # jobs the allocation does not know yet
    alloc = np.round(alloc, int(rng.integers(1, 4)))          # few distinct values: ties everywhere
    job_time = rng.uniform(0, 500, (J, W)) * (rng.random((J, W)) < 0.7)
    worker_time = rng.uniform(100, 5000, W) * (rng.random(W) < 0.9)
    thr = rng.uniform(0.1, 10, (J, W)) * (rng.random((J, W)) < 0.9)
    deficit = np.round(rng.uniform(-5, 5, (J, W)), 1) * (rng.random((J, W)) < 0.3)
    sf = rng.choice([1, 2, 4, 8], J, p=[0.6, 0.3, 0.09, 0.01]).astype(np.int32)
    capacity = np.array(cap, dtype=np.int32)
    wid0 = np.concatenate([[0], np.cumsum(capacity)])
    lists_by_type = [list(rng.permutation(np.arange(wid0[t], wid0[t + 1])).tolist()) for t in range(W)]
    worker_lists = [lists_by_type[t] for t in order]
    # previous round: a random feasible assignment (disjoint workers)
    prev = {}
    for t in range(W):
        free = list(rng.permutation(lists_by_type[t]))
        for j in rng.permutation(J):
            if len(free) < sf[j] or rng.random() < 0.6 or j in prev:
                continue
            prev[int(j)] = (t, tuple(int(free.pop()) for _ in range(sf[j])))
    servers = {int(t): [list(lists_by_type[t])] for t in range(W)}
    prio, sel, asg = gr.gavel_round(alloc, job_time, worker_time, thr, deficit, sf, capacity, [int(t) for t in order],
                                    servers, prev, isolated_plus=isolated_plus, fifo=fifo)
    return dict(alloc=alloc, job_time=job_time, worker_time=worker_time, thr=thr, deficit=deficit, sf=sf,
                capacity=capacity, type_order=[int(t) for t in order], worker_lists=worker_lists, prev=prev,
                isolated_plus=isolated_plus, fifo=fifo, prio=prio, sel=sel, asg=list(asg.items()))


@pytest.mark.parametrize("J,W,cap", [(7, 1, [4]), (40, 3, [8, 6, 5]), (300, 3, [64, 32, 16]), (1000, 2, [256, 100]),
                                     (2048, 3, [256, 128, 128]), (4096, 1, [512]), (513, 6, [40, 30, 20, 10, 8, 4])])
def test_random_states(engine, J, W, cap):
    rng = np.random.default_rng(J * 31 + W)
    for rep in range(4):
        for flags in ((False, False), (True, False), (False, True)):
            try:
                s = _random_state(rng, J, W, cap, *flags)
            except RuntimeError:
                continue                                     # the reference itself raises on this state
            _check(engine, s)


def test_unassignable_selection_raises_like_the_reference(engine):
    # a job the allocation knows is selected but its workers are gone: "Could not assign workers to job"
    alloc = np.array([[1.0], [1.0]]); z = np.zeros((2, 1))
    with pytest.raises(RuntimeError, match="assign workers"):
        engine.gavel_round(alloc, z, [0.0], np.ones((2, 1)), z, [2, 2], [4], [0], [[0, 1, 2]], {})
