"""CPU tests: the C-ABI library loads and exports every symbol include/swb200.h declares (no compute
without a GPU), struct layouts agree, and the host state machine of the drop-in class."""
import ctypes as C
import os
import re
from collections import OrderedDict

import numpy as np
import pytest

from shockwave_b200 import engine
from shockwave_b200.scheduler import ShockwaveScheduler, schedules_from_matrices

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "swb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(swb_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    names = _header_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(engine.EXPORTS) == names


def test_struct_layouts_match_header():
    # sizes implied by include/swb200.h (4 int32, 4 double, 2 x 16 double ; 6 int32 + 5 double)
    assert C.sizeof(engine.Params) == 16 + 32 + 2 * 16 * 8
    assert C.sizeof(engine.Result) == 24 + 5 * 8
    assert engine.load_library().swb_version() >= 101


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        engine.Engine(0)


def test_make_params():
    p = engine.make_params(32, 20, 120, 1e-3, 12.0, 1.0, [0.0, 0.2, 0.4, 0.6, 0.8, 1.0], {0.0: 1e-6}, 7)
    assert (p.ngpus, p.future_rounds, p.round_ptr, p.nbases) == (32, 20, 7, 6)
    assert abs(p.logv[0] - np.log(1e-6)) < 1e-15 and p.logv[5] == 0.0


class _FakeBackend(ShockwaveScheduler):
    """State machine only: the three backend hooks are replaced, nothing touches the GPU."""
    def _on_add(self, jobid, job): self.log.append(("add", jobid))
    def _on_remove(self, jobid): self.log.append(("rm", jobid))
    def _resolve(self, jobids, jobobjs):
        self.log.append(("solve", tuple(jobids), self.round_ptr, self.reestimate_share))
        return OrderedDict((self.round_ptr + t, list(jobids)) for t in range(self.future_nrounds))


class _J:
    def __init__(self): self.epoch_progress, self.waiting_delay, self.epochs = 0, 0, 10
    def set_epoch_progress(self, c): self.epoch_progress = c
    def reset_waiting_delay(self): self.waiting_delay = 0
    def add_waiting_delay(self, d): self.waiting_delay += d


def _mk():
    _FakeBackend.log = []
    return _FakeBackend(8, 16, OrderedDict(), 4, 120, ["GUROBI"], 1e-3, 24, 15, 30, [0.0, 1.0], {0.0: 1e-6}, 1e-3, 12.0, 1.0)


def test_resolve_and_cache_replay_state_machine():
    """shockwave.py:122-210: resolve set by add/remove/set_resolve, cleared by a solve; cached window
    replayed while round_ptr is inside it; reestimate_share only on add/remove."""
    sw = _mk()
    sw.add_metadata(1, _J()); sw.add_metadata(2, _J())
    assert sw.resolve and sw.reestimate_share
    assert sw.round_schedule() == [1, 2]
    assert not sw.resolve and not sw.reestimate_share
    n = len(sw.log)
    for _ in range(3):
        sw.increment_round_ptr()
        assert sw.round_schedule() == [1, 2]          # replay, no solve
    assert len(sw.log) == n
    sw.increment_round_ptr()                          # round 4 is outside the cached window [0,4)
    sw.round_schedule()
    assert sw.log[-1][0] == "solve" and sw.log[-1][3] is False
    sw.set_resolve(); sw.round_schedule()
    assert sw.log[-1][2] == 4 and sw.log[-1][3] is False
    sw.remove_metadata(1)
    assert sw.resolve and sw.reestimate_share and 1 in sw.completed_jobs
    assert sw.round_schedule() == [2] and sw.log[-1][3] is True
    with pytest.raises(AssertionError):
        sw.add_metadata(2, _J())
    sw.schedule_progress(2, 5); assert sw.metadata[2].epoch_progress == 5
    sw.deschedule_waiting_delay(2, 120); assert sw.metadata[2].waiting_delay == 120
    sw.deschedule_waiting_delay(99, 120)              # unknown id is ignored (shockwave.py:189-194)


def test_constructor_asserts_like_reference():
    with pytest.raises(AssertionError):
        _FakeBackend(0, 16, OrderedDict(), 4, 120, [], 1e-3, 24, 15, 30, [0.0, 1.0], {0.0: 1e-6}, 1e-3, 12.0, 1.0)
    with pytest.raises(AssertionError):
        _FakeBackend(8, 16, {}, 4, 120, [], 1e-3, 24, 15, 30, [0.0, 1.0], {0.0: 1e-6}, 1e-3, 12.0, 1.0)


def test_schedules_from_matrices_order():
    x = np.array([[1, 0], [0, 0], [0, 1], [0, 0]], dtype=np.uint8)
    bf = np.array([[0, 0], [1, 1], [0, 0], [1, 1]], dtype=np.uint8)
    s = schedules_from_matrices(x, bf, np.array([5.0, 9.0, 1.0, 9.0]), [10, 11, 12, 13], 7)
    assert s[7] == [10, 11, 13] and s[8] == [12, 11, 13]
