"""The Monte-Carlo (GBM) forecast as part of the drop-in class (round-1 review: it was reachable only from bench.py).

No reference counterpart exists — parity is pinned where it can be: with zero drift and volatility
the GBM-enabled scheduler must reproduce the deterministic (Dirichlet, JobMetaData.py:315-370) path BIT FOR BIT through
round_schedule(); with drift only (sigma = 0) the forecast must equal the closed form R0 * mean_h exp(mu h); with
volatility the Monte-Carlo mean must agree with that closed form at O(P^-1/2)."""
from collections import OrderedDict

import numpy as np
import pytest

from shockwave_b200 import ShockwaveScheduler
from shockwave_b200.forecast_mc import analytic_mean
from tests import fixtures as fx

pytestmark = pytest.mark.gpu


class _Job:      # what scheduler.py hands over (scheduler/JobMetaData.py:41-98), duck-typed
    def __init__(self, jid, rng, noisy):
        E = int(rng.integers(20, 120))
        self.nworkers, self.epochs, self.epoch_nsamples = int(rng.choice([1, 1, 2, 4])), E, 50000
        bs = [32] * (E // 2) + [64] * (E - E // 2)
        base = {32: float(rng.uniform(80, 400)), 64: float(rng.uniform(60, 300))}
        self.epoch_duration_preprofiled = [base[b] * (float(rng.uniform(0.8, 1.2)) if noisy else 1.0) for b in bs]
        self.bs_schedule = bs
        self.timestamp_submit, self.gavel_round_duration = 0.0, 120.0
        self.throughput_measurements = OrderedDict()
        self.epoch_progress, self.waiting_delay = int(rng.integers(0, E // 2)), 0

    def set_epoch_progress(self, c): self.epoch_progress = c
    def reset_waiting_delay(self): self.waiting_delay = 0
    def add_waiting_delay(self, d): self.waiting_delay += d


def _sched(J, noisy, **kw):
    rng = np.random.default_rng(11)
    sw = ShockwaveScheduler(ngpus=32, gram=16, init_metadata=OrderedDict(), future_nrounds=20, round_duration=120,
                            solver_preference=["GUROBI"], solver_rel_gap=1e-3, solver_num_threads=24, solver_timeout=15,
                            n_epoch_vars_max=64, logapx_bases=fx.BASES, logapx_origin=fx.ORIGIN, k=1e-3, lam=12.0,
                            rhomax=1.0, **kw)
    for j in range(J):
        sw.add_metadata(j, _Job(j, rng, noisy))
    return sw


def _run(sw, rounds=4):
    out = []
    for r in range(rounds):
        sw.set_resolve()
        ids = sw.round_schedule()
        out.append((list(ids), {k: v.copy() for k, v in sw.last_forecast.items()}, dict(sw.last_result)))
        for j in ids:
            if j in sw.metadata:
                sw.schedule_progress(j, min(sw.metadata[j].epochs, sw.metadata[j].epoch_progress + 1))
        sw.increment_round_ptr()
    return out


def test_zero_volatility_reproduces_the_dirichlet_path_bit_for_bit():
    a = _run(_sched(60, noisy=False))
    b = _run(_sched(60, noisy=False, forecast="gbm", gbm_paths=512))             # sigma from the profile = 0
    c = _run(_sched(60, noisy=True, forecast="gbm", gbm_paths=512, gbm_volatility=0.0))
    d = _run(_sched(60, noisy=True))
    for x, y in ((a, b), (d, c)):
        for (ids1, f1, r1), (ids2, f2, r2) in zip(x, y):
            assert ids1 == ids2
            for k in f1:
                assert np.array_equal(f1[k], f2[k]), k
            assert r1 == r2


def test_drift_only_matches_the_closed_form_and_feeds_the_solve():
    mu = 2e-3
    det = _sched(60, noisy=False)
    gbm = _sched(60, noisy=False, forecast="gbm", gbm_paths=256, gbm_volatility=(mu, 0.0))
    det.round_schedule(); gbm.round_schedule()
    R0 = det.last_forecast["rem"]
    H = np.array([min(j.epochs - j.epoch_progress, 256) for j in det.metadata.values()])
    want = analytic_mean(R0, H, np.full(60, mu))
    got = gbm.last_forecast["rem"]
    assert np.allclose(got, want, rtol=2e-5)                 # fp32 path arithmetic
    assert (got > R0 * (1 + 1e-4)).any()
    # the longer forecast reaches the solve: the plan's makespan term grows with it
    assert gbm.last_result["makespan"] >= det.last_result["makespan"]
    assert gbm.last_result["objective"] != det.last_result["objective"]


def test_profile_spread_gives_a_volatility_and_the_mc_mean_converges():
    P = 16384
    det = _sched(40, noisy=True)
    gbm = _sched(40, noisy=True, forecast="gbm", gbm_paths=P, gbm_volatility=(1e-3, 0.03), gbm_horizon=64)
    det.round_schedule(); gbm.round_schedule()
    R0 = det.last_forecast["rem"]
    H = np.array([min(j.epochs - j.epoch_progress, 64) for j in det.metadata.values()])
    want = analytic_mean(R0, H, np.full(40, 1e-3))
    rel = np.abs(gbm.last_forecast["rem"] / want - 1.0)
    assert rel.max() < 6 * 0.03 * np.sqrt(64) / np.sqrt(P) + 1e-4, rel.max()
    # default model (None): sigma from the profile's spread inside the batch-size modes, > 0 for noisy profiles
    auto = _sched(40, noisy=True, forecast="gbm", gbm_paths=4096)
    auto.round_schedule()
    assert not np.array_equal(auto.last_forecast["rem"], R0)
    assert np.allclose(auto.last_forecast["rem"], R0, rtol=0.05)
