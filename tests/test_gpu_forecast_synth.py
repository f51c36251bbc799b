"""Forecast kernels on synthetic job profiles well outside the canonical trace: long schedules (up to 6000
epochs), up to 8 batch-size modes, jobs arriving and leaving (slot reuse, pool growth), random throughput
timelines — against the numpy restatement (oracle/jobmeta.py, itself bit-identical to the reference's
JobMetaData) driven through the same call sequence."""
import numpy as np
import pytest

from oracle import jobmeta as ojm
from oracle import shockwave_milp as om
from shockwave_b200 import Engine, make_params
from tests import fixtures as fx

pytestmark = pytest.mark.gpu


def _profile(rng, jid):
    E = int(rng.choice([3, 7, 40, 300, 2500, 6000]))
    nm = int(rng.integers(1, 9))
    modes = rng.choice([8, 16, 32, 64, 128, 256, 512, 1024, 2048], size=nm, replace=False)
    cuts = np.sort(rng.integers(0, E + 1, size=nm - 1)) if nm > 1 else np.array([], dtype=int)
    bs = np.empty(E, dtype=np.int64)
    for i, (a, b) in enumerate(zip(np.r_[0, cuts], np.r_[cuts, E])):
        bs[a:b] = modes[i]
    if rng.random() < 0.3:
        rng.shuffle(bs)
    pre = np.maximum(1.0, np.round(rng.uniform(0.6, 900.0, size=E) if rng.random() < 0.5
                                   else np.full(E, rng.uniform(1, 400))))
    return dict(nworkers=int(rng.choice([1, 2, 4, 8])), epochs=E, epoch_nsamples=float(rng.choice([50000, 10000, 117907])),
                pre=pre, bs=bs, timestamp_submit=float(rng.uniform(0, 5e4)))


def test_random_profiles_with_churn():
    rng = np.random.default_rng(77)
    eng = Engine(0)
    G, T, D = 64, 20, 120.0
    jobs, series, slots, free, nxt = {}, {}, {}, [], 0
    next_id, rnd = 0, 0
    worst = dict(dbar=0.0, rem=0.0, ftobj=0.0, bfkey=0.0)
    for step in range(40):
        # churn
        for jid in list(jobs):
            if rng.random() < 0.12:
                eng.job_remove(slots[jid]); free.append(slots.pop(jid)); del jobs[jid]; series.pop(jid, None)
        changed = False
        while len(jobs) < 30 or rng.random() < 0.3:
            p = _profile(rng, next_id)
            s = free.pop() if free else nxt
            if s == nxt:
                nxt += 1
            eng.job_add(s, p["nworkers"], p["epochs"], p["epoch_nsamples"], p["timestamp_submit"], p["pre"], p["bs"])
            jobs[next_id] = ojm.JobState(next_id, p["nworkers"], p["epochs"], p["epoch_nsamples"], p["pre"], p["bs"],
                                         p["timestamp_submit"], D)
            slots[next_id] = s
            next_id += 1
            changed = True
            if len(jobs) >= 60:
                break
        rnd += int(rng.integers(1, 5))
        live = list(jobs.keys())
        J = len(live)
        reest = changed or step == 0 or rng.random() < 0.3
        c = np.zeros(J, dtype=np.int32); ns = np.zeros(J); end = np.full(J, -1, dtype=np.int32)
        for k, jid in enumerate(live):
            job = jobs[jid]
            job.epoch_progress = int(rng.integers(0, job.epochs + 1)) if rng.random() < 0.9 else job.epochs
            c[k] = job.epoch_progress
            if rng.random() < 0.7:
                job.timeline[rnd] = (float(rng.uniform(0.05, 40.0)), int(rng.choice(job.modes)))
            summ = job.timeline_summary()
            if summ is not None:
                ns[k], end[k] = summ
        prm = make_params(G, T, D, 1e-3, 12.0, 1.0, fx.BASES, fx.ORIGIN, round_ptr=rnd)
        f = eng.forecast(prm, [slots[j] for j in live], c, ns, end, reest, D)
        fb = bool(rng.random() < 0.5)
        ncal = rng.integers(0, T + 1, size=J).astype(np.int32)
        want = {k: np.empty(J) for k in worst}
        for k, jid in enumerate(live):
            job = jobs[jid]
            if reest:
                series.setdefault(jid, []).append((rnd, job.finish_time_estimate(G, J)))
            want["ftobj"][k] = om.finish_time_momentumed_average(list(series[jid]), rnd)
            want["dbar"][k] = job.interpolate_epoch_duration()
            want["rem"][k] = job.remaining()
            keep = job.amp
            want["bfkey"][k] = job.remaining()      # key of the no-fallback continuation (what swb_forecast returns)
            job.amp = keep
            if fb:
                job.calibrate(); job.remaining()
            if job.epoch_progress < job.epochs:
                for _ in range(int(ncal[k])):
                    job.calibrate()
        eng.forecast_commit(fb, ncal)
        for key in worst:
            rel = np.abs(f[key] - want[key]) / np.maximum(1e-300, np.abs(want[key]))
            worst[key] = max(worst[key], float(rel.max()))
    print("worst relative deviation:", worst)
    for key, v in worst.items():
        assert v <= 1e-10, (key, v)
    eng.close()


def test_error_paths(engine):
    prm = make_params(8, 4, 120.0, 1e-3, 12.0, 1.0, fx.BASES, fx.ORIGIN)
    with pytest.raises(RuntimeError):
        engine.forecast(prm, [9999], [0], [0.0], [-1], True, 120.0)           # unknown slot
    engine.job_add(0, 1, 5, 100.0, 0.0, np.ones(5), np.full(5, 32))
    with pytest.raises(RuntimeError):
        engine.forecast(prm, [0], [6], [0.0], [-1], True, 120.0)              # progress > epochs
    with pytest.raises(RuntimeError):
        engine.job_add(1, 300, 5, 100.0, 0.0, np.ones(5), np.full(5, 32))     # gang wider than 255
    with pytest.raises(RuntimeError):
        engine.solve(make_params(8, 200, 120.0, 1e-3, 12.0, 1.0, fx.BASES, fx.ORIGIN), [1], [5], [0], [10.0], [50.0], [1e9])
    engine.job_remove(0)
