"""Replays the recorded canonical simulation (tests/golden/tacc32_solves.npz) at the forecast level:
the same per-solve host inputs the ShockwaveScheduler boundary saw, in order, so that the stateful
parts (calibration factor, share series) evolve exactly as they did under the reference's own
JobMetaData objects."""
import numpy as np

from oracle import jobmeta as ojm
from oracle import shockwave_milp as om
from tests import fixtures as fx


def replay_with_oracle_jobmeta(upto=None):
    """Yields (solve_index, recorded, dict(dbar, rem, ftobj, bfkey)) computed by oracle/jobmeta.py with
    the call sequence the device kernel replays (forecast.cu header)."""
    st = fx.job_statics()
    G = fx.TACC["G"]
    jobs, series = {}, {}
    n = fx.n_solves() if upto is None else min(upto, fx.n_solves())
    for i in range(n):
        s = fx.solve(i)
        J, r = s["J"], s["round_ptr"]
        live = [int(j) for j in s["jobids"]]
        for jid in list(jobs):
            if jid not in live:
                del jobs[jid]; series.pop(jid, None)
        out = dict(dbar=np.empty(J), rem=np.empty(J), ftobj=np.empty(J), bfkey=np.empty(J), rem_fb=np.empty(J))
        ncal = ncal_of(s["x"], s["g"], G)
        for k, jid in enumerate(live):
            if jid not in jobs:
                p = st[jid]
                jobs[jid] = ojm.JobState(jid, p["nworkers"], p["epochs"], p["epoch_nsamples"], p["pre"], p["bs"],
                                         p["timestamp_submit"], p["grd"])
            job = jobs[jid]
            job.epoch_progress = int(s["c"][k])
            ns, end = float(s["meas_ns"][k]), int(s["meas_end"][k])
            job.timeline_summary = (lambda ns=ns, end=end: None if end < 0 else (ns, end))
            if s["reestimate"]:
                series.setdefault(jid, []).append((r, job.finish_time_estimate(G, J)))
            out["ftobj"][k] = om.finish_time_momentumed_average(list(series[jid]), r)
            out["dbar"][k] = job.interpolate_epoch_duration()
            out["rem"][k] = job.remaining()
            if s["status"] == om.STATUS_FALLBACK:
                job.calibrate()
                out["rem_fb"][k] = job.remaining()
            else:
                out["rem_fb"][k] = out["rem"][k]
            # sort key of construct_schedules: one dirichlet() per (round with idle GPUs, unscheduled
            # job) — each one calibrates (shockwave.py:254-267).  The key reported is the first one.
            keep = job.amp
            out["bfkey"][k] = job.remaining()
            job.amp = keep
            if job.epoch_progress < job.epochs:
                for _ in range(int(ncal[k])):
                    job.calibrate()
        out["ncal"] = ncal
        yield i, s, out


def ncal_of(x, g, G):
    """rounds with idle GPUs in which job j is not scheduled by the solver (per job)."""
    x = np.asarray(x, dtype=np.int64)
    idle = G - x.T @ np.asarray(g, dtype=np.int64)
    return ((x == 0) & (idle[None, :] > 0)).sum(axis=1).astype(np.int32)
