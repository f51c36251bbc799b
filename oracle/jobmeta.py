"""CPU ORACLE (test infrastructure, not product code): the dynamic-adaptation forecast.

Plain-Python/numpy restatement of the parts of `scheduler/JobMetaData.py` that feed the solver:
  * `calibrate_profiled_epoch_duration`         JobMetaData.py:225-288
  * `get_bs_epoch_duration_map`                 JobMetaData.py:301-313
  * `dirichlet_posterior_remaining_runtime`     JobMetaData.py:315-370
  * `interpolate_epoch_duration`                shockwave.py:322-324
  * the finish-time estimate of `finish_time_uniform_share`   shockwave.py:88-120

It works on plain arrays (no reference objects) so that it travels to the GPU box, where
`/root/reference` does not exist.  Pinned: `tests/test_oracle_jobmeta.py` checks it value-for-value
against the reference's own `JobMetaData` (imported from /root/reference when present) and against
the committed fixtures `tests/golden/jobmeta_*.npz` made by `tests/golden/make_jobmeta_golden.py`.

Reference behaviour worth knowing (kept, because parity is judged against it):
  * calibration is STATEFUL: `epoch_duration[i] = preprofiled[i] * amp` only changes when the
    measured/profiled sample counts differ by > 40 %, and the in-epoch deficit term divides by the
    CURRENT (already rescaled) `epoch_duration[iepoch]`, so repeated calls are not idempotent.  We
    keep one scalar `amp` per job (exact, because every rescale multiplies the untouched
    pre-profiled copy) and replay the reference's call sequence.
"""
from __future__ import annotations

import numpy as np


class JobState:
    """Array view of one JobMetaData (JobMetaData.py:41-98)."""

    def __init__(self, jobid, nworkers, epochs, epoch_nsamples, epoch_duration_pre, bs_schedule,
                 timestamp_submit=0.0, gavel_round_duration=None):
        self.jobid = jobid
        self.nworkers = int(nworkers)
        self.epochs = int(epochs)
        self.epoch_nsamples = epoch_nsamples
        # max(1, round(d)) then /overclock(1.0)  — JobMetaData.py:105-114 (callers pass that result)
        self.pre = np.asarray(epoch_duration_pre, dtype=np.float64)
        self.bs = np.asarray(bs_schedule, dtype=np.int64)
        assert len(self.pre) == len(self.bs) == self.epochs
        self.modes = sorted(set(int(b) for b in self.bs))          # JobMetaData.py:296
        self.prior = self.epochs / len(self.modes)                 # JobMetaData.py:297-299
        self.amp = 1.0
        self.epoch_progress = 0
        self.timestamp_submit = timestamp_submit
        self.timeline = {}          # round -> (throughput, bs); shared dict in the reference
        self.gavel_round_duration = gavel_round_duration

    # --- JobMetaData.py:225-288 -------------------------------------------------------------
    def timeline_summary(self):
        """(measured_nsamples, end_round) of the throughput timeline (JobMetaData.py:235-249)."""
        if len(self.timeline) == 0:
            return None
        prev = 0
        nsamp = 0
        for cur in sorted(self.timeline.keys()):
            thr, bs = self.timeline[cur][0], self.timeline[cur][1]
            niters = thr * self.gavel_round_duration * (cur - prev)
            nsamp += bs * niters
            prev = cur
        return nsamp, max(self.timeline.keys())

    def calibrate(self):
        summ = self.timeline_summary()
        if summ is None:
            return
        measured_nsamples, end_round = summ
        measured_time_range = self.gavel_round_duration * end_round
        pre_range = 0
        pre_nsamples = 0
        iepoch = 0
        for iepoch, duration in enumerate(self.pre):
            if pre_range + duration > measured_time_range:
                break
            pre_range += duration
            pre_nsamples += self.epoch_nsamples
        deficit = measured_time_range - pre_range
        if deficit > 0:
            epoch_duration = self.pre[iepoch] * self.amp      # the CURRENT (rescaled) duration
            pre_nsamples += self.epoch_nsamples * deficit / epoch_duration
        if (measured_nsamples <= 0 or pre_nsamples <= 0
                or abs(measured_nsamples - pre_nsamples) / pre_nsamples <= 0.4):
            return
        self.amp = pre_nsamples / measured_nsamples

    # --- shockwave.py:322-324 ----------------------------------------------------------------
    def interpolate_epoch_duration(self):
        self.calibrate()
        return float(np.mean(self.pre[: self.epoch_progress + 1] * self.amp))

    def elapsed(self):
        """sum(job.epoch_duration[:epoch_progress])   (shockwave.py:106)."""
        return float(sum((self.pre[: self.epoch_progress] * self.amp).tolist()))

    # --- JobMetaData.py:315-370 --------------------------------------------------------------
    def remaining(self, progress=None):
        if progress is None:
            progress = self.epoch_progress
        assert 0 <= progress <= self.epochs
        observed = self.bs[: progress + 1]
        post = {m: self.prior for m in self.modes}
        for b in observed:
            post[int(b)] += 1
        csum = sum(post.values())
        reb = {m: self.epochs * v / csum for m, v in post.items()}
        for b in observed:
            if reb[int(b)] >= 1:
                reb[int(b)] -= 1
        inflated = int(sum(reb.values()) + 1)
        rem_epochs = self.epochs - self.epoch_progress
        if inflated < rem_epochs:
            inflated = rem_epochs
        if inflated <= 0 or rem_epochs <= 0:
            return 1.0
        self.calibrate()                                   # get_bs_epoch_duration_map, :302
        dur = self.pre * self.amp
        out = 0.0
        for m in self.modes:
            out += reb[m] * float(np.mean(dur[self.bs == m]))
        out *= rem_epochs / inflated
        return out

    # --- shockwave.py:101-112 ----------------------------------------------------------------
    def finish_time_estimate(self, ngpus, njobs):
        share = min(1.0, ngpus / njobs)
        self.calibrate()
        return self.timestamp_submit + (self.elapsed() + self.remaining(self.epoch_progress)) / share
