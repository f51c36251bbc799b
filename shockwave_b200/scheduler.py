"""Drop-in `ShockwaveScheduler` backed by the sm_100a kernels of libswb200.so.

Mirrors the public surface of the reference class (scheduler/shockwave.py:20-210) that
`scheduler/scheduler.py` drives: same constructor signature, same attributes read
from outside (`metadata`, `round_duration`), same methods and the same resolve / cache-replay state
machine.  `JobMetaData` objects are built and mutated by the caller (scheduler.py:757-778) and are
only READ here, at solve time.

What moved to the GPU: the per-job forecast walk (calibration, Dirichlet remaining-runtime,
interpolated epoch duration, finish-time share series), the whole model-build-and-solve, the
placement into rounds and the work-conserving back-fill.  What stays on the host: this bookkeeping
and the conversion of the returned J x T byte matrices into the reference's
`OrderedDict{round -> [job ids]}`.
"""
from __future__ import annotations

import operator
import random
from collections import OrderedDict
from itertools import repeat

import numpy as np

from . import engine as _eng


class ShockwaveScheduler(object):
    def __init__(self, ngpus: int, gram: int, init_metadata: OrderedDict, future_nrounds: int,
                 round_duration: int, solver_preference: list, solver_rel_gap: float,
                 solver_num_threads: int, solver_timeout: float, n_epoch_vars_max: int,
                 logapx_bases: list, logapx_origin: dict, k: float, lam: float, rhomax: float,
                 device: int = 0):
        # same argument checks as shockwave.py:40-69
        self.ngpus = ngpus
        self.gram = gram
        assert self.ngpus > 0
        assert self.gram > 0
        self.future_nrounds = future_nrounds
        assert self.future_nrounds > 0
        self.round_duration = round_duration
        assert self.round_duration > 0
        self.solver_preference = solver_preference      # kept for signature parity; Gurobi is gone
        self.solver_rel_gap = solver_rel_gap
        self.solver_num_threads = solver_num_threads
        self.solver_timeout = solver_timeout
        assert self.solver_timeout > 0
        self.n_epoch_vars_max = n_epoch_vars_max
        assert type(self.n_epoch_vars_max) == int and self.n_epoch_vars_max > 0
        self.logapx_bases = logapx_bases
        assert type(self.logapx_bases) == list
        self.logapx_origin = logapx_origin
        assert type(self.logapx_origin) == dict
        self.k = k
        assert self.k > 0
        self.lam = lam
        self.rhomax = rhomax

        self._device = device
        self._engine = None
        self._slots = {}          # jobid -> slot of the device-resident job table
        self._free_slots = []
        self._next_slot = 0
        # per-slot cache of the throughput-timeline summaries (see _timeline_summaries)
        self._tl_len = np.full(64, -1, dtype=np.int64)      # signature: number of entries ...
        self._tl_key = np.full(64, -2, dtype=np.int64)      # ... last round ...
        self._tl_val = np.empty(64, dtype=object)           # ... and its value
        self._tl_ns = np.zeros(64, dtype=np.float64)        # summary: measured samples
        self._tl_end = np.full(64, -1, dtype=np.int32)      # summary: last measured round
        self._slot_arr = None     # slots in metadata order, rebuilt after add / remove
        self.last_result = None   # scalars of the latest solve (objective, status, ...)

        assert type(init_metadata) == OrderedDict
        self.metadata = OrderedDict()
        for jobid, jobobj in init_metadata.items():
            self.add_metadata(jobid, jobobj)

        self.schedules = OrderedDict()
        self.round_ptr = 0
        self.resolve = True
        self.completed_jobs = OrderedDict()
        self.reestimate_share = True
        self.share_series = {}    # lives on the device (swb_ctx share-series state); kept for parity

    # ---- backend hooks (the oracle harness overrides these three, nothing else) -------------------
    def _eng(self):
        if self._engine is None:
            self._engine = _eng.Engine(self._device)   # raises when the CUDA library / GPU is missing
        return self._engine

    def _on_add(self, jobid, job):
        slot = self._free_slots.pop() if self._free_slots else self._next_slot
        if slot == self._next_slot:
            self._next_slot += 1
        ts = job.timestamp_submit
        self._eng().job_add(slot, job.nworkers, job.epochs, job.epoch_nsamples,
                            float("nan") if ts is None else ts,
                            job.epoch_duration_preprofiled, job.bs_schedule)
        self._slots[jobid] = slot
        self._slot_arr = None
        if slot >= len(self._tl_len):
            grow = max(2 * len(self._tl_len), slot + 1)
            for name, fill in (("_tl_len", -1), ("_tl_key", -2), ("_tl_val", None), ("_tl_ns", 0.0), ("_tl_end", -1)):
                old = getattr(self, name)
                new = np.empty(grow, dtype=old.dtype)
                new[:len(old)] = old
                new[len(old):] = fill
                setattr(self, name, new)
        self._tl_len[slot] = -1           # a new tenant of the slot: force a recompute

    def _on_remove(self, jobid):
        slot = self._slots.pop(jobid)
        self._eng().job_remove(slot)
        self._free_slots.append(slot)
        self._slot_arr = None

    def _timeline_summary(self, jobid, job):
        """(measured_nsamples, end_round) of JobMetaData.py:235-249 for one job."""
        tl = job.throughput_measurements
        assert tl is not None                       # JobMetaData.py:229
        if len(tl) == 0:
            return 0.0, -1
        grd = job.gavel_round_duration
        prev = 0
        nsamp = 0
        for cur in sorted(tl.keys()):
            thr, bs = tl[cur][0], tl[cur][1]
            nsamp += bs * (thr * grd * (cur - prev))
            prev = cur
        return float(nsamp), int(max(tl.keys()))

    def _timeline_summaries(self, jobids, jobobjs, slots):
        """Vector form, cached per slot while the job's shared throughput OrderedDict is unchanged (signature =
        length + last (round, value) entry; the caller only appends / rewrites the current round,
        scheduler.py:568-571).  The per-job work runs in C (map / numpy); Python only touches the jobs whose
        timeline changed since the previous re-solve."""
        J = len(jobids)
        tls = list(map(operator.attrgetter("throughput_measurements"), jobobjs))
        assert None not in tls                      # JobMetaData.py:229
        lens = np.fromiter(map(len, tls), dtype=np.int64, count=J)
        lastk_l = list(map(next, map(reversed, tls), repeat(-1)))
        lastk = np.fromiter(lastk_l, dtype=np.int64, count=J)
        lastv = list(map(dict.get, tls, lastk_l))
        changed = (lens != self._tl_len[slots]) | (lastk != self._tl_key[slots])
        changed |= np.fromiter(map(operator.ne, lastv, self._tl_val[slots]), dtype=bool, count=J)
        for i in np.flatnonzero(changed).tolist():
            sl = slots[i]
            ns_i, end_i = self._timeline_summary(jobids[i], jobobjs[i])
            self._tl_len[sl], self._tl_key[sl], self._tl_val[sl] = lens[i], lastk[i], lastv[i]
            self._tl_ns[sl], self._tl_end[sl] = ns_i, end_i
        return self._tl_ns[slots], self._tl_end[slots]

    def _resolve(self, jobids, jobobjs):
        """One re-solve on the device; returns OrderedDict{round -> [job ids]} (shockwave.py:129-161)."""
        J = len(jobids)
        if J == 0:      # nothing to schedule: an empty window (the reference is never called like this)
            return OrderedDict((self.round_ptr + t, []) for t in range(self.future_nrounds))
        if self._slot_arr is None or len(self._slot_arr) != J:
            self._slot_arr = np.fromiter(map(self._slots.__getitem__, jobids), dtype=np.int32, count=J)
        slots = self._slot_arr
        prog = np.fromiter(map(operator.attrgetter("epoch_progress"), jobobjs), dtype=np.int32, count=J)
        ns, end = self._timeline_summaries(jobids, jobobjs, slots)
        prm = _eng.make_params(self.ngpus, self.future_nrounds, self.round_duration, self.k, self.lam,
                               self.rhomax, self.logapx_bases, self.logapx_origin, self.round_ptr)
        grd = jobobjs[0].gavel_round_duration
        out = self._eng().round_solve(prm, slots, prog, ns, end, self.reestimate_share, grd,
                                      want_forecast=True)
        self.last_result = out["result"]
        self.last_forecast = {k: out[k] for k in ("dbar", "rem", "ftobj", "bfkey")}
        return schedules_from_matrices(out["x"], out["backfill"], out["bfkey"], jobids, self.round_ptr)

    # ---- reference surface ------------------------------------------------------------------------
    def round_schedule(self):
        if not self.resolve:
            if len(self.schedules) > 0:
                if self.round_ptr in self.schedules.keys():
                    return self.schedules[self.round_ptr]
        jobids = list(self.metadata.keys())
        jobobjs = list(self.metadata.values())
        # the reference re-seeds both global RNGs on every solver call (call_cvxpy_solver,
        # shockwave.py:451-452); the simulator draws from them afterwards, so this side effect is part
        # of the drop-in contract
        random.seed(0)
        np.random.seed(0)
        schedules = self._resolve(jobids, jobobjs)
        self.reestimate_share = False               # shockwave.py:120
        self.schedules = schedules
        self.clear_resolve()
        return self.schedules[self.round_ptr]

    def increment_round_ptr(self):
        self.round_ptr += 1

    def set_resolve(self):
        self.resolve = True

    def clear_resolve(self):
        self.resolve = False

    def schedule_progress(self, jobid, epoch_progress, share_update=True):
        assert jobid in self.metadata.keys()
        job = self.metadata[jobid]
        job.set_epoch_progress(epoch_progress)
        job.reset_waiting_delay()

    def deschedule_waiting_delay(self, jobid, delay):
        if jobid in self.metadata.keys():
            self.metadata[jobid].add_waiting_delay(delay)

    def add_metadata(self, jobid, jobobj, share_update=True):
        assert jobid not in self.metadata.keys()
        self.metadata[jobid] = jobobj
        self._on_add(jobid, jobobj)
        self.set_resolve()
        if share_update:
            self.reestimate_share = True

    def remove_metadata(self, jobid, share_update=True):
        assert jobid not in self.completed_jobs.keys()
        self.completed_jobs[jobid] = self.metadata[jobid]
        if share_update:
            self.reestimate_share = True
        assert jobid in self.metadata.keys()
        self.metadata.pop(jobid)
        self._on_remove(jobid)
        self.set_resolve()


class LazySchedules(OrderedDict):
    """OrderedDict{round -> [job ids]} whose lists are built on first access.

    The reference materialises all T lists in construct_schedules (shockwave.py:233-283), but only
    `schedules[round_ptr]` is ever read (shockwave.py:127,166) and a window is usually replaced by the next
    re-solve after one or a few rounds; at 4096 jobs x 64 rounds building every list costs more host time
    than the whole GPU solve."""

    def __init__(self, x, backfill, bfkey, jobids, round_ptr):
        super().__init__()
        self._x, self._bf, self._ids, self._r0 = x, backfill, list(jobids), round_ptr
        self._order = np.argsort(-np.asarray(bfkey, dtype=np.float64), kind="stable")
        for t in range(x.shape[1]):
            super().__setitem__(round_ptr + t, None)

    def __getitem__(self, rnd):
        cur = super().__getitem__(rnd)
        if cur is None:
            t = rnd - self._r0
            ids = self._ids
            cur = [ids[j] for j in np.flatnonzero(self._x[:, t]).tolist()]
            order = self._order
            cur += [ids[j] for j in order[np.flatnonzero(self._bf[order, t])].tolist()]
            super().__setitem__(rnd, cur)
        return cur

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]


def schedules_from_matrices(x, backfill, bfkey, jobids, round_ptr):
    """J x T byte matrices -> {round_ptr+t -> [job ids]} in the reference's list order
    (construct_schedules, shockwave.py:233-283): solver-scheduled jobs in metadata order, then the
    back-filled ones in descending remaining-runtime order (stable)."""
    return LazySchedules(x, backfill, bfkey, jobids, round_ptr)
