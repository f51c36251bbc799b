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
                 device: int = 0, timeline_check: str = "dirty", forecast: str = "dirichlet",
                 gbm_paths: int = 8192, gbm_seed: int = 0, gbm_volatility=None, gbm_horizon: int = 256):
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
        # how changes of the caller-owned JobMetaData objects are noticed at re-solve time.  "dirty": only the jobs
        # touched through this class since the previous re-solve (add_metadata / schedule_progress — the simulator
        # appends a throughput measurement exactly for the jobs it then reports progress for, scheduler.py:555-571,
        # :2274-2341) are re-read; "exact": every job's timeline signature and epoch_progress are re-read each time
        # (for callers that mutate JobMetaData behind this class's back).  Identical results on the reference's loops.
        # forecast = "dirichlet": the reference's deterministic remaining-runtime estimate (JobMetaData.py:315-370);
        # "gbm": that estimate becomes the start of `gbm_paths` geometric-Brownian-motion sample paths per job on the
        # device (gbm.cu) and the solve plans against their mean.  gbm_volatility: None -> per-job sigma from the spread
        # of the profile inside its batch-size modes (0 for reference-generated profiles); a float -> that sigma for
        # every job; a (mu, sigma) pair; or a callable(jobid, JobMetaData) -> (mu, sigma).  Jobs with mu = sigma = 0
        # keep the deterministic value bit for bit.
        assert forecast in ("dirichlet", "gbm")
        self.forecast, self.gbm_paths, self.gbm_seed = forecast, int(gbm_paths), int(gbm_seed)
        self.gbm_volatility, self.gbm_horizon = gbm_volatility, int(gbm_horizon)
        assert timeline_check in ("dirty", "exact")
        self._timeline_check = timeline_check
        self._dirty = set()
        self._prog = np.zeros(64, dtype=np.int32)           # epoch_progress by slot (kept by schedule_progress)
        self._engine = None
        self._slots = {}          # jobid -> slot of the device-resident job table
        self._slot_job = {}       # slot -> jobid
        self._free_slots = []
        self._next_slot = 0
        # per-slot cache of the throughput-timeline summaries (see _timeline_summaries)
        self._tl_len = np.full(64, -1, dtype=np.int64)      # signature: number of entries ...
        self._tl_key = np.full(64, -2, dtype=np.int64)      # ... last round ...
        self._tl_val = np.empty(64, dtype=object)           # ... and its value
        self._tl_ns = np.zeros(64, dtype=np.float64)        # summary: measured samples
        self._tl_end = np.full(64, -1, dtype=np.int32)      # summary: last measured round
        self._tl_base = np.zeros(64, dtype=np.float64)      # running sum WITHOUT the last entry ...
        self._tl_prev = np.zeros(64, dtype=np.int64)        # ... and the round before the last one (incremental update)
        self._slot_arr = None     # slots in metadata order, rebuilt after add / remove
        self._ids_cache = None    # (job ids, job objects) in metadata order, rebuilt after add / remove
        self.last_result = None   # scalars of the latest solve (objective, status, ...)
        self.last_forecast = None

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
            if self.forecast == "gbm":
                self._engine.set_option(_eng.OPT_GBM_PATHS, self.gbm_paths)
                self._engine.set_option(_eng.OPT_GBM_SEED, self.gbm_seed)
                self._engine.set_option(_eng.OPT_GBM_HORIZON, self.gbm_horizon)
        return self._engine

    def _gbm_model(self, jobid, job):
        v = self.gbm_volatility
        if v is None:
            return None
        if callable(v):
            return tuple(float(a) for a in v(jobid, job))
        if isinstance(v, (tuple, list)):
            return float(v[0]), float(v[1])
        return 0.0, float(v)

    def _on_add(self, jobid, job):
        slot = self._free_slots.pop() if self._free_slots else self._next_slot
        if slot == self._next_slot:
            self._next_slot += 1
        ts = job.timestamp_submit
        self._eng().job_add(slot, job.nworkers, job.epochs, job.epoch_nsamples,
                            float("nan") if ts is None else ts,
                            job.epoch_duration_preprofiled, job.bs_schedule)
        if self.forecast == "gbm":
            ms = self._gbm_model(jobid, job)
            if ms is not None:
                self._eng().job_set_gbm(slot, ms[0], ms[1])
        self._slots[jobid] = slot
        self._slot_job[slot] = jobid
        self._slot_arr = None
        if slot >= len(self._tl_len):
            grow = max(2 * len(self._tl_len), slot + 1)
            for name, fill in (("_tl_len", -1), ("_tl_key", -2), ("_tl_val", None), ("_tl_ns", 0.0), ("_tl_end", -1),
                               ("_prog", 0), ("_tl_base", 0.0), ("_tl_prev", 0)):
                old = getattr(self, name)
                new = np.empty(grow, dtype=old.dtype)
                new[:len(old)] = old
                new[len(old):] = fill
                setattr(self, name, new)
        self._tl_len[slot] = -1           # a new tenant of the slot: force a recompute
        self._prog[slot] = job.epoch_progress
        self._dirty.add(slot)

    def _on_remove(self, jobid):
        slot = self._slots.pop(jobid)
        self._slot_job.pop(slot, None)
        self._dirty.discard(slot)
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

    def _touch_timeline(self, sl, jobid, job):
        """O(1) refresh of one job's summary after the caller appended (or rewrote) the measurement of the latest
        round: the reference's sum runs over the sorted rounds (JobMetaData.py:235-249), so continuing it with the new
        last term gives the same float; anything but "same entries" / "one entry appended at the end" falls back to
        the full walk."""
        tl = job.throughput_measurements
        assert tl is not None                       # JobMetaData.py:229
        n = len(tl)
        if n == 0:
            self._tl_len[sl], self._tl_ns[sl], self._tl_end[sl] = 0, 0.0, -1
            return
        lk = next(reversed(tl))
        lv = tl[lk]
        cl, ck = self._tl_len[sl], self._tl_key[sl]
        if n == cl and lk == ck:
            base, pk = self._tl_base[sl], int(self._tl_prev[sl])
        elif cl >= 0 and n == cl + 1 and (cl == 0 or (lk > ck and tl.get(int(ck)) == self._tl_val[sl])):
            base, pk = (self._tl_ns[sl], int(ck)) if cl > 0 else (0.0, 0)
        else:
            ns_i, end_i = self._timeline_summary(jobid, job)
            ks = sorted(tl.keys())
            if ks[-1] == lk:                        # insertion order == sorted order: the cache can continue from here
                last = tl[lk]
                pk = ks[-2] if n > 1 else 0
                base = 0
                prev = 0
                for cur in ks[:-1]:
                    base += tl[cur][1] * (tl[cur][0] * job.gavel_round_duration * (cur - prev))
                    prev = cur
                self._tl_len[sl], self._tl_key[sl], self._tl_val[sl] = n, lk, last
                self._tl_base[sl], self._tl_prev[sl] = float(base), pk
            else:
                self._tl_len[sl] = -1
            self._tl_ns[sl], self._tl_end[sl] = ns_i, end_i
            return
        ns = base + lv[1] * (lv[0] * job.gavel_round_duration * (lk - pk))
        self._tl_len[sl], self._tl_key[sl], self._tl_val[sl] = n, lk, lv
        self._tl_base[sl], self._tl_prev[sl] = base, pk
        self._tl_ns[sl], self._tl_end[sl] = ns, lk

    def _timeline_summaries(self, jobids, jobobjs, slots):
        """Vector form, cached per slot while the job's shared throughput OrderedDict is unchanged (signature =
        length + last (round, value) entry; the caller only appends / rewrites the current round,
        scheduler.py:568-571).  The per-job work runs in C (map / numpy); Python only touches the jobs whose
        timeline changed since the previous re-solve."""
        J = len(jobids)
        if self._timeline_check == "dirty":
            for sl in self._dirty:                  # jobs added since the previous re-solve
                jid = self._slot_job.get(sl)
                if jid is not None:
                    self._touch_timeline(sl, jid, self.metadata[jid])
            self._dirty.clear()
            return self._tl_ns[slots], self._tl_end[slots]
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
        if self._timeline_check == "dirty":
            prog = self._prog[slots]
        else:
            prog = np.fromiter(map(operator.attrgetter("epoch_progress"), jobobjs), dtype=np.int32, count=J)
        ns, end = self._timeline_summaries(jobids, jobobjs, slots)
        prm = _eng.make_params(self.ngpus, self.future_nrounds, self.round_duration, self.k, self.lam,
                               self.rhomax, self.logapx_bases, self.logapx_origin, self.round_ptr)
        grd = jobobjs[0].gavel_round_duration
        out = self._eng().round_solve(prm, slots, prog, ns, end, self.reestimate_share, grd,
                                      want_forecast=True, packed=True)
        self.last_result = out["result"]
        self.last_forecast = {k: out[k] for k in ("dbar", "rem", "ftobj", "bfkey")}
        return LazySchedules(None, None, out["bfkey"], jobids, self.round_ptr, xmask=out["xmask"],
                             bfmask=out["bfmask"], T=self.future_nrounds)

    # ---- extension: planning on several worker types / capacities that change inside the window ---------------
    def heterogeneous_plan(self, type_speed, capacity, full_iters=400, coarse_iters=1000):
        """Fractional plan x[job][worker type][round] of the live jobs on W worker types (market.py: the dense
        price-response solve on the device).  type_speed [W]: progress relative to the type the epoch durations were
        profiled on; capacity [W][T]: workers of type w available in round round_ptr + t.  Uses the forecast of the
        latest re-solve (mean epoch duration, remaining runtime: shockwave.py:322-324, JobMetaData.py:315-370), so
        call it after round_schedule().  Not part of the reference's surface (it plans on one type only).
        Returns (job ids, dict of market.solve_relaxation)."""
        from . import market
        if self.last_forecast is None:
            raise RuntimeError("heterogeneous_plan() needs the forecast of a re-solve: call round_schedule() first")
        jobids = list(self.metadata.keys())
        if len(jobids) != len(self.last_forecast["dbar"]):
            raise RuntimeError("jobs were added or removed since the latest re-solve: call round_schedule() first")
        jobs = [self.metadata[j] for j in jobids]
        g = np.array([j.nworkers for j in jobs], dtype=np.int32)
        E = np.array([j.epochs for j in jobs], dtype=np.float64)
        c = np.array([j.epoch_progress for j in jobs], dtype=np.float64)
        prm = _eng.make_params(self.ngpus, self.future_nrounds, self.round_duration, self.k, self.lam,
                               self.rhomax, self.logapx_bases, self.logapx_origin, self.round_ptr)
        plan = market.solve_relaxation(self._eng(), prm, g, E, c, self.last_forecast["dbar"], self.last_forecast["rem"],
                                       type_speed, capacity, full_iters=full_iters, coarse_iters=coarse_iters)
        return jobids, plan

    # ---- reference surface ------------------------------------------------------------------------
    def round_schedule(self):
        if not self.resolve:
            if len(self.schedules) > 0:
                if self.round_ptr in self.schedules.keys():
                    return self.schedules[self.round_ptr]
        if self._ids_cache is None:                 # rebuilt after add / remove only
            self._ids_cache = (list(self.metadata.keys()), list(self.metadata.values()))
        jobids, jobobjs = self._ids_cache
        # the reference re-seeds both global generators on every solver call (call_cvxpy_solver, shockwave.py:451-452):
        # kept, so that user code drawing from them sees the same state after a re-solve.  Not reproduced: the
        # reference then consumes one `random.choice` draw per back-fill sort-key evaluation (JobMetaData.py:365 with
        # noise_level = 0 — the VALUE is unaffected).  The simulator itself never reads the global generators (it
        # draws from private random.Random instances, scheduler.py:502-511), so closed-loop results do not depend on it.
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
        slot = self._slots.get(jobid)
        if slot is not None:
            self._prog[slot] = job.epoch_progress
            if self._timeline_check == "dirty":
                self._touch_timeline(slot, jobid, job)

    def schedule_progress_batch(self, jobids, epoch_progress, measured_nsamples, end_round):
        """schedule_progress() for many jobs at once, with the throughput-timeline summaries
        (JobMetaData.py:235-249: measured samples, last measured round) supplied by the caller instead of being read
        from each job's `throughput_measurements` dict: the entry point of the device round loop (simulate.py /
        swb_sim_step), which keeps those sums as running state.  Needs timeline_check="dirty" (the default)."""
        if self._timeline_check != "dirty":
            raise RuntimeError("schedule_progress_batch needs timeline_check='dirty'")
        if len(jobids) == 0:
            return
        slots = np.fromiter(map(self._slots.__getitem__, jobids), dtype=np.int64, count=len(jobids))
        ep = np.asarray(epoch_progress, dtype=np.int64)
        for jid, e in zip(jobids, ep.tolist()):
            job = self.metadata[jid]
            job.set_epoch_progress(e)
            job.reset_waiting_delay()
        self._prog[slots] = ep
        self._tl_ns[slots] = np.asarray(measured_nsamples, dtype=np.float64)
        self._tl_end[slots] = np.asarray(end_round, dtype=self._tl_end.dtype)
        self._tl_len[slots] = 0                      # summaries are caller-owned from now on: never re-derived from the dict

    def deschedule_waiting_delay(self, jobid, delay):
        if jobid in self.metadata.keys():
            self.metadata[jobid].add_waiting_delay(delay)

    def add_metadata(self, jobid, jobobj, share_update=True):
        assert jobid not in self.metadata.keys()
        self.metadata[jobid] = jobobj
        self._ids_cache = None
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
        self._ids_cache = None
        self._on_remove(jobid)
        self.set_resolve()


class LazySchedules(OrderedDict):
    """OrderedDict{round -> [job ids]} whose lists are built on first access.

    The reference materialises all T lists in construct_schedules (shockwave.py:233-283), but only
    `schedules[round_ptr]` is ever read (shockwave.py:127,166) and a window is usually replaced by the next
    re-solve after one or a few rounds; at 4096 jobs x 64 rounds building every list costs more host time
    than the whole GPU solve."""

    def __init__(self, x, backfill, bfkey, jobids, round_ptr, xmask=None, bfmask=None, T=None):
        super().__init__()
        self._x, self._bf, self._ids, self._r0 = x, backfill, jobids, round_ptr
        self._xm, self._bm = xmask, bfmask      # packed form: [J, 2] uint64, bit t of the 128-bit row = round t
        self._bfkey, self._order = bfkey, None
        for t in range(x.shape[1] if x is not None else T):
            super().__setitem__(round_ptr + t, None)

    def _column(self, which, t):
        if self._xm is not None:
            m = self._xm if which == 0 else self._bm
            return np.flatnonzero((m[:, t >> 6] >> np.uint64(t & 63)) & np.uint64(1))
        return np.flatnonzero((self._x if which == 0 else self._bf)[:, t])

    def __getitem__(self, rnd):
        cur = super().__getitem__(rnd)
        if cur is None:
            t = rnd - self._r0
            ids = self._ids
            cur = [ids[j] for j in self._column(0, t).tolist()]
            bfj = self._column(1, t)
            if len(bfj):
                if self._order is None:      # descending remaining runtime, stable (shockwave.py:261-267)
                    self._order = np.argsort(-np.asarray(self._bfkey, dtype=np.float64), kind="stable")
                    self._rank = np.empty(len(self._order), dtype=np.int64)
                    self._rank[self._order] = np.arange(len(self._order))
                cur += [ids[j] for j in bfj[np.argsort(self._rank[bfj], kind="stable")].tolist()]
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
