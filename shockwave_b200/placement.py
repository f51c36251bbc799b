"""Drop-in for Gavel's per-round priority -> selection -> worker-assignment step, backed by gavel.cu.

`GavelRoundMixin` goes IN FRONT of the reference's `Scheduler` class (scheduler/scheduler.py) in the MRO:

    class Scheduler(shockwave_b200.placement.GavelRoundMixin, scheduler.Scheduler): pass

and overrides three of its methods for the vanilla-Gavel, simulation, single-job (no packing) case — everything else,
and every other case, runs the reference's own code through super():

  _update_priorities                 scheduler/scheduler.py:3600-3724   allocation reset logic kept line for line, the
                                                                          per-(job, type) priority loop deferred to ...
  _schedule_jobs_on_workers_helper   scheduler/scheduler.py:1113-1271   ... ONE device call (swb_gavel_round) that
                                                                          computes priorities, the sorted queues, the
                                                                          greedy selection and the worker assignment
  _assign_workers_to_job             scheduler/scheduler.py:1049-1110   hands out the workers that call computed (the
                                                                          lease-extension loop of the caller stays in
                                                                          Python and reaches the same decisions)

The reference keeps its state in dicts that its own code mutates everywhere, so the per-round packing is a walk over
those dicts (O(J W), what `_update_priorities` itself does); results are bit-for-bit the dict code's
(tests/test_gpu_gavel_round.py; closed loop on the golden pickles in tests/test_closed_loop.py).
`_swb_backend` may be replaced by oracle.gavel_round (tests only) to pin the restatement without a GPU."""
from __future__ import annotations

import numpy as np

from . import policies as _pol


class _DeviceBackend:
    def gavel_round(self, alloc, job_time, worker_time, thr, deficit, sf, capacity, type_order, worker_lists, prev,
                    isolated_plus=False, fifo=False):
        return _pol._engine().gavel_round(alloc, job_time, worker_time, thr, deficit, sf, capacity, type_order,
                                          worker_lists, prev, isolated_plus, fifo)


class GavelRoundMixin:
    _swb_backend = _DeviceBackend()
    _swb_round = None
    swb_round_calls = 0

    def _swb_applicable(self):
        return (self._simulate and self._policy.name != "shockwave" and not getattr(self, "_job_packing", False)
                and not self._enable_global_queue)

    def _update_priorities(self):
        if not self._swb_applicable():
            return super()._update_priorities()
        # scheduler.py:3611-3636, unchanged: when to reset the time accounting and recompute the allocation
        current_time = self.get_current_timestamp()
        time_since_last_reset = current_time - self._last_reset_time
        reset_interval_elapsed = time_since_last_reset >= self._minimum_time_between_allocation_resets
        need_to_reset_time_run_so_far = reset_interval_elapsed or self._last_reset_time == 0
        need_to_reset_time_run_so_far = self._need_to_update_allocation and need_to_reset_time_run_so_far
        if need_to_reset_time_run_so_far:
            self._reset_time_run_so_far()
            self._allocation = self._compute_allocation()
            self._need_to_update_allocation = False
        self._swb_round = None      # priorities are computed with the selection, in one device call

    def _schedule_jobs_on_workers_helper(self, worker_types):
        if not self._swb_applicable() or not worker_types:
            return super()._schedule_jobs_on_workers_helper(worker_types)
        all_types = sorted(self._worker_types)            # type index = position in this list
        tix = {wt: i for i, wt in enumerate(all_types)}
        W = len(all_types)
        # job order = iteration order of the reference's priority dicts (identical for every worker type)
        jobs = list(self._priorities[worker_types[0]].keys())
        J = len(jobs)
        if J == 0:
            return {wt: [] for wt in worker_types}
        jix = {job_id: j for j, job_id in enumerate(jobs)}
        alloc = np.full((J, W), np.nan)
        job_time = np.zeros((J, W)); thr = np.zeros((J, W)); deficit = np.zeros((J, W))
        A = self._allocation if self._allocation is not None else {}
        for j, job_id in enumerate(jobs):
            a = A.get(job_id)
            jt = self._job_time_so_far.get(job_id, {})
            th = self._throughputs[job_id]
            for wt in all_types:
                w = tix[wt]
                if a is not None:
                    alloc[j, w] = a[wt]
                job_time[j, w] = jt.get(wt, 0.0)   # a job the accounting does not know has fraction 0 (:3684-3689)
                thr[j, w] = th[wt]
                deficit[j, w] = self._deficits[wt][job_id]
        worker_time = np.array([self._worker_time_so_far[wt] for wt in all_types], dtype=np.float64)
        sf = np.array([self._jobs[job_id].scale_factor for job_id in jobs], dtype=np.int32)
        rest = [wt for wt in all_types if wt not in worker_types]
        order = [tix[wt] for wt in worker_types] + [tix[wt] for wt in rest]
        worker_lists = [[wid for server in self._worker_type_to_worker_id_mapping[wt] for wid in server]
                        for wt in worker_types] + [[] for _ in rest]
        capacity = np.array([self._cluster_spec[wt] if wt in worker_types else 0 for wt in all_types], dtype=np.int32)
        prev = {}
        for job_id, wids in self._current_worker_assignments.items():
            if job_id in jix:
                prev[jix[job_id]] = (tix[self._worker_id_to_worker_type_mapping[wids[0]]], tuple(wids))
        name = self._policy.name
        prio, sel, asg = self._swb_backend.gavel_round(alloc, job_time, worker_time, thr, deficit, sf, capacity, order,
                                                       worker_lists, prev, isolated_plus=(name == "Isolated_plus"),
                                                       fifo=name.startswith("FIFO"))
        GavelRoundMixin.swb_round_calls += 1
        # the reference's dicts get the same values its own loop would have written (other code reads them)
        for wt in all_types:
            w = tix[wt]
            pw = self._priorities[wt]
            for j, job_id in enumerate(jobs):
                pw[job_id] = float(prio[j, w])
        self._swb_round = {jobs[j]: ws for j, ws in asg}
        return {wt: [(jobs[j], int(sf[j])) for j in sel[tix[wt]]] for wt in worker_types}

    def _assign_workers_to_job(self, job_id, scale_factor, worker_type, worker_state, worker_assignments):
        if not self._swb_applicable() or self._swb_round is None or job_id not in self._swb_round:
            return super()._assign_workers_to_job(job_id, scale_factor, worker_type, worker_state, worker_assignments)
        ws = self._swb_round[job_id]
        assert len(ws) == scale_factor
        worker_assignments[job_id] = tuple(ws)
        worker_state["assigned_worker_ids"].update(ws)     # the caller's lease-extension test reads this set
        for single_job_id in job_id.singletons():          # scheduler.py:1103-1110
            self._per_job_latest_timestamps[single_job_id] = self.get_current_timestamp()
            self._running_jobs.add(single_job_id)
