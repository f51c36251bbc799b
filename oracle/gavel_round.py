"""CPU ORACLE (test infrastructure, not product code) for Gavel's per-round priority -> selection -> worker-assignment
step, the part of the reference's round mechanism that runs right after `policy.get_allocation()`:

  priorities    scheduler/scheduler.py:3669-3724  (`_update_priorities`, the "Compute priorities" half)
  selection     scheduler/scheduler.py:1146-1265  (`_schedule_jobs_on_workers_helper`, the vanilla-Gavel branch,
                                                   single jobs — no packed pairs)
  assignment    scheduler/scheduler.py:1306-1378  (`_schedule_jobs_on_workers`, largest scale factor first, lease
                                                   extension, then strided assignment) with
                scheduler/scheduler.py:1049-1110  (`_assign_workers_to_job`)

restated on plain arrays (job index = position in the reference's dict iteration order, worker type index = position in
the caller's `worker_types` list) with the reference's exact float64 operations and tie-breaking, so results are
bit-for-bit those of the dict code.  PINNED: `tests/test_oracle_gavel_round.py` swaps these functions into the
unmodified reference simulator (oracle/ref_harness.py + shockwave_b200.placement.GavelRoundMixin with the oracle
backend) and demands the golden pickles' `per_round_schedule` round by round.  Only tests/, smoke() and bench.py's CPU
legs may import this module.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np


def update_priorities(alloc, job_time, worker_time, thr_ok):
    """scheduler.py:3669-3724.  alloc, job_time [J, W] float64; worker_time [W]; thr_ok [J, W] bool (throughput != 0).
    A job absent from the allocation has alloc = NaN (priority 0, :3700-3701)."""
    J, W = alloc.shape
    prio = np.zeros((J, W), dtype=np.float64)
    for w in range(W):
        wt = worker_time[w]
        for j in range(J):
            a = alloc[j, w]
            if np.isnan(a):
                prio[j, w] = 0.0
                continue
            fraction = 0.0 if wt == 0.0 else job_time[j, w] / wt      # :3682-3696
            new_priority = a * 1e9                                     # :3703
            if a == 0.0:
                new_priority = 0.0
            elif not thr_ok[j, w]:
                new_priority = 0.0                                     # :3706-3717
            elif fraction > 0.0:
                new_priority = a / fraction                            # :3718-3722
            prio[j, w] = new_priority
    return prio


def select_jobs(prio, deficit, alloc, thr_pos, sf, capacity, type_order, isolated_plus=False, fifo=False):
    """scheduler.py:1166-1258 with _enable_global_queue False.  thr_pos [J, W] bool (throughput > 0).
    Returns {w: [job indices in selection order]}."""
    J, W = prio.shape
    already = np.zeros(J, dtype=bool)
    left = {w: int(capacity[w]) for w in type_order}
    out = {w: [] for w in type_order}
    queue = []
    for w in type_order:
        entries = [(j, w, prio[j, w], deficit[j, w], 0.0 if np.isnan(alloc[j, w]) else alloc[j, w]) for j in range(J)]
        queue += sorted(entries, key=lambda x: (x[2], x[3], x[4]), reverse=True)      # stable, :1206-1210
    for j, w, *_ in queue:
        if left[w] == 0:
            continue
        if already[j]:
            continue
        if not thr_pos[j, w]:
            continue
        if fifo and prio[j, w] <= 0.0:
            continue
        if sf[j] > left[w]:
            if not isolated_plus:
                continue
            break          # NB: the reference's `break` leaves the WHOLE concatenated queue (:1243-1250)
        left[w] -= int(sf[j])
        already[j] = True
        out[w].append(j)
    return out


def assign_workers(selected, sf, in_alloc, servers, prev, type_order):
    """scheduler.py:1306-1378 + :1049-1110.  selected {w: [job idx]} (selection order); servers {w: [[worker ids] per
    server]}; prev {job idx: (w, (worker ids))} = the previous round's assignment.  Returns OrderedDict{job idx: tuple}."""
    new = OrderedDict()
    for w in type_order:
        jobs = sorted(selected[w], key=lambda j: sf[j], reverse=True)      # stable, :1311
        worker_ids = [list(s) for s in servers[w]]
        assigned = set()
        ptr = 0
        for cur in sorted(set(int(sf[j]) for j in jobs), reverse=True):
            for j in jobs:                                                   # lease extension, :1335-1356
                if sf[j] != cur:
                    continue
                if j in prev and prev[j][0] == w:
                    pw = prev[j][1]
                    if all(p not in assigned for p in pw):
                        new[j] = tuple(pw)
                        assigned.update(pw)
            for j in jobs:                                                   # remaining jobs, :1359-1378
                if sf[j] != cur:
                    continue
                if not in_alloc[j]:
                    continue
                mine = list(new[j]) if j in new else []
                while len(mine) < sf[j] and ptr < len(worker_ids):           # :1084-1095
                    if len(worker_ids[ptr]) == 0:
                        ptr += 1
                        continue
                    cand = worker_ids[ptr][0]
                    if cand not in assigned:
                        mine.append(cand)
                        assigned.add(cand)
                    worker_ids[ptr].pop(0)
                if len(mine) != sf[j]:
                    raise RuntimeError("Could not assign workers to job %s!" % (j,))
                new[j] = tuple(mine)
    return new


def gavel_round(alloc, job_time, worker_time, thr, deficit, sf, capacity, type_order, servers, prev, in_alloc=None,
                isolated_plus=False, fifo=False):
    """The three steps back to back on plain arrays (what shockwave_b200's swb_gavel_round computes on the device)."""
    alloc = np.asarray(alloc, dtype=np.float64)
    thr = np.asarray(thr, dtype=np.float64)
    if in_alloc is None:
        in_alloc = ~np.isnan(alloc).all(axis=1)
    prio = update_priorities(alloc, np.asarray(job_time, dtype=np.float64), np.asarray(worker_time, dtype=np.float64),
                             thr != 0)
    sel = select_jobs(prio, np.asarray(deficit, dtype=np.float64), alloc, thr > 0, sf, capacity, type_order,
                      isolated_plus, fifo)
    asg = assign_workers(sel, sf, in_alloc, servers, prev, type_order)
    return prio, sel, asg
