"""TEST INFRASTRUCTURE (oracle): plain-array restatement of the reference simulator's round loop for trace-driven runs of
STATIC jobs on one worker type — `Scheduler.simulate()` scheduler/scheduler.py:1878-2250 with the pieces of
`_get_job_steps_and_finish_times` :1467-1512, `_get_num_steps` :1425-1465, `_done_callback` :4341-4700,
`_update_throughput` :549-571, `_remove_job` :808-830 and `_update_shockwave_scheduler` :2270-2342 that act on such
jobs.  The policy is a callback (`select`), so the loop can replay a schedule the reference recorded.

Pinned: tests/test_oracle_sim_loop.py replays the per-round schedules recorded from the UNMODIFIED reference
(tests/golden/make_sim_pins.py -> sim_static_pins.json; fifo and max_min_fairness on 32 and 12 GPUs, 303-643 rounds) and
demands the same completion time of every job, the same makespan, the same number of rounds and the same measured-
throughput timeline, to the last bit.

Not restated (rejected by the product too): accordion / gns batch-size rescaling (`_simulate_accordion`, `_simulate_gns`,
`_scale_bs_and_iters`), job pairs (packing), several worker types, the `ideal` and generated-arrival modes.
One quirk is kept on purpose: the over-deadline rule sums `_cumulative_run_time` per worker id and divides by the scale
factor (scheduler.py:4378-4390); here it is the running sum of the per-round execution times — the same real number,
possibly another rounding when a gang moves between workers (matters only at an exact tie with int(1.5 duration)).

Only tests/ and bench.py's CPU legs may import this module."""
import math

import numpy as np

PREEMPTION_OVERHEAD = 20.0       # scheduler.py:1947-1949


def steps_and_finish(now, thr, tpi, remaining):
    """scheduler.py:1441-1465 + :1500-1508 for a single static job: (num_steps, finish_time)."""
    n = min(int(thr * tpi), remaining)
    return n, now + n / thr


def run(trace, select, tpi=120.0, on_round=None, max_rounds=None):
    """trace: dict of equal-length sequences `arrival, total_steps, scale_factor, throughput, duration, batch_size,
    dataset_len`.  select(c, now, active) -> iterable of job indices to run in round c (active = ascending list of live
    jobs).  on_round(c, now, info) sees what the shockwave hook would see (epoch progress, measured throughput).
    Returns dict(makespan, rounds, jct[J] (nan = never completed), timeline {job: [(round, throughput, bs)]},
    per_round_schedule)."""
    arrival = [float(a) for a in trace["arrival"]]
    J = len(arrival)
    total = [int(v) for v in trace["total_steps"]]
    thr = [float(v) for v in trace["throughput"]]
    sf = [int(v) for v in trace["scale_factor"]]
    dur = [float(v) for v in trace["duration"]]
    bs = [int(v) for v in trace["batch_size"]]
    spe = [math.ceil(d / b) for d, b in zip(trace["dataset_len"], bs)]       # steps per epoch, scheduler.py:2332-2335
    assert all(arrival[i] <= arrival[i + 1] for i in range(J - 1))            # scheduler.py:1842-1843
    status = [0] * J                                      # 0 queued, 1 live, 2 completed
    steps_run = [0] * J
    run_time = [0.0] * J
    latest = [None] * J
    jct = [math.nan] * J
    timeline = {j: [] for j in range(J)}
    schedule = []
    running = []                                          # (finish_time, job, num_steps)
    queue = 0                                             # index of the first job not yet admitted
    now = arrival[0]                                      # scheduler.py:1847
    round_start, round_end = 0.0, None                    # scheduler.py:1819-1820
    remaining_jobs = J
    c = 0
    while True:
        if remaining_jobs == 0:
            break
        next_arrival = arrival[queue] if queue < J else None
        max_ts = 0
        if running:
            m = max(r[0] for r in running)
            if m > max_ts:
                max_ts = m
                if round_end is not None:
                    round_start = round_end
                round_end = max_ts
        now = max_ts if max_ts > 0 else next_arrival
        ran = []
        for finish, j, n in sorted(running, key=lambda r: (-r[0], r[1])):     # heap order: latest finish first
            ex = finish - round_start
            slow = 1
            if c != 1 and j not in schedule[c - 2]:
                if ex != 0 and tpi - 5 < ex:
                    slow = (ex - PREEMPTION_OVERHEAD) / ex
                    ex -= PREEMPTION_OVERHEAD
            latest[j] = finish
            n = int(n * slow)
            run_time[j] += ex
            over = run_time[j] > int(dur[j] * 1.5)
            steps_run[j] += n
            timeline[j].append((c, 0.0 if ex <= 0 else n / ex, bs[j]))
            if total[j] - steps_run[j] <= 0 or over:
                status[j] = 2
                jct[j] = latest[j] - arrival[j]
                remaining_jobs -= 1
            ran.append(j)
        running = []
        if on_round is not None and now != 0.0:
            prev = schedule[c - 1] if c >= 1 else []
            on_round(c, now, dict(scheduled=list(prev), epoch={j: (None if status[j] == 2 else steps_run[j] // spe[j])
                                                              for j in prev}))
        while queue < J and arrival[queue] <= now:
            status[queue] = 1
            queue += 1
        active = [j for j in range(J) if status[j] == 1]
        if not active:
            break                                         # scheduler.py:2173-2178 (even with jobs still queued)
        chosen = [j for j in select(c, now, active) if status[j] == 1]
        schedule.append(set(chosen))
        for j in chosen:
            n, fin = steps_and_finish(now, thr[j], tpi, total[j] - steps_run[j])
            running.append((fin, j, n))
        c += 1
        if max_rounds is not None and c >= max_rounds:
            break
    return dict(makespan=now, rounds=c, jct=jct, timeline=timeline, per_round_schedule=[sorted(s) for s in schedule],
                steps_run=steps_run)


def trace_arrays(rec):
    """The trace part of a sim_static_pins.json record as numpy arrays (the layout swb_sim_* takes)."""
    return dict(arrival=np.asarray(rec["arrival"], np.float64), total_steps=np.asarray(rec["total_steps"], np.int64),
                scale_factor=np.asarray(rec["scale_factor"], np.int32), throughput=np.asarray(rec["throughput"], np.float64),
                duration=np.asarray(rec["duration"], np.float64), batch_size=np.asarray(rec["batch_size"], np.int32),
                dataset_len=np.asarray(rec["dataset_len"], np.int64))
