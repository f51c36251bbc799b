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
`_scale_bs_and_iters`) on more than one worker type, job pairs (packing: the reference's loop raises on the first scheduled pair,
scheduler.py:1408, tests/test_closed_loop_packed_host.py), the `ideal` and generated-arrival modes
(the generated-arrival mode cannot run in the reference itself: `Scheduler.__init__` opens its trace pickle
unconditionally, scheduler.py:437).
One quirk is kept on purpose: the over-deadline rule sums `_cumulative_run_time` per worker id and divides by the scale
factor (scheduler.py:4378-4390); here it is the running sum of the per-round execution times — the same real number,
possibly another rounding when a gang moves between workers (matters only at an exact tie with int(1.5 duration)).

Only tests/ and bench.py's CPU legs may import this module."""
import math

import numpy as np

PREEMPTION_OVERHEAD = 20.0       # scheduler.py:1947-1949
MAX_FAILED_ATTEMPTS = 5          # scheduler.py:66


def steps_and_finish(now, thr, tpi, remaining):
    """scheduler.py:1441-1465 + :1500-1508 for a single static job: (num_steps, finish_time)."""
    n = min(int(thr * tpi), remaining)               # negative after a rescale that rounds the progress up past the total
    return n, max(now, now + n / thr)                # max_finish_time starts at the current timestamp (:1470, :1507)


def run(trace, select, tpi=120.0, on_round=None, max_rounds=None, dyn=None, throughput_w=None):
    """throughput_w: [J][W] per-worker-type throughputs (`Scheduler._throughputs[job][worker_type]`, static jobs only);
    select() then returns {job: worker type index} and the round's steps / finish time use the throughput of that type
    (`_get_job_steps_and_finish_times(job_id, worker_type)`, scheduler.py:1467-1512).
    trace: dict of equal-length sequences `arrival, total_steps, scale_factor, throughput, duration, batch_size,
    dataset_len`.  select(c, now, active) -> iterable of job indices to run in round c (active = ascending list of live
    jobs).  on_round(c, now, info) sees what the shockwave hook would see (epoch progress, measured throughput).
    dyn: tables of the dynamic-adaptation jobs (accordion / gns batch-size rescaling, scheduler.py:1604-1727 and
    :4731-4935) as shockwave_b200.simulate.build_dynamic_tables lays them out; None = every job static.
    Returns dict(makespan, rounds, jct[J] (nan = never completed), timeline {job: [(round, throughput, bs)]},
    per_round_schedule)."""
    arrival = [float(a) for a in trace["arrival"]]
    J = len(arrival)
    total = [int(v) for v in trace["total_steps"]]
    thr = [float(v) for v in trace["throughput"]]
    sf = [int(v) for v in trace["scale_factor"]]
    dur = [float(v) for v in trace["duration"]]
    bs = [int(v) for v in trace["batch_size"]]
    ds = [int(v) for v in trace["dataset_len"]]
    spe = [math.ceil(d / b) for d, b in zip(ds, bs)]                          # steps per epoch, scheduler.py:2332-2335
    orig_bs = list(bs)
    fails = [0] * J                                       # _num_failures_per_job
    flag = [0] * J                                        # 1 = big_bs, 2 = small_bs (scheduler.py:_bs_flags)
    errs = set()
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

    def request(j):
        """_simulate_accordion :1658-1727 / _simulate_gns :1604-1656: raise a rescale request from the job's progress."""
        cur = -(-steps_run[j] // spe[j])                  # _get_num_epochs: ceil(steps / ceil(dataset / bs))
        pat = dyn["pattern"][j]
        if dyn["mode"][j] == 1:
            if dyn["acc_skip"][j]:
                return
            if cur >= len(pat):
                errs.add("pattern")
                return
            crit = bool(pat[cur])
            if bs[j] == orig_bs[j] and not crit:
                if bs[j] != dyn["bs_max"][j]:
                    flag[j] = 1
            elif bs[j] != orig_bs[j] and crit:
                if bs[j] != dyn["bs_min"][j]:
                    flag[j] = 2
        else:
            if cur + 1 >= len(pat):
                errs.add("pattern")
                return
            # bs_gns = get_gns_bs_pattern(.., max(760, cur + 2), ..): its last entry is never scaled (utils.py:801-1010)
            nxt = orig_bs[j] if cur + 1 >= 759 else pat[cur + 1]
            if nxt > bs[j] or pat[cur] > bs[j]:
                if bs[j] != dyn["bs_max"][j]:
                    flag[j] = 1

    def rescale(j):
        """_scale_bs_and_iters :4731-4935 for one job whose request flag is up."""
        if dyn["orig_locked"][j]:
            return
        old = bs[j]
        if dyn["mode"][j] == 2:
            new = 2 * old
        elif flag[j] == 1:
            new = dyn["bs_big"][j]
        else:
            new = orig_bs[j]
        lv = dyn["lvl_bs"][j]
        if new not in lv or not (dyn["lvl_thr"][j][lv.index(new)] > 0):
            return                                        # :4803-4818: batch size not in the throughput file
        factor = new / old
        it = 1 / factor
        bs[j] = new
        thr[j] = dyn["lvl_thr"][j][lv.index(new)]
        spe_old, spe_new = math.ceil(ds[j] / old), math.ceil(ds[j] / new)
        old_epochs = math.ceil(total[j] / spe_old)
        new_total = math.ceil(total[j] * it)
        if math.ceil(new_total / spe_new) != old_epochs:
            new_total = spe_new * old_epochs
        done_epochs = math.ceil(steps_run[j] / spe_old)
        total[j] = new_total
        steps_run[j] = done_epochs * spe_new
        spe[j] = spe_new

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
            if n <= 0 and ex <= 0:                        # micro-task failure, scheduler.py:4497-4570: no progress booked,
                fails[j] += 1                             # the job is dropped after MAX_FAILED_ATTEMPTS in a row
                done = fails[j] >= MAX_FAILED_ATTEMPTS
            else:
                fails[j] = 0
                steps_run[j] += n
                done = total[j] - steps_run[j] <= 0 or over
            timeline[j].append((c, 0.0 if ex <= 0 else n / ex, bs[j]))
            if dyn is not None and flag[j]:
                rescale(j)
            flag[j] = 0                                   # scheduler.py:4700-4712: reset in every _done_callback
            if done:
                status[j] = 2
                jct[j] = latest[j] - arrival[j]
                remaining_jobs -= 1
            ran.append(j)
        running = []
        if dyn is not None:                               # scheduler.py:2019-2027: every live job, every iteration
            for j in range(J):
                if status[j] == 1 and dyn["mode"][j]:
                    request(j)
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
        picked = select(c, now, active)
        chosen = [j for j in picked if status[j] == 1]
        schedule.append(set(chosen))
        for j in chosen:
            th = thr[j] if throughput_w is None else float(throughput_w[j][picked[j]])
            n, fin = steps_and_finish(now, th, tpi, total[j] - steps_run[j])
            running.append((fin, j, n))
        c += 1
        if max_rounds is not None and c >= max_rounds:
            break
    return dict(makespan=now, rounds=c, jct=jct, timeline=timeline, per_round_schedule=[sorted(s) for s in schedule],
                steps_run=steps_run, errs=sorted(errs), final_bs=bs)


def trace_arrays(rec):
    """The trace part of a sim_static_pins.json record as numpy arrays (the layout swb_sim_* takes)."""
    return dict(arrival=np.asarray(rec["arrival"], np.float64), total_steps=np.asarray(rec["total_steps"], np.int64),
                scale_factor=np.asarray(rec["scale_factor"], np.int32), throughput=np.asarray(rec["throughput"], np.float64),
                duration=np.asarray(rec["duration"], np.float64), batch_size=np.asarray(rec["batch_size"], np.int32),
                dataset_len=np.asarray(rec["dataset_len"], np.int64))
