#!/usr/bin/env python
"""bench.py — schedule rounds/sec of the per-round market solve on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE.json config D — 4096 jobs x 512 GPUs x 64-round planning window.
One "step" = the Monte-Carlo (GBM) throughput forecast of the live job set (8192 paths/job, sharded over
the ranks, one NCCL allreduce) followed by one pass of the market solve + round placement +
work-conserving back-fill over a batch of S independent scenarios of that size (a hyper-parameter / trace-ensemble sweep: every
scenario has its own synthetic job set and its own k).  A scenario is one CTA, so S = 2 x 148 fills
the GPU twice.  `value` = scenarios solved per second with the inputs resident in HBM; `e2e` = the
same through the public host API (Engine.solve, pinned host buffers, H2D + D2H inside the call);
`latency_ms_S1` in `config` = one single re-solve (what ShockwaveScheduler.round_schedule() pays).

N > 1 (torchrun): scenarios shard across ranks with no data-path collective ("weak" scaling, S per
GPU fixed); the only collective is the max-over-ranks of the timed region.

--impl reference: the CPU path on the host cores — the HiGHS restatement of the reference's MILP
(oracle/shockwave_milp.py; Gurobi/cvxpy cannot be installed here), one scenario of the same config
per step with the reference's own solver settings (MIPGap 1e-3, TimeLimit 15 s).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BASES = [0.0, 0.2, 0.4, 0.6, 0.8, 1.0]
ORIGIN = {0.0: 1e-6}
J, G, T, D = 4096, 512, 64, 120.0
K_SWEEP = [1e-3, 1e1, 1e5]          # the k values of the shipped configurations/*.json
METRIC = "schedule rounds/sec (4096 jobs x 512 GPUs x 64-round window)"


def synth_batch(S, seed0):
    from tests.synth import synth_problem
    pbs = [synth_problem(J, G, T, D, seed=seed0 + s, tight=3.0 if s % 4 else 0.5) for s in range(S)]
    st = lambda k, dt: np.ascontiguousarray(np.stack([p[k] for p in pbs]).astype(dt))
    arrs = dict(g=st("g", np.int32), E=st("E", np.int32), c=st("c", np.int32), dbar=st("dbar", np.float64),
                rem=st("rem", np.float64), ftobj=st("ftobj", np.float64))
    return arrs, [p["round_ptr"] for p in pbs]


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons DURING the timed region, through NVML (nvidia_ml_py) every 5 ms — the timed
    region is ~50 ms, far too short for spawning nvidia-smi (B200_PROFILING.md's query, same fields)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.mx, self.power = index, False, [], set(), None, []
        self.t0, self.t1, self.ts, self.bits = None, None, [], []

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            while not self.stop_flag:
                self.ts.append(time.perf_counter())
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                    self.bits.append(int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
                except Exception:
                    self.power.append(0.0); self.bits.append(0)
                time.sleep(0.002)
        except Exception as e:                      # NVML missing: fall back to one nvidia-smi query
            self.error = repr(e)
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                a, b = [float(v) for v in out.stdout.strip().split(",")]
                self.sm.append(a); self.mx = b
            except Exception:
                pass

    def summary(self):
        """Samples taken inside [t0, t1] (the timed region); the sampler itself starts before the warm-up."""
        idx = [i for i, t in enumerate(self.ts) if self.t0 is not None and self.t0 <= t <= (self.t1 or 1e300)]
        if not idx:
            idx = list(range(len(self.sm)))
        sm = [self.sm[i] for i in idx]
        reasons = set(self.reasons)
        for i in idx:
            if i < len(self.bits):
                for b, name in self.REASONS.items():
                    if self.bits[i] & b:
                        reasons.add(name)
        pw = [self.power[i] for i in idx if i < len(self.power)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


def _ref_one(seed):
    """One config-D scenario through the HiGHS restatement of the reference's first MILP."""
    from oracle import shockwave_milp as om
    from tests.synth import synth_problem
    logv = om.pwl_log_values(BASES, ORIGIN)
    pb = synth_problem(J, G, T, D, seed=seed, tight=3.0)
    t0 = time.perf_counter()
    cap = om.ftf_caps(pb["rem"], pb["ftobj"], G, J, T, D, pb["round_ptr"], 1.0)
    ok, x, p, obj = om._solve(pb["g"].astype(np.int64), pb["E"].astype(float), pb["c"].astype(float),
                              pb["dbar"], pb["rem"], np.ones(J), G, T, D, 1e-3, BASES, logv, cap, 1e-3, 15.0)
    return bool(ok), time.perf_counter() - t0


def run_reference(args):
    """CPU arm: HiGHS restatement of the reference MILP on ALL host cores — HiGHS' branch-and-bound is
    single-threaded, so one independent config-D scenario per core and step (a throughput-fair figure)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    # one single-threaded HiGHS process per host core, at most 16: with 64 concurrent processes on the 128-core GPU box
    # one step took 416 s instead of ~70 s (memory-bandwidth contention) for barely more throughput (0.154 vs 0.10
    # rounds/s) — past 16 the run no longer finishes "within a few minutes"
    cores = max(1, min(os.cpu_count() or 1, 16))
    steps = max(1, min(args.steps, 2))          # one step is ~70-100 s of CPU at this size
    solved, walls = [], []
    with mp.get_context("spawn").Pool(cores) as pool:
        for i in range(steps):
            t0 = time.perf_counter()
            out = pool.map(_ref_one, [1000 + i * cores + w for w in range(cores)])
            walls.append(time.perf_counter() - t0)
            solved += [o[0] for o in out]
    total = float(sum(walls))
    val = steps * cores / total
    sample = (f"{steps} step(s) x {cores} concurrent scenarios of the config-D workload (one per host core), first MILP "
              f"of dynamic_eisenberg_gale_scheduling only (model build + HiGHS, mip_rel_gap=1e-3, time_limit=15 s = "
              f"the reference's Gurobi settings); incumbents found: {sum(solved)}/{len(solved)}")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "rounds/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": 0, "ms_per_step": 1e3 * total / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"4096 jobs x 512 GPUs x 64-round window (BASELINE config D), {cores} scenarios/step",
                       "solver": "HiGHS via scipy.optimize.milp — stand-in for Gurobi (not installable here)"},
            "cpu_baseline": {"value": val, "unit": "rounds/s", "cores": cores, "kind": "port", "sample": sample,
                             "host_cores": os.cpu_count()},
            "e2e": {"value": val, "unit": "rounds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenarios", type=int, default=296, help="scenarios per GPU per step (2 x 148 SMs)")
    ap.add_argument("--mc-paths", type=int, default=8192, help="GBM sample paths per job (global, sharded over ranks)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from shockwave_b200 import Engine, make_params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    eng = Engine(local)
    S = args.scenarios
    W = max(args.warmup, 3)
    K = args.steps

    arrs, rptr = synth_batch(S, seed0=rank * S)
    prms = [make_params(G, T, D, K_SWEEP[s % len(K_SWEEP)], 12.0, 1.0, BASES, ORIGIN, round_ptr=rptr[s])
            for s in range(S)]
    # ---- device-resident inputs / outputs (value) ----
    dten = {k: torch.from_numpy(v).to(dev) for k, v in arrs.items()}
    outs = dict(x=torch.empty((S, J, T), dtype=torch.uint8, device=dev),
                backfill=torch.empty((S, J, T), dtype=torch.uint8, device=dev),
                nrounds=torch.empty((S, J), dtype=torch.int32, device=dev),
                weights=torch.empty((S, J), dtype=torch.float64, device=dev))
    ptrs = {k: v.data_ptr() for k, v in dten.items()}
    ptrs["bfkey"] = dten["rem"].data_ptr()
    optrs = {k: v.data_ptr() for k, v in outs.items()}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    # Engine calls are synchronous (they end with a stream synchronize), so events recorded on torch's
    # current stream bracket everything a step launches: our kernels on the engine stream and NCCL.
    est = torch.cuda.current_stream(dev)

    # ---- Monte-Carlo (GBM) forecast of the live job set: P paths/job sharded over the ranks, ONE allreduce ----
    from shockwave_b200.forecast_mc import path_range
    P_MC = args.mc_paths
    rng = np.random.default_rng(12345)
    mc_R0 = arrs["rem"][0].copy()
    mc_H = np.minimum(arrs["E"][0] - arrs["c"][0], 256).astype(np.int32)
    mc_mu = rng.uniform(-1e-3, 1e-3, J)
    mc_sg = rng.uniform(0.0, 0.05, J)
    mc_lo, mc_n = path_range(P_MC, rank, world)
    mc_out = torch.zeros((2, J), dtype=torch.float64, device=dev)

    def step_resident():
        eng.gbm_forecast(mc_R0, mc_H, mc_mu, mc_sg, mc_n, mc_lo, 7, out_device_ptr=mc_out.data_ptr())
        if world > 1:
            dist.all_reduce(mc_out, op=dist.ReduceOp.SUM)      # NCCL over NVLink: 2*J float64 = 64 KiB
        return eng.solve_device(prms, J, ptrs, optrs)

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(W):
        step_resident()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.t0 = time.perf_counter()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    ksolve, kplace, launches = [], [], 0
    torch.cuda.synchronize()
    for i in range(K):
        flush.fill_(i)                          # L2 flush between timed iterations (not timed)
        torch.cuda.synchronize()
        ev[i][0].record(est)
        res = step_resident()
        ev[i][1].record(est)
        tm = eng.last_timings()
        ksolve.append(tm["ms_solve"]); kplace.append(tm["ms_place"])
        launches += 1 + 2 * tm["passes"] + (tm["passes"] - 1)       # gbm + (solve, place) per pass + tighten
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.t1 = time.perf_counter()
    sampler.stop_flag = True
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    value = world * S * K / (total_ms * 1e-3)
    nfallback = sum(1 for r in res if r["status"] == 1)

    # ---- single re-solve latency (S = 1, resident) ----
    one = [prms[1]]
    p1 = {k: v[1:2].contiguous().data_ptr() for k, v in dten.items()}
    p1["bfkey"] = p1["rem"]
    lat = []
    for i in range(W + 5):
        t0 = time.perf_counter()
        eng.solve_device(one, J, p1, optrs)
        lat.append((time.perf_counter() - t0) * 1e3)
    lat_ms = float(np.median(lat[W:]))
    lat_k = eng.last_timings()

    # ---- e2e through the public host API, pinned host buffers, copies inside the timed region ----
    hin = {k: torch.from_numpy(v).pin_memory() for k, v in arrs.items()}
    hnp = {k: v.numpy() for k, v in hin.items()}
    hout = {"xmask": torch.empty((S, J, 2), dtype=torch.uint64).pin_memory().numpy(),
            "bfmask": torch.empty((S, J, 2), dtype=torch.uint64).pin_memory().numpy(),
            "nrounds": torch.empty((S, J), dtype=torch.int32).pin_memory().numpy(),
            "weights": torch.empty((S, J), dtype=torch.float64).pin_memory().numpy()}
    # two engines (two CUDA streams, two sets of device + pinned buffers) driven by two host threads: step i+1's
    # host->device copies overlap step i's kernels and device->host copies.  Every copy of every step stays inside
    # the timed region; ctypes releases the GIL during the library calls.
    from concurrent.futures import ThreadPoolExecutor
    eng_b = Engine(local)
    hout_b = {k: torch.from_numpy(np.empty_like(v)).pin_memory().numpy() for k, v in hout.items()}
    lanes = [(eng, hout, ThreadPoolExecutor(1)), (eng_b, hout_b, ThreadPoolExecutor(1))]

    def e2e_step(i):
        e, ho, _ = lanes[i % 2]
        sums_ = e.gbm_forecast(mc_R0, mc_H, mc_mu, mc_sg, mc_n, mc_lo, 7)
        e.solve(prms, hnp["g"], hnp["E"], hnp["c"], hnp["dbar"], hnp["rem"], hnp["ftobj"], packed=True, out=ho)
        return sums_

    for f in [lanes[i % 2][2].submit(e2e_step, i) for i in range(6)]:
        f.result()
    # the same K steps one after the other on one engine, for reference (reported in config)
    t0 = time.perf_counter()
    for i in range(K):
        e2e_step(0)
    e2e_serial_s = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    futs = [lanes[i % 2][2].submit(e2e_step, i) for i in range(K)]
    for f in futs:
        sums = f.result()
        if world > 1:       # collectives stay on the main thread, in step order on every rank
            tsum = torch.from_numpy(sums).to(dev)
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            sums = tsum.cpu().numpy()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * S * K / float(te.item())
    for _, _, ex_ in lanes:
        ex_.shutdown()
    del eng_b
    h2d = S * J * (3 * 4 + 4 * 8) + S * 256 + J * (3 * 8 + 4)
    d2h = S * J * 16 * 2 + S * J * (4 + 8) + S * 56 + 2 * J * 8      # round masks (128 bit/job) x2, counts, weights
    # MC kernel alone (device time)
    eng.gbm_forecast(mc_R0, mc_H, mc_mu, mc_sg, mc_n, mc_lo, 7, out_device_ptr=mc_out.data_ptr())
    mc_ms = eng.last_timings()["ms_solve"]

    # ---- the drop-in class itself: ShockwaveScheduler.round_schedule() with a forced re-solve (S = 1) ----
    from collections import OrderedDict
    from shockwave_b200 import ShockwaveScheduler

    class _Job:                      # what scheduler.py hands over (scheduler/JobMetaData.py:41-98), duck-typed
        def __init__(self, jid, r):
            E = int(arrs["E"][0][jid])
            self.jobid, self.nworkers, self.epochs, self.epoch_nsamples = jid, int(arrs["g"][0][jid]), E, 50000
            dur = max(1.0, round(float(arrs["dbar"][0][jid])))
            self.epoch_duration_preprofiled = [dur] * E
            self.bs_schedule = [32] * (E // 2) + [64] * (E - E // 2)
            self.timestamp_submit, self.gavel_round_duration = 0.0, D
            self.throughput_measurements = OrderedDict()
            self.epoch_progress, self.waiting_delay = int(arrs["c"][0][jid]), 0
        def set_epoch_progress(self, c): self.epoch_progress = c
        def reset_waiting_delay(self): self.waiting_delay = 0
        def add_waiting_delay(self, d): self.waiting_delay += d

    sw = ShockwaveScheduler(ngpus=G, gram=16, init_metadata=OrderedDict(), future_nrounds=T, round_duration=int(D),
                            solver_preference=["GUROBI"], solver_rel_gap=1e-3, solver_num_threads=24, solver_timeout=15,
                            n_epoch_vars_max=64, logapx_bases=BASES, logapx_origin=ORIGIN, k=1e-3, lam=12.0, rhomax=1.0,
                            device=local)
    for jid in range(J):
        sw.add_metadata(jid, _Job(jid, None))
    rs_ms = []
    for i in range(W + 5):
        sw.round_ptr = 10 + i
        sw.set_resolve()
        t0 = time.perf_counter()
        sw.round_schedule()
        rs_ms.append((time.perf_counter() - t0) * 1e3)
    rs_ms = float(np.median(rs_ms[W:]))
    del sw

    # ---- dense PR-dynamics pass over X[S][J][W][T]: S*J*W*T*4 B = 512 MiB > L2 ----
    from shockwave_b200.engine import market_pgd
    Sd, Wd = 512, 1
    Xd = torch.zeros((Sd, J, Wd, T), dtype=torch.float32, device=dev)
    dj = {k: torch.from_numpy(np.ascontiguousarray(arrs[k][0]).astype(np.float64 if k != "g" else np.int32)).to(dev)
          for k in ("g", "E", "c", "dbar", "rem")}
    drate = torch.from_numpy((D / arrs["dbar"][0])[:, None].astype(np.float32)).to(dev)
    mptr = dict(shape=(Sd, J, Wd, T), per_scenario_jobs=0, g=dj["g"].data_ptr(), E=dj["E"].data_ptr(),
                c=dj["c"].data_ptr(), dbar=dj["dbar"].data_ptr(), rem=dj["rem"].data_ptr(), rate=drate.data_ptr(),
                X=Xd.data_ptr())
    mprm = [make_params(G, T, D, 1e-9, 12.0, 1.0, BASES, ORIGIN) for _ in range(Sd)]
    dense_ms = []
    for i in range(W + K):
        _, ms = market_pgd(eng, mprm, None, None, None, None, None, None, [G], None, 8, 0.1, 0.3, float(J * T),
                           device_ptrs=mptr, eta_decay=50.0)
        if i >= W:
            dense_ms.append(ms)
    dense_ms = float(np.mean(dense_ms))
    dense_bytes = 8.0 * Sd * J * Wd * T + Sd * (24.0 * J + 8.0 * Wd * T)
    del Xd

    # ---- config E (J=2048 policy sweep): Gavel get_allocation() latency, W=1 (pooled) and W=3 (heterogeneous) ----
    pol_ms = {}
    if rank == 0:
        from shockwave_b200 import policies as GP
        GP._shared_engine = eng
        rngp = np.random.default_rng(7)
        JE = 2048
        mE = rngp.uniform(0.5, 20.0, size=(JE, 1)) * np.sort(rngp.uniform(0.1, 1.0, size=(JE, 3)), axis=1)
        wts = ["k80", "p100", "v100"]
        thr3 = {j: {w: float(mE[j, i]) for i, w in enumerate(wts)} for j in range(JE)}
        thr1 = {j: {w: float(mE[j, 2]) for w in wts} for j in range(JE)}
        sfE = {j: int(rngp.choice([1, 2, 4, 8], p=[0.6, 0.3, 0.09, 0.01])) for j in range(JE)}
        prE = {j: 1.0 for j in range(JE)}
        stE = {j: float(rngp.uniform(1e4, 1e6)) for j in range(JE)}
        tE = {j: float(rngp.uniform(0, 5e3)) for j in range(JE)}
        for tag, thr_, spec_ in (("W1", thr1, {"k80": 0, "p100": 0, "v100": 512}),
                                 ("W3", thr3, {"k80": 256, "p100": 128, "v100": 128})):
            for name, call in (
                    ("max_min_fairness_perf", lambda: GP.MaxMinFairnessPolicyWithPerf("ECOS").get_allocation(thr_, sfE, prE, spec_)),
                    ("finish_time_fairness_perf", lambda: GP.FinishTimeFairnessPolicyWithPerf("GUROBI").get_allocation(
                        thr_, sfE, prE, tE, stE, spec_))):
                call()
                t0 = time.perf_counter()
                for _ in range(3):
                    call()
                pol_ms[f"{name}_{tag}"] = (time.perf_counter() - t0) / 3 * 1e3

    if rank != 0:
        return
    clocks = sampler.summary()
    # ---- roofline of the dominant kernel (solve_kernel): algorithmic bytes per launch ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ms_solve, ms_place = float(np.mean(ksolve)), float(np.mean(kplace))
    # algorithmic bytes per launch: inputs + outputs each kernel must touch once
    kern = {
        "solve_kernel": (ms_solve, S * J * (3 * 4 + 3 * 8) + S * J * (1 + 8)),
        "place_kernel": (ms_place, S * J * (1 + 8 + 8) + S * J * T * 2 + S * J * 4),
        "gbm_kernel": (mc_ms, J * (3 * 8 + 4) + 2 * J * 8),
    }
    dom = max(kern, key=lambda k: kern[k][0])
    dom_ms, dom_bytes = kern[dom]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    bound_note = {
        "solve_kernel": "price/makespan search on per-job scalars staged in shared memory: instruction-issue bound "
                        "(79 % issue-active in profiles/), not HBM bound",
        "place_kernel": "sorts + water-filling over 64 rounds in shared memory: latency bound, not HBM bound",
        "gbm_kernel": "Monte-Carlo paths (xorshift128+, Box-Muller, exp): ALU/SFU bound (86 % issue-active in profiles/), "
                      "28 bytes in / 16 bytes out per job — the HBM fraction is reported only because the contract asks for it",
    }[dom]
    line = {
        "metric": METRIC, "value": value, "unit": "rounds/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE config D: 4096 jobs x 512 GPUs x 64-round window",
                   "scenarios_per_gpu_per_step": S, "k_sweep": K_SWEEP, "fallback_scenarios": nfallback,
                   "latency_ms_S1": lat_ms, "latency_kernels_ms_S1": lat_k,
                   "gavel_get_allocation_ms_J2048": pol_ms,   # config E: dict in -> dict out, host packing included
                   "round_schedule_ms": rs_ms,   # ShockwaveScheduler.round_schedule(): host packing + forecast + solve + lists
                   "mc_forecast": {"paths_per_job": P_MC, "paths_this_rank": mc_n, "horizon_epochs": "min(E-c, 256)",
                                   "kernel_ms": mc_ms, "allreduce": "NCCL SUM of [2][J] float64 (64 KiB)" if world > 1 else "none (1 GPU)"},
                   "e2e_pipeline": "2 engines x 2 host threads (double buffering): the copies of step i+1 overlap the kernels "
                                   "of step i; all host<->device copies of all steps inside the timed region; the same "
                                   f"steps one after the other on one engine: {S * K / e2e_serial_s:.0f} rounds/s on this rank",
                   "l2": "flushed between timed steps (256 MiB write); per-step CUDA events on the launching stream, summed",
                   "parallelism": f"scenario-sharded x{world}, no data-path collective",
                   "scaling_note": "scenarios per GPU are fixed (weak scaling); the 8192 Monte-Carlo paths per job are a fixed "
                                   "TOTAL sharded over the ranks (BASELINE config D), so the forecast's share of the step "
                                   "shrinks with N and value can grow faster than N"},
        "clocks": clocks,
        "gpu_launches": launches,
        "kernels_ms": {"solve_kernel": ms_solve, "place_kernel": ms_place, "gbm_kernel": mc_ms},
        "e2e": {"value": e2e_val, "unit": "rounds/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        # the bounding roofline of this path is HBM bandwidth of the dense PR-dynamics kernel in its
        # scenario-batched form; the Monte-Carlo / search kernels of the timed step are ALU- and issue-bound and are
        # listed with those bounds in step_kernels (never against HBM).
        "roofline": {
            "bound": "hbm", "kernel": "market_step_kernel<1,4> (dense PR-dynamics pass over X[S][J][W][T], fp32)",
            "achieved": dense_bytes / (dense_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": dense_bytes / (dense_ms * 1e-3) / 1e9 / peak,
            "traffic": 1031965952,      # dram__bytes_read+write per launch, profiles/ncu_summary_r01.txt (ncu --set full)
            "algorithmic_bytes_per_launch": dense_bytes, "shape": [Sd, J, Wd, T], "ms_per_launch": dense_ms,
            "note": "kernel designated by SURVEY.md §8(d); algorithmic bytes = 8*S*J*W*T + S*(24*J + 8*W*T); tensor "
                    "(512 MiB) larger than L2; timed live in this run with CUDA events on the launching stream around "
                    "each pass of an 8-iteration run, in its own leg (the collapsed Shockwave solve of the timed step "
                    "never materialises X, DESIGN.md §2); peak = MEASURED_PEAKS.json hbm_gbs"
                    + ("" if peaks else " (fallback 6650)")},
        "step_kernels": [
            {"kernel": "gbm_kernel", "ms": mc_ms, "bound": "alu/sfu",
             "evidence": "86 % issue-active, 75 % SM throughput (profiles/ncu_summary_r01.txt); 28 B in / 16 B out per job"},
            {"kernel": "solve_kernel", "ms": ms_solve, "bound": "instruction issue",
             "evidence": "77 % issue-active, 59 % SM throughput; per-job scalars staged once in shared memory, 76 MB DRAM per launch"},
            {"kernel": "place_kernel", "ms": ms_place, "bound": "latency (sorts, scans and a sequential packer in shared memory)",
             "evidence": "44 % issue-active, 202 MB DRAM per launch (145 MB of it the J x T byte matrices it writes)"}],
        "step_dominant_kernel": {"kernel": dom, "hbm_gbs": achieved, "note": bound_note},
    }
    if world == 1 and not args.no_cpu_baseline:
        from oracle import shockwave_milp as om
        from tests.synth import synth_problem
        logv = om.pwl_log_values(BASES, ORIGIN)
        pb = synth_problem(J, G, T, D, seed=1000, tight=3.0)
        t0 = time.perf_counter()
        cap = om.ftf_caps(pb["rem"], pb["ftobj"], G, J, T, D, pb["round_ptr"], 1.0)
        ok, _, _, _ = om._solve(pb["g"].astype(np.int64), pb["E"].astype(float), pb["c"].astype(float), pb["dbar"],
                                pb["rem"], np.ones(J), G, T, D, 1e-3, BASES, logv, cap, 1e-3, 15.0)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {
            "value": 1.0 / dt, "unit": "rounds/s", "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
            "sample": ("1 scenario of the same workload: first MILP of dynamic_eisenberg_gale_scheduling (model build + "
                       f"HiGHS, mip_rel_gap=1e-3, time_limit=15 s = the reference's Gurobi settings); took {dt:.1f} s, "
                       f"incumbent found: {bool(ok)}; HiGHS stand-in for Gurobi (not installable: no package/licence/network)")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
