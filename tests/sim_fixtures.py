"""Helpers of the simulator-loop tests: ctypes binding of the HOST build of sim_core.cuh (tests/native/sim_host.cpp),
random static traces, a random work-conserving policy."""
import ctypes as C
import math
import os
import shutil
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Scn(C.Structure):
    _fields_ = [("now", C.c_double), ("round_start", C.c_double), ("round_end", C.c_double), ("rounds", C.c_int32),
                ("remaining", C.c_int32), ("n_active", C.c_int32), ("done", C.c_int32), ("err", C.c_int32),
                ("pad", C.c_int32)]


def host_sim_lib():
    if shutil.which("g++") is None:
        return None
    src = os.path.join(ROOT, "tests", "native", "sim_host.cpp")
    out = os.path.join(ROOT, "tests", "native", "libsim_host.so")
    core = os.path.join(ROOT, "shockwave_b200", "csrc", "sim_core.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", src, "-o", out])
    lib = C.CDLL(out)
    lib.sim_host_create.restype = C.c_void_p
    lib.sim_host_create.argtypes = [C.c_int] + [C.c_void_p] * 7
    lib.sim_host_begin.argtypes = [C.c_void_p, C.POINTER(Scn)]
    lib.sim_host_set_dynamic.argtypes = [C.c_void_p] * 9 + [C.c_int] + [C.c_void_p] * 2
    lib.sim_host_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.POINTER(Scn)] + [C.c_void_p] * 5
    lib.sim_host_results.argtypes = [C.c_void_p] * 4
    lib.sim_host_destroy.argtypes = [C.c_void_p]
    lib.sim_host_job_state.argtypes = [C.c_void_p] * 8
    lib.sim_host_set_worker_types.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return lib


class HostSim:
    """One scenario on the host build; same call sequence as swb_sim_begin / swb_sim_step / swb_sim_results."""

    def __init__(self, lib, tr, ngpus, tpi, grd):
        self.lib, self.ngpus, self.tpi, self.grd = lib, ngpus, float(tpi), float(grd)
        self.J = J = len(tr["arrival"])
        a = lambda k, dt: np.ascontiguousarray(tr[k], dtype=dt)
        self._keep = [a("arrival", np.float64), a("total_steps", np.int64), a("scale_factor", np.int32),
                      a("throughput", np.float64), a("duration", np.float64), a("batch_size", np.int32),
                      a("dataset_len", np.int64)]
        self.h = lib.sim_host_create(J, *[x.ctypes.data for x in self._keep])
        self.scn = Scn()
        self.status = np.zeros(J, np.uint8); self.epoch = np.zeros(J, np.int32); self.tl_ns = np.zeros(J)
        self.tl_end = np.zeros(J, np.int32); self.thr_meas = np.zeros(J)

    def set_dynamic(self, dyn):
        from shockwave_b200.simulate import pack_dynamic_tables
        a = self._dyn = pack_dynamic_tables(dyn, self.J)
        p = lambda k: a[k].ctypes.data
        self.lib.sim_host_set_dynamic(self.h, p("mode"), p("bs_max"), p("bs_min"), p("bs_big"), p("orig_locked"),
                                      p("acc_skip"), p("pat_off"), p("pattern"), a["K"], p("lvl_bs"), p("lvl_thr"))

    def set_worker_types(self, throughput, ngpus):
        self._wt = (np.ascontiguousarray(throughput, dtype=np.float64), np.ascontiguousarray(ngpus, dtype=np.int32))
        assert self._wt[0].shape == (self.J, self._wt[1].shape[0])
        self.lib.sim_host_set_worker_types(self.h, int(self._wt[1].shape[0]), self._wt[0].ctypes.data, self._wt[1].ctypes.data)

    def begin(self):
        self.lib.sim_host_begin(self.h, C.byref(self.scn))
        return self.scn

    def step(self, chosen):
        """chosen: job indices (one worker type), {job: worker type index}, or the [J] uint8 row swb_sim_step takes."""
        if isinstance(chosen, np.ndarray) and chosen.shape == (self.J,) and chosen.dtype == np.uint8:
            ch = np.ascontiguousarray(chosen)
        else:
            ch = np.zeros(self.J, np.uint8)
            if isinstance(chosen, dict):
                for j, w in chosen.items():
                    ch[j] = 1 + w
            else:
                ch[list(chosen)] = 1
        self.lib.sim_host_step(self.h, ch.ctypes.data, self.ngpus, self.tpi, self.grd, C.byref(self.scn),
                               self.status.ctypes.data, self.epoch.ctypes.data, self.tl_ns.ctypes.data,
                               self.tl_end.ctypes.data, self.thr_meas.ctypes.data)
        return self.scn

    def job_state(self):
        J = self.J
        o = dict(total_steps=np.zeros(J, np.int64), throughput=np.zeros(J), batch_size=np.zeros(J, np.int32),
                 exec_time=np.zeros(J), finish_time=np.zeros(J), failed_attempts=np.zeros(J, np.uint8), ran=np.zeros(J, np.uint8))
        self.lib.sim_host_job_state(self.h, *[o[k].ctypes.data for k in (
            "total_steps", "throughput", "batch_size", "exec_time", "finish_time", "failed_attempts", "ran")])
        return o

    def results(self):
        jct = np.zeros(self.J); sr = np.zeros(self.J, np.int64); rt = np.zeros(self.J)
        self.lib.sim_host_results(self.h, jct.ctypes.data, sr.ctypes.data, rt.ctypes.data)
        return jct, sr, rt

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.sim_host_destroy(self.h)
            self.h = None


def random_trace(J, seed, gap=False):
    rng = np.random.default_rng(seed)
    arrival = np.sort(rng.uniform(0, 4000.0 if not gap else 60000.0, J))
    if not gap:
        arrival[: max(1, J // 4)] = 0.0
    thr = rng.uniform(0.3, 40.0, J)
    bs = rng.choice([16, 32, 64, 128], J)
    ds = rng.choice([50000, 100000, 10000, 59675], J)
    spe = np.ceil(ds / bs)
    epochs = rng.integers(1, 12, J)
    total = (epochs * spe - rng.integers(0, 3, J)).astype(np.int64)
    total = np.maximum(total, 1)
    sf = rng.choice([1, 1, 1, 2, 4], J)
    ideal = total / thr
    duration = ideal * rng.uniform(0.55, 1.6, J)          # some jobs hit the over-deadline rule (1.5 x duration)
    return dict(arrival=arrival, total_steps=total, scale_factor=sf.astype(np.int32), throughput=thr,
                duration=duration, batch_size=bs.astype(np.int32), dataset_len=ds.astype(np.int64))


def random_policy(tr, ngpus, seed):
    """A random, mostly work-conserving gang selection: stable within a run (keyed by the round index)."""
    sf = np.asarray(tr["scale_factor"])

    def select(c, now, active):
        rng = np.random.default_rng(seed * 100003 + c)
        order = rng.permutation(len(active))
        left = ngpus
        out = []
        for i in order:
            j = active[i]
            if sf[j] <= left and (rng.random() < 0.9 or not out):
                out.append(j)
                left -= int(sf[j])
        return out
    return select


def timeline_summary(entries, grd):
    """JobMetaData.py:235-249 on a list of (round, throughput, bs)."""
    ns, prev = 0, 0
    for cur, thr, bs in entries:
        ns += bs * (thr * grd * (cur - prev))
        prev = cur
    return float(ns), (entries[-1][0] if entries else -1)


class HostDeviceSim:
    """Stand-in for shockwave_b200.simulate.DeviceSim on the host build (same attributes and call sequence): lets the
    CPU tests run the ensemble drivers' host logic.  Test tooling only."""

    def __init__(self, trace, S, ngpus, time_per_iteration=120.0, round_duration=None, device=0):
        from shockwave_b200.simulate import SCN_DTYPE
        lib = host_sim_lib()
        rd = time_per_iteration if round_duration is None else round_duration
        self.S, self.J = int(S), len(trace["arrival"])
        self.sims = [HostSim(lib, trace, ngpus, time_per_iteration, rd) for _ in range(self.S)]
        self.scn = np.zeros(self.S, dtype=SCN_DTYPE)
        self.status = np.zeros((self.S, self.J), np.uint8)
        self.epoch = np.zeros((self.S, self.J), np.int32)
        self.tl_ns = np.zeros((self.S, self.J))
        self.tl_end = np.full((self.S, self.J), -1, np.int32)

    def set_dynamic(self, dyn):
        for m in self.sims:
            m.set_dynamic(dyn)

    def set_worker_types(self, throughput, ngpus):
        for m in self.sims:
            m.set_worker_types(throughput, ngpus)

    def _pull(self, s, z):
        for k in ("now", "round_start", "round_end", "rounds", "remaining", "n_active", "done", "err"):
            self.scn[k][s] = getattr(z, k)

    def begin(self):
        arrival = np.asarray(self.sims[0]._keep[0])
        for s, m in enumerate(self.sims):
            z = m.begin()
            self._pull(s, z)
            self.status[s] = arrival <= z.now
        return self.scn

    def step(self, chosen):
        for s, m in enumerate(self.sims):
            z = m.step(np.ascontiguousarray(chosen[s], dtype=np.uint8))
            self._pull(s, z)
            self.status[s], self.epoch[s], self.tl_ns[s], self.tl_end[s] = m.status, m.epoch, m.tl_ns, m.tl_end
        return self.scn

    def replay(self, schedule):
        self.begin()
        for r in range(len(schedule)):
            ch = schedule[r] if schedule.ndim == 3 else np.broadcast_to(schedule[r], (self.S, self.J))
            self.step(ch)
        return self.scn

    def results(self):
        r = [m.results() for m in self.sims]
        return dict(jct=np.stack([x[0] for x in r]), steps_run=np.stack([x[1] for x in r]),
                    run_time=np.stack([x[2] for x in r]), measured_throughput=np.stack([m.thr_meas for m in self.sims]))

    def job_state(self):
        parts = [m.job_state() for m in self.sims]
        return {k: np.stack([p[k] for p in parts]) for k in parts[0]}

    def close(self):
        pass


def make_rule_scheduler_cls():
    """A ShockwaveScheduler whose device calls are replaced by a deterministic RULE on exactly the arrays the real
    re-solve uploads (epoch progress, timeline summaries, round pointer): run inside the reference's loop and inside the
    ensemble driver it makes any difference in how those arrays are maintained visible as a different schedule."""
    from collections import OrderedDict

    from shockwave_b200.scheduler import ShockwaveScheduler as Base

    class _NoEngine:
        def job_add(self, *a, **k):
            pass

        def job_remove(self, *a, **k):
            pass

        def set_option(self, *a, **k):
            pass

    class RuleScheduler(Base):
        log = None

        def _eng(self):
            if self._engine is None:
                self._engine = _NoEngine()
            return self._engine

        def _resolve(self, jobids, jobobjs):
            J = len(jobids)
            slots = np.fromiter(map(self._slots.__getitem__, jobids), dtype=np.int64, count=J)
            prog = self._prog[slots]
            ns, end = self._timeline_summaries(jobids, jobobjs, slots)
            g = np.array([j.nworkers for j in jobobjs])
            sub = np.array([j.timestamp_submit for j in jobobjs])
            if self.log is not None:
                self.log.append((self.round_ptr, list(jobids), prog.copy(), np.array(ns, float), np.array(end, int), sub))
            out = OrderedDict()
            for t in range(self.future_nrounds):
                score = (prog * 7 + np.asarray(end) * 3 + (np.floor(np.asarray(ns)) % 11) + np.array(jobids) * 5 +
                         (self.round_ptr + t) * 2 + np.floor(sub) % 3) % 13
                left, names = self.ngpus, []
                for i in sorted(range(J), key=lambda i: (score[i], jobids[i])):
                    if g[i] <= left:
                        names.append(jobids[i])
                        left -= int(g[i])
                out[self.round_ptr + t] = names
            return out
    return RuleScheduler
