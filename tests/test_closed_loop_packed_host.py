"""P5 for the packing policies on the CPU: the UNMODIFIED reference simulator (`Scheduler.simulate()` with job packing on:
job pairs in the allocation, in the priorities and in the per-round schedule, scheduler.py:1146-1265, :1425-1512,
:3390-3490) drives the PRODUCT's `*_packed` policy classes (shockwave_b200/packing.py: the reference's column model,
flatten / unflatten, search loops) with
  (a) the device simplex SOURCE compiled for the host behind `packing._lp` (tests/native/lp_host.cpp), and
  (b) HiGHS behind the same hook (the yardstick: cvxpy / ECOS are not installable),
on the first jobs of the canonical trace.  What this pins is how far the reference itself gets: its loop takes the
product's allocation (singles AND pairs), builds priorities, and schedules the first job PAIR — and then raises in its
own bookkeeping (`_schedule_jobs_on_workers` records the round with `job_id.integer_job_id()`, scheduler.py:1408, which
asserts on a pair: the Shockwave additions to Gavel's loop assume single jobs).  So a closed-loop (P5) yardstick for the
packed policies does not exist in this reference — with either solver behind the policy — and "job pairs in the
simulator loop" has nothing to be a drop-in for; the packed policies are held to the LP-level parity of
tests/test_gpu_packed.py / tests/test_oracle_packed.py instead."""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest
import scipy.sparse as sp
from scipy.optimize import linprog

from oracle import gavel_backend as gb
from oracle import ref_harness as rh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="staged reference not present")


def host_simplex_solver():
    if shutil.which("g++") is None:
        return None
    src = os.path.join(ROOT, "tests", "native", "lp_host.cpp")
    out = os.path.join(ROOT, "tests", "native", "liblp_host.so")
    core = os.path.join(ROOT, "shockwave_b200", "csrc", "lp_core.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", src, "-o", out])
    lib = C.CDLL(out)
    lib.lp_host_solve.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_void_p]

    def solve(colp, rowi, val, c, b, max_iter=0):
        val = np.atleast_2d(np.ascontiguousarray(val, float)); c = np.atleast_2d(np.ascontiguousarray(c, float))
        b = np.atleast_2d(np.ascontiguousarray(b, float))
        colp = np.ascontiguousarray(colp, np.int32); rowi = np.ascontiguousarray(rowi, np.int32)
        S, n = c.shape
        m = b.shape[1]
        x = np.zeros((S, n)); obj = np.zeros(S); st = np.zeros(S, np.int32); stats = np.zeros((S, 4), np.int32)
        p = lambda a: a.ctypes.data
        for s in range(S):
            out8 = np.zeros(8)
            v, cc, bb = (np.ascontiguousarray(a[s % a.shape[0]]) for a in (val, c, b))
            lib.lp_host_solve(m, n, p(colp), p(rowi), p(v), p(cc), p(bb), max_iter or 50 * (m + n) + 1000, p(x[s]), p(out8))
            obj[s], st[s], stats[s] = out8[0], int(out8[1]), out8[2:6]
        return x, obj, st, stats
    return solve


def highs_solver(colp, rowi, val, c, b, max_iter=0):
    """max c x, A x <= b, x >= 0 per scenario (the contract of swb_lp_solve): status 0 optimal, 1 infeasible, 2 unbounded."""
    val = np.atleast_2d(np.asarray(val, float)); c = np.atleast_2d(np.asarray(c, float)); b = np.atleast_2d(np.asarray(b, float))
    S, n = c.shape
    m = b.shape[1]
    x = np.zeros((S, n)); obj = np.zeros(S); st = np.zeros(S, np.int32); stats = np.zeros((S, 4), np.int32)
    for s in range(S):
        A = sp.csc_matrix((val[s % val.shape[0]], np.asarray(rowi), np.asarray(colp)), shape=(m, n))
        r = linprog(-c[s], A_ub=A, b_ub=b[s % b.shape[0]], bounds=(0, None), method="highs")
        if r.status == 0:
            x[s], obj[s] = r.x, -r.fun
        else:
            st[s] = 1 if r.status == 2 else 2
    return x, obj, st, stats


def _run(policy, lp, keep, cluster, seen):
    from shockwave_b200 import packing as pk
    from tests.golden import make_sim_pins as pins
    scratch = tempfile.mkdtemp(prefix="swpk_")
    pins.stage_static_trace(scratch, keep=keep, static=True)
    saved = pk._lp
    pk._lp = lp
    try:
        with gb.cpu_backend() as P:
            pol = P.get_policy(policy, solver="ECOS", seed=0)
            assert type(pol).__module__ == "shockwave_b200.packing"
            inner = pol.get_allocation

            def recording(*a, **k):
                out = inner(*a, **k)
                seen.append(out)
                return out
            pol.get_allocation = recording
            return rh.simulate(policy, policy_obj=pol, trace=pins.REL, scratch=scratch, cluster=cluster)
    finally:
        pk._lp = saved


@pytest.mark.parametrize("backend", ["host_simplex", "highs"])
def test_packed_policy_in_the_unmodified_reference_loop_reaches_the_first_pair(backend):
    lp = host_simplex_solver() if backend == "host_simplex" else highs_solver
    if lp is None:
        pytest.skip("g++ not available")
    import traceback
    seen = []
    with pytest.raises(AssertionError) as err:
        _run("max_min_fairness_packed", lp, 24, "6:0:0", seen)
    frames = traceback.extract_tb(err.value.__traceback__)
    assert frames[-1].name == "integer_job_id" and frames[-2].name == "_schedule_jobs_on_workers"   # the reference's own code
    # before that the loop consumed the product's allocations: base constraints hold and a pair carries real share
    assert len(seen) >= 2
    alloc = seen[-1]
    pairs = {k: v for k, v in alloc.items() if k.is_pair()}
    assert pairs and max(v["v100"] for v in pairs.values()) > 0.1
    singles = [k for k in alloc if not k.is_pair()]
    for j in singles:
        share = sum(v["v100"] for k, v in alloc.items() if j.overlaps_with(k))
        assert share <= 1.0 + 1e-6
    assert sum(v["v100"] for v in alloc.values()) <= 6.0 + 1e-6


def test_generated_arrival_mode_does_not_start_in_the_reference():
    """`simulate_scheduler_with_generated_jobs.py` / `scripts/sweeps/run_sweep_*.py` build a Scheduler WITHOUT a trace
    pickle (drivers/simulate_scheduler_with_generated_jobs.py:45-56); this fork's `Scheduler.__init__` opens the pickle
    unconditionally (scheduler.py:437), so the generated-arrival mode cannot run in the reference: nothing to drop into."""
    with gb.cpu_backend() as P:
        dst = rh.prepare_tree(None)
        cwd = os.getcwd()
        os.chdir(dst)
        try:
            ref_sched, _ = rh.import_reference(dst)
            with pytest.raises(TypeError):
                ref_sched.Scheduler(P.get_policy("max_min_fairness", solver="ECOS", seed=0),
                                    throughputs_file=os.path.join(dst, "tacc_throughputs.json"), simulate=True, seed=0,
                                    time_per_iteration=120)
        finally:
            os.chdir(cwd)
