"""CPU tests of the packing-policy path (no GPU):
  * the oracle's array construction (oracle/gavel_packed.py flatten / scale_factors_array) is pinned BIT FOR BIT on the
    reference's own `PolicyWithPacking` (policy.py:68-160), imported from the staged reference with a stub `cvxpy`
    (those methods are pure numpy; only the solve needs cvxpy);
  * shockwave_b200/packing.py's HOST logic (sparse column model, multi-section, stateful bookkeeping, the bisection
    replay) against the oracle, with HiGHS standing in for swb_lp_solve through the module's `_lp` hook;
  * the DEVICE simplex source (csrc/lp_core.cuh) compiled for the host (tests/native/lp_host.cpp, one thread) against
    HiGHS on random programs — optimal, infeasible, unbounded, degenerate — and on packed programs."""
import ctypes as C
import importlib
import os
import shutil
import subprocess
import sys
import types

import numpy as np
import pytest
from scipy.optimize import linprog

from oracle import gavel_packed as gp
from oracle import ref_harness
from tests.packing_fixtures import JobId, instance

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPEC = {"v100": 8, "p100": 6, "k80": 4}
COSTS = {"k80": 1.0, "p100": 2.0, "v100": 3.0}


@pytest.fixture()
def highs_backend(monkeypatch):
    from shockwave_b200 import packing as pk
    monkeypatch.setattr(pk, "_lp", gp.lp_backend)
    return pk


def _base_ok(alloc, thr, sf, spec):
    ids = sorted(alloc.keys())
    wts = sorted(spec.keys())
    x = np.array([[alloc[c][w] for w in wts] for c in ids])
    assert x.min() >= 0 and x.max() <= 1 + 1e-12
    sfc = gp.scale_factors_array(sf, ids, len(ids), len(wts))
    assert np.all((x * sfc).sum(axis=0) <= np.array([spec[w] for w in wts]) * (1 + 1e-9) + 1e-9)
    assert np.all(x[sfc == 0] == 0)
    for s in [c for c in ids if not c.is_pair()]:
        share = sum(x[i].sum() for i, c in enumerate(ids) if s in c.singletons())
        assert share <= 1 + 1e-9
    return x


@pytest.mark.skipif(not ref_harness.reference_available(), reason="staged reference not present")
def test_oracle_flatten_pinned_on_reference():
    pol_dir = os.path.join(ref_harness.REF, "policies")
    saved = {k: sys.modules.get(k) for k in ("cvxpy", "policy", "job_id_pair")}
    sys.modules["cvxpy"] = types.ModuleType("cvxpy")          # policy.py imports it at the top; flatten never calls it
    sys.path[:0] = [pol_dir, ref_harness.REF]
    try:
        for k in ("policy", "job_id_pair"):
            sys.modules.pop(k, None)
        ref_policy = importlib.import_module("policy")
        ref_pair = importlib.import_module("job_id_pair").JobIdPair
        for ns, pf, seed in [(5, 1.0, 1), (12, 0.6, 2), (20, 0.3, 3)]:
            thr, sf, prio, _, _, spec, _ = instance(ns, SPEC, seed, pf, make_id=ref_pair)
            for pw in (None, prio):
                P = ref_policy.PolicyWithPacking()
                ref_m, ref_idx = P.flatten(thr, spec, priority_weights=pw)
                my_m, my_idx = gp.flatten(thr, spec, priority_weights=pw)
                assert ref_m.dtype == my_m.dtype and np.array_equal(ref_m, my_m)
                assert list(ref_idx[0]) == list(my_idx[0]) and list(ref_idx[1]) == list(my_idx[1])
                assert ref_idx[2] == my_idx[2] and ref_idx[3] == my_idx[3]
                m, n = ref_m[0].shape
                assert np.array_equal(P.scale_factors_array(sf, ref_idx[0], m, n),
                                      gp.scale_factors_array(sf, my_idx[0], m, n))
            # the tests' own JobId orders, hashes and splits like the reference's JobIdPair
            a = sorted(thr.keys())
            b = sorted(JobId(*k.as_tuple()) for k in thr)
            assert [k.as_tuple() for k in a] == [k.as_tuple() for k in b]
    finally:
        sys.path[:] = [p for p in sys.path if p not in (pol_dir, ref_harness.REF)]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.parametrize("ns,pf", [(6, 1.0), (14, 0.5), (24, 1.0)])
def test_host_logic_against_oracle(highs_backend, ns, pf):
    pk = highs_backend
    thr, sf, prio, t0, steps, spec, singles = instance(ns, SPEC, seed=100 + ns, pair_fraction=pf)
    z, _, _ = gp.max_min_fairness_packed(thr, sf, prio, spec)
    pol = pk.MaxMinFairnessPolicyWithPacking("ECOS")
    _base_ok(pol.get_allocation(thr, sf, prio, spec), thr, sf, spec)
    assert abs(pol.last_objective - z) <= 1e-7 * z

    T, _, _ = gp.min_total_duration_packed(thr, sf, steps, spec)
    pol = pk.MinTotalDurationPolicyWithPacking("ECOS")
    x = _base_ok(pol.get_allocation(thr, sf, steps, spec), thr, sf, spec)
    assert pol.last_objective == T                      # the same probe sequence as the reference's bisection
    ids = sorted(thr.keys())
    P = gp.Packed(thr, sf, spec)
    for i, s in enumerate(singles):                     # every job finishes within the returned T
        assert P.coef(i) @ x.ravel() >= steps[s] / T * (1 - 1e-7)

    # two stateful rounds of finish-time fairness (cumulative isolated time carried over)
    pol = pk.FinishTimeFairnessPolicyWithPacking("ECOS")
    cum = {s: 0.0 for s in singles}
    steps_now = dict(steps)
    iso_prev = None
    for rnd in range(2):
        if rnd == 1:
            Pk = gp.Packed(thr, sf, spec, prio)
            iso = gp.isolated_throughputs(Pk.thr_single, Pk.sf_single, Pk.N)
            steps_next = {s: steps_now[s] * 0.9 for s in singles}
            for i, s in enumerate(singles):
                cum[s] += (steps_now[s] - steps_next[s]) / iso[i]
            steps_now = steps_next
        rho, _, _ = gp.finish_time_fairness_packed(thr, sf, prio, t0, steps_now, cum, spec)
        _base_ok(pol.get_allocation(thr, sf, prio, t0, steps_now, spec), thr, sf, spec)
        assert abs(pol.last_objective - rho) <= 1e-6 * rho, (rnd, pol.last_objective, rho)
        assert pol.last_passes <= 12

    slo = {singles[i]: steps[singles[i]] / (thr[singles[i]]["k80"] * 0.1) for i in range(0, ns, 5)}
    o, _, used, _ = gp.max_sum_throughput_packed_slos(thr, sf, spec, COSTS, slo, steps)
    pol = pk.ThroughputNormalizedByCostSumWithPackingSLOs("ECOS")
    _base_ok(pol.get_allocation(thr, sf, spec, COSTS, slo, steps), thr, sf, spec)
    assert pol.used_SLOs == used and abs(pol.last_objective - o) <= 1e-7 * o
    # impossible SLOs: the reference's "solve again without the SLO rows"
    slo_bad = {singles[0]: 1e-3}
    o2, _, used2, _ = gp.max_sum_throughput_packed_slos(thr, sf, spec, COSTS, slo_bad, steps)
    pol.get_allocation(thr, sf, spec, COSTS, slo_bad, steps)
    assert not used2 and not pol.used_SLOs and abs(pol.last_objective - o2) <= 1e-7 * o2


def test_packed_without_pairs_is_the_perf_program(highs_backend):
    """Only single-job keys: the packed program IS MaxMinFairnessPolicyWithPerf's LP (max_min_fairness.py:53-113)."""
    from oracle import gavel_lp as gl
    pk = highs_backend
    thr, sf, prio, _, _, spec, singles = instance(18, SPEC, seed=5, pair_fraction=0.0)
    pol = pk.MaxMinFairnessPolicyWithPacking("ECOS")
    pol.get_allocation(thr, sf, prio, spec)
    wts = sorted(spec.keys())
    a = np.array([[thr[s][w] for w in wts] for s in singles])
    z, _ = gl.max_min_fairness_perf(a, np.array([sf[s] for s in singles], float),
                                    np.array([prio[s] for s in singles]), [spec[w] for w in wts])
    assert abs(pol.last_objective - z) <= 1e-6 * z       # all_m is float32 in the packed classes (policy.py:132)


def test_empty_and_mismatched_scale_factors(highs_backend):
    pk = highs_backend
    assert pk.MaxMinFairnessPolicyWithPacking("ECOS").get_allocation({}, {}, {}, SPEC) is None
    thr, sf, prio, _, _, spec, singles = instance(4, SPEC, seed=9, pair_fraction=1.0)
    sf = {s: (1 if i % 2 == 0 else 2) for i, s in enumerate(singles)}        # pairs (0,1), (0,3), ... mismatch
    alloc = pk.MaxMinFairnessPolicyWithPacking("ECOS").get_allocation(thr, sf, prio, spec)
    for c, row in alloc.items():
        if c.is_pair() and sf[c.singletons()[0]] != sf[c.singletons()[1]]:
            assert all(v == 0 for v in row.values())
    z, _, _ = gp.max_min_fairness_packed(thr, sf, prio, spec)
    # a cluster with an empty worker type
    spec0 = dict(spec, k80=0)
    alloc = pk.MaxMinFairnessPolicyWithPacking("ECOS").get_allocation(thr, sf, prio, spec0)
    assert all(row["k80"] == 0 for row in alloc.values())


# ---- the device simplex source on the host ----
@pytest.fixture(scope="module")
def host_simplex():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
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
            v, cc, bb = (np.ascontiguousarray(a[s]) for a in (val, c, b))
            lib.lp_host_solve(m, n, p(colp), p(rowi), p(v), p(cc), p(bb), max_iter or 50 * (m + n) + 1000, p(x[s]), p(out8))
            obj[s], st[s], stats[s] = out8[0], int(out8[1]), out8[2:6]
        return x, obj, st, stats
    return solve


def random_lp(rng, kind):
    import scipy.sparse as sp
    m = int(rng.integers(2, 40)); n = int(rng.integers(2, 80))
    A = sp.random(m, n, density=rng.uniform(0.05, 0.5), random_state=int(rng.integers(1 << 30)),
                  data_rvs=lambda k: rng.uniform(-1, 2, k)).tocsc()
    A.sort_indices()
    b = rng.uniform(0.5, 5, m) if kind == 0 else rng.uniform(-2, 5, m)
    if kind == 2:
        b = np.round(b)            # ties in the ratio test, zero right-hand sides
    return A, b, rng.uniform(-1, 2, n)


def test_device_simplex_source_on_random_programs(host_simplex):
    rng = np.random.default_rng(0)
    seen = set()
    for trial in range(240):
        A, b, c = random_lp(rng, trial % 3)
        x, obj, st, _ = host_simplex(A.indptr, A.indices, A.data, c, b)
        r = linprog(-c, A_ub=A, b_ub=b, bounds=(0, None), method="highs")
        if r.status == 4:
            continue
        assert int(st[0]) == {0: 0, 2: 1, 3: 2}[r.status], (trial, st, r.status)
        seen.add(int(st[0]))
        if r.status == 0:
            assert abs(obj[0] + r.fun) <= 1e-7 * (1 + abs(r.fun))
            assert (A @ x[0] - b).max() <= 1e-7 and x[0].min() >= -1e-9
    assert seen == {0, 1, 2}


def test_device_simplex_source_on_packed_programs(host_simplex, monkeypatch):
    from shockwave_b200 import packing as pk
    monkeypatch.setattr(pk, "_lp", host_simplex)
    for ns, pf, ident in [(16, 1.0, False), (30, 0.5, False), (24, 1.0, True)]:
        thr, sf, prio, t0, steps, spec, singles = instance(ns, SPEC, seed=ns, pair_fraction=pf)
        if ident:       # every job alike: maximally degenerate max-min program (Bland's rule engages)
            for k in thr:
                for w in thr[k]:
                    thr[k][w] = [0.6, 0.6] if k.is_pair() else 1.0
            sf = {s: 1 for s in singles}
            prio = {s: 1.0 for s in singles}
        z, _, _ = gp.max_min_fairness_packed(thr, sf, prio, spec)
        pol = pk.MaxMinFairnessPolicyWithPacking("ECOS")
        _base_ok(pol.get_allocation(thr, sf, prio, spec), thr, sf, spec)
        assert abs(pol.last_objective - z) <= 1e-9 * z
        cum = {s: 0.0 for s in singles}
        rho, _, _ = gp.finish_time_fairness_packed(thr, sf, prio, t0, steps, cum, spec)
        pol = pk.FinishTimeFairnessPolicyWithPacking("ECOS")
        pol.get_allocation(thr, sf, prio, t0, steps, spec)
        assert abs(pol.last_objective - rho) <= 1e-6 * rho


WT6 = ["k80", "k80_unconsolidated", "p100", "p100_unconsolidated", "v100", "v100_unconsolidated"]


def six_type_instance(J, seed):
    """All six worker types of tacc_throughputs.json with capacity: beyond hetero.cu's master (W <= 4)."""
    rng = np.random.default_rng(seed)
    speed = np.array([1.0, 0.8, 2.2, 1.9, 3.5, 3.0])
    thr = {j: {w: float(rng.uniform(0.5, 20.0) * speed[k] * rng.uniform(0.6, 1.0)) for k, w in enumerate(WT6)}
           for j in range(J)}
    sf = {j: int(rng.choice([1, 2, 4], p=[0.7, 0.2, 0.1])) for j in range(J)}
    prio = {j: float(rng.choice([1.0, 2.0, 5.0])) for j in range(J)}
    t0 = {j: float(rng.uniform(0.0, 5000.0)) for j in range(J)}
    steps = {j: float(rng.uniform(2e3, 4e5)) for j in range(J)}
    spec = {w: int(n) for w, n in zip(WT6, [6, 3, 5, 2, 8, 4])}
    return thr, sf, prio, t0, steps, spec


def check_six_types(P, J, seed):
    """Shared by the CPU test (HiGHS / host simplex behind packing._lp) and the GPU test."""
    from oracle import gavel_backend as gb
    from oracle import gavel_lp as gl
    thr, sf, prio, t0, steps, spec = six_type_instance(J, seed)
    a = np.array([[thr[j][w] for w in WT6] for j in range(J)])
    s = np.array([sf[j] for j in range(J)], float)
    N = np.array([spec[w] for w in WT6], float)

    def mat(alloc):
        x = np.array([[alloc[j][w] for w in WT6] for j in range(J)])
        assert x.min() >= 0 and np.all(x.sum(axis=1) <= 1 + 1e-9) and np.all((x * s[:, None]).sum(axis=0) <= N * (1 + 1e-9))
        return x
    pol = P.MaxMinFairnessPolicyWithPerf(solver="ECOS")
    mat(pol.get_allocation(thr, sf, prio, spec))
    z, _ = gl.max_min_fairness_perf(a, s, np.array([prio[j] for j in range(J)]), N)
    assert abs(pol.last_objective - z) <= 1e-6 * z
    pol = P.MinTotalDurationPolicyWithPerf(solver="ECOS")
    x = mat(pol.get_allocation(thr, sf, steps, spec))
    n = np.array([steps[j] for j in range(J)])
    T, _ = gl.min_total_duration_perf(a, s, n, N)
    assert pol.last_objective == T and np.all((a * x).sum(axis=1) >= n / T * (1 - 1e-7))
    pol = P.FinishTimeFairnessPolicyWithPerf(solver="ECOS")
    x = mat(pol.get_allocation(thr, sf, prio, t0, steps, spec))
    iso = gp.isolated_throughputs(a, s, N)
    t = np.array([t0[j] for j in range(J)])
    _, rho = gb._ftf(a, s, t, n, n / iso, N)
    assert abs(pol.last_objective - rho) <= 1e-6 * rho
    assert np.max((t + n / (a * x).sum(axis=1)) / (n / iso)) <= rho * (1 + 1e-6)
    costs = {w: c for w, c in zip(WT6, [1.0, 0.9, 2.0, 1.8, 3.0, 2.7])}
    slo = {j: steps[j] / (thr[j]["k80"] * 0.2) for j in range(0, J, 4)}
    pol = P.ThroughputNormalizedByCostSumWithPerfSLOs(solver="ECOS")
    x = mat(pol.get_allocation(thr, sf, spec, instance_costs=costs, SLOs=slo, num_steps_remaining=steps))
    need = np.zeros(J)
    for j in slo:
        need[j] = steps[j] / slo[j]
    cst = np.array([costs[w] for w in WT6])
    v, _ = gl.max_sum_throughput(a, s, N, costs=cst, need=need)
    assert v is not None and abs(pol.last_objective - v) <= 1e-6 * v
    assert np.all((a * x).sum(axis=1) >= need * (1 - 1e-7))


@pytest.mark.parametrize("backend", ["highs", "host_simplex"])
def test_six_worker_types_host_logic(backend, monkeypatch, request):
    from shockwave_b200 import packing as pk
    from shockwave_b200 import policies as P
    monkeypatch.setattr(pk, "_lp", gp.lp_backend if backend == "highs" else request.getfixturevalue("host_simplex"))
    # IsolatedPolicy._alloc (finish-time fairness) is a device call too: restate it for this CPU test
    monkeypatch.setattr(P.IsolatedPolicy, "_alloc", lambda self, thr, sf: (
        lambda x: x / np.maximum(x.sum(axis=1), 1.0)[:, None])((np.asarray(self._num_workers, float)[None, :] / len(sf)) / sf[:, None]))
    check_six_types(P, 40, seed=11)


@pytest.mark.parametrize("backend", ["highs", "host_simplex"])
def test_water_filling_packed_host_logic(backend, monkeypatch, request):
    """The packed water-filling class: tight LP relaxation of the bottleneck MILP vs the MILP itself (scipy milp),
    iteration by iteration, with entity re-weighting on the last instance."""
    from shockwave_b200 import packing as pk
    monkeypatch.setattr(pk, "_lp", gp.lp_backend if backend == "highs" else request.getfixturevalue("host_simplex"))
    for ns, pf, seed in [(6, 1.0, 1), (10, 1.0, 2), (14, 0.5, 3), (12, 0.7, 6)]:
        thr, sf, prio, _, _, spec, singles = instance(ns, SPEC, seed=seed, pair_fraction=pf)
        x, eff, it, log, _ = gp.water_filling_packed(thr, sf, prio, spec)
        pol = pk.MaxMinFairnessWaterFillingPolicyWithPacking()
        e2, ids = pol.get_allocation(thr, sf, prio, spec, return_effective_throughputs=True)
        assert ids == singles and pol.last_iterations == it
        assert np.max(np.abs(eff - e2) / np.maximum(eff, 1e-9)) <= 1e-6
        _base_ok(pol.get_allocation(thr, sf, prio, spec), thr, sf, spec)
    ent = {"A": singles[:5], "B": singles[5:]}
    ew = {"A": 1.0, "B": 2.0}
    pols = {"A": "fairness", "B": "fifo"}
    x, eff, it, log, _ = gp.water_filling_packed(thr, sf, prio, spec, entity_weights=ew,
                                                 entity_to_job_mapping={k: list(v) for k, v in ent.items()}, policies=pols)
    pol = pk.MaxMinFairnessWaterFillingPolicyWithPacking(priority_reweighting_policies=pols)
    e2, _ = pol.get_allocation(thr, sf, prio, spec, entity_weights=ew,
                               entity_to_job_mapping={k: list(v) for k, v in ent.items()}, return_effective_throughputs=True)
    assert pol.last_iterations == it and np.max(np.abs(eff - e2) / np.maximum(eff, 1e-9)) <= 1e-6


# ---- the job-TYPE formulation (max_min_fairness.py:122-316 + policy.py:195-260) ----
@pytest.mark.parametrize("backend", ["highs", "host_simplex"])
def test_job_type_formulation_host_logic(backend, monkeypatch, request):
    from shockwave_b200 import packing as pk
    from tests.packing_fixtures import job_type_instance
    monkeypatch.setattr(pk, "_lp", gp.lp_backend if backend == "highs" else request.getfixturevalue("host_simplex"))
    for n, nt, seed in [(6, 2, 1), (12, 3, 2), (30, 5, 3), (9, 9, 4)]:
        thr, j2k, sf, prio = job_type_instance(n, nt, SPEC, seed)
        z, x_or, (job_ids, keys, wts) = gp.max_min_fairness_job_types(thr, j2k, sf, prio, SPEC)
        pol = pk.MaxMinFairnessPolicyWithPacking("ECOS")
        out = pol.get_allocation_using_job_type_throughputs(thr, j2k, sf, prio, SPEC)
        assert abs(pol.last_objective - z) <= 1e-7 * max(1.0, abs(z)), (n, nt, pol.last_objective, z)
        # the job x job-type point satisfies the reference's rows
        al = pol.last_job_type_allocation
        for j in job_ids:
            assert sum(al[j][w][o] for w in wts for o in [None] + keys) <= 1 + 1e-7
        for w in wts:
            used = sum(sf[j] * (al[j][w][None] + 0.5 * sum(al[j][w][o] for o in keys)) for j in job_ids)
            assert used <= SPEC[w] + 1e-6
            for a_i, ka in enumerate(keys):
                mem_a = [j for j in job_ids if j2k[j] == ka]
                vals = [al[j][w][ka] for j in mem_a]
                assert max(vals) - min(vals) <= 1e-9                      # i-A variables of type A all equal
                for kb in keys[a_i + 1:]:
                    if kb[1] != ka[1]:
                        continue
                    mem_b = [j for j in job_ids if j2k[j] == kb]
                    assert abs(sum(al[j][w][kb] for j in mem_a) - sum(al[j][w][ka] for j in mem_b)) <= 1e-7
        # conversion: the oracle's restatement of policy.py:195-260 on the same job-type point
        ref = gp.convert_job_type_allocation(al, j2k, JobId)
        assert set(map(repr, ref)) == set(map(repr, out))
        byrepr = {repr(k): v for k, v in out.items()}
        for k, row in ref.items():
            for w in wts:
                assert abs(row[w] - byrepr[repr(k)][w]) <= 1e-12


@pytest.mark.skipif(not ref_harness.reference_available(), reason="staged reference not present")
def test_job_type_conversion_pinned_on_reference():
    """oracle convert_job_type_allocation == the reference's Policy.convert_job_type_allocation (policy.py:195-260)
    on a random job-type point, with the reference's own JobIdPair."""
    from tests.packing_fixtures import job_type_instance
    pol_dir = os.path.join(ref_harness.REF, "policies")
    saved = {k: sys.modules.get(k) for k in ("cvxpy", "policy", "job_id_pair")}
    sys.modules["cvxpy"] = types.ModuleType("cvxpy")
    sys.path[:0] = [pol_dir, ref_harness.REF]
    try:
        for k in ("policy", "job_id_pair"):
            sys.modules.pop(k, None)
        ref_policy = importlib.import_module("policy")
        ref_pair = importlib.import_module("job_id_pair").JobIdPair
        rng = np.random.default_rng(5)
        for n, nt, seed in [(10, 3, 7), (7, 2, 8)]:
            thr, j2k, sf, prio = job_type_instance(n, nt, SPEC, seed)
            j2k = {ref_pair(j[0], None): k for j, k in j2k.items()}
            keys = sorted(thr.keys())
            al = {j: {w: {o: float(rng.uniform(0, 0.2)) for o in [None] + keys} for w in sorted(SPEC)} for j in j2k}
            want = ref_policy.PolicyWithPacking().convert_job_type_allocation(al, j2k)
            got = gp.convert_job_type_allocation(al, j2k, lambda a, b: ref_pair(a, b))
            assert set(want.keys()) == set(got.keys())
            for k in want:
                for w in want[k]:
                    assert abs(want[k][w] - got[k][w]) <= 1e-15
    finally:
        sys.path[:] = [p for p in sys.path if p not in (pol_dir, ref_harness.REF)]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
