"""CPU tests of the Gavel-policy oracle (oracle/gavel_lp.py, oracle/gavel_backend.py): closed-loop pin against the
reference's golden pickles, and the interior-point selection rule."""
import json
import os

import numpy as np

from oracle import gavel_backend as gb
from oracle import gavel_lp as gl

HERE = os.path.dirname(os.path.abspath(__file__))


def test_policy_oracle_pinned_against_golden_pickles():
    """The UNMODIFIED reference simulator driven by policies.py + the HiGHS oracle backend reproduces the golden
    pickles of scheduler/reproduce/pickles/tacc_32gpus/ (numbers recorded by tests/golden/make_policy_pins.py):
    exactly for max_min_fairness and gandiva_fair, within 1 % (makespan, avg JCT) for the others with the
    interior-point selection — and NOT with a simplex vertex of the same objective, which is why the selection
    rule is part of the parity contract."""
    pins = json.load(open(os.path.join(HERE, "golden", "tacc32_policy_pins.json")))
    for name in ("max_min_fairness", "gandiva_fair"):
        c, g = pins[name]["centre"], pins[name]["golden"]
        assert abs(c["makespan"] - g["makespan"]) <= 1e-9 * g["makespan"]
        assert abs(c["avg_jct"] - g["avg_jct"]) <= 1e-9 * g["avg_jct"]
        assert c["rounds"] == g["rounds"]
    for name in ("finish_time_fairness", "min_total_duration", "max_sum_throughput_perf", "allox"):
        c, g = pins[name]["centre"], pins[name]["golden"]
        assert abs(c["makespan"] - g["makespan"]) <= 0.01 * g["makespan"], name
        assert abs(c["avg_jct"] - g["avg_jct"]) <= 0.01 * g["avg_jct"], name
        assert abs(c["rounds"] - g["rounds"]) <= 2
    worse = 0
    for name in ("finish_time_fairness", "min_total_duration", "max_sum_throughput_perf"):
        v, c, g = pins[name]["vertex"], pins[name]["centre"], pins[name]["golden"]
        err = lambda r: max(abs(r["makespan"] / g["makespan"] - 1), abs(r["avg_jct"] / g["avg_jct"] - 1))
        worse += err(v) > 2.5 * err(c)
    assert worse == 3


def test_analytic_centre_satisfies_its_optimality_conditions():
    rng = np.random.default_rng(0)
    for trial in range(6):
        J = int(rng.integers(5, 200))
        sf = rng.choice([1.0, 2.0, 4.0], J)
        lo = rng.uniform(0.0, 0.6, J)
        lo[rng.integers(0, J)] = 1.0                      # one job pinned at x = 1
        N = float((sf * np.minimum(lo, 1)).sum() * rng.uniform(1.05, 3.0))
        for w_lo, w_x in ((1.0, 1.0), (2.0, 0.0)):
            x = gl.analytic_centre_box(lo, sf, N, w_lo, w_x)
            free = lo < 1.0
            lam = 1.0 / (N - (sf * x).sum())
            grad = w_x / x[free] - 1 / (1 - x[free]) + w_lo / (x[free] - lo[free]) - lam * sf[free]
            assert np.all(x[~free] == 1.0) and np.all(x[free] > lo[free]) and np.all(x[free] < 1)
            assert np.abs(grad).max() <= 1e-6 * lam * sf.max()
    sf = rng.choice([1.0, 2.0, 4.0], 40)
    C = 0.37 * sf.sum()
    x = gl.analytic_centre_tied(sf, C)
    assert abs((sf * x).sum() - C) <= 1e-9 * C
    nu = (1 / x - 1 / (1 - x)) / sf
    assert np.ptp(nu) <= 1e-6 * max(1.0, abs(nu).max())


def test_backend_selection_keeps_the_lp_objective():
    rng = np.random.default_rng(1)
    J, N = 40, 16.0
    thr = rng.uniform(0.5, 20.0, J)
    sf = rng.choice([1.0, 2.0, 4.0], J)
    n = rng.uniform(1e4, 1e6, J)
    for mode, kw in ((gb.POL_MTD, dict(n=n)), (gb.POL_MAXSUM, {}), (gb.POL_MAXMIN, {})):
        xv, ov, _ = gb.pooled_cpu(mode, N, thr, sf, select="vertex", **kw)
        xc, oc, _ = gb.pooled_cpu(mode, N, thr, sf, select="centre", **kw)
        assert abs(ov - oc) <= 1e-9 * abs(ov)
        assert (sf * xc).sum() <= N * (1 + 1e-9) and xc.min() >= 0 and xc.max() <= 1
        if mode == gb.POL_MAXSUM:
            assert abs((thr * xc).sum() - (thr * xv).sum()) <= 1e-7 * (thr * xv).sum()
        if mode == gb.POL_MTD:
            assert np.all(thr * xc >= n / oc * (1 - 1e-9))


def test_live_closed_loop_max_min_fairness_reproduces_golden():
    """Live version of the pin above for one policy (needs /root/reference; ~5 s): the unmodified reference
    simulator + policies.py + the HiGHS backend reproduce the max_min_fairness golden pickle exactly."""
    import pickle
    import pytest
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("reference tree not present")
    with gb.cpu_backend() as P:
        res = rh.simulate("max_min_fairness", policy_obj=P.get_policy("max_min_fairness", solver="ECOS", seed=0))
    gold = pickle.load(open(os.path.join(
        rh.REF, "reproduce/pickles/tacc_32gpus",
        "max_min_fairness_120_0.2_5_100_40_25_0,0.5,0.5_0.6,0.3,0.09,0.01_multigpu_dynamic_simulation.pickle"), "rb"))
    assert abs(res["makespan"] - gold["makespan"]) <= 1e-9 * gold["makespan"]
    assert abs(res["avg_jct"] - gold["avg_jct"]) <= 1e-9 * gold["avg_jct"]
    assert len(res["per_round_schedule"]) == len(gold["per_round_schedule"])
