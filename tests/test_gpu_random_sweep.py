"""Randomised parity sweep of the market solve against the live HiGHS oracle (gap 1e-6) over many small
and medium instances: all widths, cluster sizes that are not powers of two, every shipped k, FTF-feasible
and fallback cases.  P1 (feasibility) and P3 (verdict) must hold on every instance; P2 (objective within
the reference's own 1e-3 MIPGap) on every instance of realistic size (J >= 16) — tiny instances with gangs
as wide as the cluster are pure integer knapsacks where the exact MILP can win by a few 1e-3."""
import numpy as np
import pytest

from oracle import shockwave_milp as om
from shockwave_b200 import make_params
from tests import fixtures as fx
from tests.synth import synth_problem

pytestmark = pytest.mark.gpu
LOGV = om.pwl_log_values(fx.BASES, fx.ORIGIN)


def test_random_sweep(engine):
    rng = np.random.default_rng(2024)
    gaps, sizes = [], []
    for it in range(70):
        J = int(rng.integers(6, 90))
        G = int(rng.choice([8, 12, 16, 24, 32, 48, 64]))
        T = int(rng.integers(4, 25))
        k = float(rng.choice([1e-6, 1e-3, 1e-1, 1e1, 1e5]))
        tight = float(rng.choice([0.4, 1.0, 3.0]))
        D = float(rng.choice([60.0, 120.0, 360.0]))
        pb = synth_problem(J, G, T, D, seed=1000 + it, tight=tight)
        prm = make_params(G, T, D, k, float(rng.choice([5.0, 12.0, 15.0])), 1.0, fx.BASES, fx.ORIGIN,
                          round_ptr=pb["round_ptr"])
        out = engine.solve(prm, pb["g"], pb["E"], pb["c"], pb["dbar"], pb["rem"], pb["ftobj"])
        ora = om.dynamic_eisenberg_gale(pb["g"], pb["E"], pb["c"], pb["dbar"], pb["rem"], pb["ftobj"], G, T, D,
                                        pb["round_ptr"], k, prm.lam, 1.0, fx.BASES, LOGV, rel_gap=1e-6,
                                        time_limit=30.0, do_rank=False)
        res, x, w = out["results"][0], out["x"][0], out["weights"][0]
        assert res["status"] == ora["status"], (it, J, G, T)
        obj, _, _, n, cap_ok = om.evaluate(x, pb["g"], pb["E"].astype(float), pb["c"].astype(float), pb["dbar"],
                                           pb["rem"], w, G, T, D, k, fx.BASES, LOGV)
        assert cap_ok
        assert abs(obj - res["objective"]) <= 1e-9 * max(1.0, abs(obj))
        ora_obj = om.evaluate(ora["x"], pb["g"], pb["E"].astype(float), pb["c"].astype(float), pb["dbar"],
                              pb["rem"], w, G, T, D, k, fx.BASES, LOGV)[0]
        gaps.append((ora_obj - obj) / max(1e-12, abs(ora_obj)))
        sizes.append(J)
    gaps, sizes = np.array(gaps), np.array(sizes)
    print("instances", len(gaps), "worst gap", gaps.max(), "p95", np.quantile(gaps, 0.95),
          "share within 1e-3", (gaps <= 1e-3).mean(), "worst gap at J>=16", gaps[sizes >= 16].max())
    assert gaps[sizes >= 16].max() <= 1e-3
    assert gaps.max() <= 1e-2


def test_recorded_random_sweep(engine):
    """240 random instances up to J = 320, T = 32, G = 128 whose oracle verdicts / objectives were computed offline
    (tests/golden/make_random_sweep.py, HiGHS gap 1e-6): P1 + P3 on every instance, P2 (1e-3) on every instance with
    J >= 16."""
    import json
    import os
    recs = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "random_sweep_oracle.json")))
    gaps, sizes, verdict_miss = [], [], []
    for p in recs:
        pb = synth_problem(p["J"], p["G"], p["T"], p["D"], seed=p["seed"], tight=p["tight"])
        prm = make_params(p["G"], p["T"], p["D"], p["k"], p["lam"], 1.0, fx.BASES, fx.ORIGIN, round_ptr=pb["round_ptr"])
        out = engine.solve(prm, pb["g"], pb["E"], pb["c"], pb["dbar"], pb["rem"], pb["ftobj"])
        res, x, w = out["results"][0], out["x"][0], out["weights"][0]
        if res["status"] != p["status"]:
            verdict_miss.append(p["it"])
            continue
        obj, _, _, n, cap_ok = om.evaluate(x, pb["g"], pb["E"].astype(float), pb["c"].astype(float), pb["dbar"],
                                           pb["rem"], w, p["G"], p["T"], p["D"], p["k"], fx.BASES, LOGV)
        assert cap_ok, p
        assert abs(obj - res["objective"]) <= 1e-9 * max(1.0, abs(obj)), p
        gaps.append((p["objective"] - obj) / max(1e-12, abs(p["objective"])))
        sizes.append(p["J"])
    gaps, sizes = np.array(gaps), np.array(sizes)
    print("recorded sweep:", len(recs), "instances, verdict mismatches", verdict_miss, "worst gap", gaps.max(),
          "p95", np.quantile(gaps, 0.95), "worst at J>=16", gaps[sizes >= 16].max(),
          "share within 1e-3", (gaps <= 1e-3).mean())
    assert not verdict_miss
    assert gaps[sizes >= 16].max() <= 1e-3
    assert gaps.max() <= 1e-2
