"""CPU tests of the HiGHS oracle itself (oracle/shockwave_milp.py) and its pinning."""
import json
import os

import numpy as np
import pytest

from oracle import shockwave_milp as om
from tests import fixtures as fx
from tests.synth import synth_problem

LOGV = om.pwl_log_values(fx.BASES, fx.ORIGIN)
HERE = os.path.dirname(os.path.abspath(__file__))


def _args(pb, G, T, D=120.0, k=1e-3):
    return (pb["g"], pb["E"], pb["c"], pb["dbar"], pb["rem"], pb["ftobj"], G, T, D, pb["round_ptr"], k, 12.0, 1.0,
            fx.BASES, LOGV)


def test_pinned_against_reference_golden_pickle():
    """End-to-end pin: the UNMODIFIED reference simulator driven by this oracle reproduces the golden
    pickle (scheduler/reproduce/pickles/tacc_32gpus/shockwave_*.pickle) within the 3 % spread of the
    three Shockwave pickles the reference ships.  Numbers were recorded by
    tests/golden/make_solve_fixtures.py."""
    pin = json.load(open(os.path.join(HERE, "golden", "tacc32_oracle_pin.json")))
    o, g = pin["oracle"], pin["golden"]
    for key in ("makespan", "avg_jct", "cluster_util"):
        assert abs(o[key] - g[key]) / g[key] < 0.03, (key, o[key], g[key])
    assert abs(o["rounds"] - g["rounds"]) <= 6
    assert abs(o["worst_ftf"] - g["worst_ftf"]) / g["worst_ftf"] < 0.10
    assert abs(o["unfair_frac"] - g["unfair_frac"]) < 0.03


def test_sos2_binaries_are_redundant():
    """The reference's z_{j,b} SOS2 rows (shockwave.py:403-419) do not change the optimum of a concave
    maximise: same objective with and without them."""
    for seed in range(3):
        pb = synth_problem(10, 8, 6, seed=seed, tight=3.0)
        a = om.dynamic_eisenberg_gale(*_args(pb, 8, 6), rel_gap=1e-9, with_sos2=False, do_rank=False)
        b = om.dynamic_eisenberg_gale(*_args(pb, 8, 6), rel_gap=1e-9, with_sos2=True, do_rank=False)
        assert a["status"] == b["status"]
        assert abs(a["objective"] - b["objective"]) <= 1e-7 * abs(a["objective"])


def test_solution_is_feasible_and_objective_consistent():
    pb = synth_problem(40, 16, 12, seed=5, tight=3.0)
    r = om.dynamic_eisenberg_gale(*_args(pb, 16, 12), rel_gap=1e-6)
    ev = om.evaluate(r["x"], pb["g"], pb["E"].astype(float), pb["c"].astype(float), pb["dbar"], pb["rem"],
                     r["weights"], 16, 12, 120.0, 1e-3, fx.BASES, LOGV)
    assert ev[4]
    assert abs(ev[0] - r["objective"]) <= 1e-6 * abs(r["objective"])
    lp = om.dynamic_eisenberg_gale(*_args(pb, 16, 12), relax=True)
    assert lp["objective"] >= r["objective"] - 1e-9


def test_fallback_path_and_rank():
    pb = synth_problem(24, 16, 10, seed=2, tight=0.3)       # FTF rows infeasible
    r = om.dynamic_eisenberg_gale(*_args(pb, 16, 10), rel_gap=1e-6, do_rank=False)
    assert r["status"] == om.STATUS_FALLBACK
    prio = r["weights"]
    assert prio.max() > 1.0
    y = om.rank_in_schedule(r["x"], prio, pb["g"].astype(np.int64), 16, 1e-6, 15.0)
    assert np.array_equal(y.sum(axis=1), r["x"].sum(axis=1))           # counts kept
    assert np.all(y.T @ pb["g"] <= 16)
    assert om.rank_objective(y, prio) <= om.rank_objective(r["x"], prio) + 1e-9


def test_recorded_solves_are_reproducible():
    """Re-solving recorded inputs gives the recorded verdict and an objective within the MIP gap."""
    for i in (0, 17, 60, 97, 128):
        s = fx.solve(i)
        r = om.dynamic_eisenberg_gale(s["g"], s["E"], s["c"], s["dbar"], s["rem"], s["ftobj"], fx.TACC["G"],
                                      fx.TACC["T"], fx.TACC["D"], s["round_ptr"], fx.TACC["k"], fx.TACC["lam"],
                                      fx.TACC["rhomax"], fx.BASES, LOGV, rel_gap=1e-3, do_rank=False)
        assert r["status"] == s["status"]
        if s["status"] == om.STATUS_FTF_FEASIBLE:
            assert abs(r["objective"] - s["objective"]) <= 2e-3 * abs(s["objective"]) + 1e-9


def test_construct_schedules_backfill_order():
    x = np.array([[1, 0], [0, 0], [0, 1], [0, 0]])
    g = np.array([2, 1, 2, 1])
    R = np.array([5.0, 9.0, 1.0, 9.0])
    s = om.construct_schedules(x, [10, 11, 12, 13], g, R, 7, 4)
    assert s[7] == [10, 11, 13]      # back-fill by descending R, stable
    assert s[8] == [12, 11, 13]


def test_momentumed_average_matches_reference_formula():
    series = [(3, 100.0), (5, 200.0), (9, 50.0)]
    got = om.finish_time_momentumed_average(list(series), 11)
    want = 0.9 * (2 / 8 * 100 + 4 / 8 * 200 + 2 / 8 * 50) + 0.1 * 50
    assert abs(got - want) < 1e-12
    assert om.finish_time_momentumed_average([(4, 77.0)], 4) == 0.9 * 77.0 + 0.1 * 77.0


def test_live_pin_is_what_the_current_oracle_produces():
    """The pin above is a recorded number; this re-runs the unmodified reference simulator with the CURRENT oracle
    (needs /root/reference, ~1 min) and demands the recorded metrics exactly — HiGHS is deterministic, so any drift
    means the oracle's model changed (even a row reordering moves the closed loop by a few per cent)."""
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("reference tree not present")
    pin = json.load(open(os.path.join(HERE, "golden", "tacc32_oracle_pin.json")))["oracle"]
    out = rh.simulate("shockwave", shockwave_scheduler_cls=rh.make_oracle_scheduler_cls())
    assert abs(out["makespan"] - pin["makespan"]) <= 1e-9 * pin["makespan"], (out["makespan"], pin["makespan"])
    assert abs(out["avg_jct"] - pin["avg_jct"]) <= 1e-9 * pin["avg_jct"]
    assert len(out["per_round_schedule"]) == pin["rounds"]


def test_product_placement_rule_closed_loop_pin():
    """P5 as far as it can be taken without a GPU next to the reference: the unmodified reference simulator with HiGHS
    for the round counts and the PRODUCT's placement rule (tests/ref_placement.py, which place.cu matches bit for bit)
    for the rounds themselves stays inside the 3 % spread of the reference's own Shockwave pickles
    (tests/golden/make_placement_pin.py)."""
    pin = json.load(open(os.path.join(HERE, "golden", "tacc32_placement_pin.json")))
    g = pin["golden"]
    for variant in ("product_placement", "product_placement_tight_counts"):
        o = pin[variant]
        for key in ("makespan", "avg_jct", "cluster_util"):
            assert abs(o[key] - g[key]) / g[key] < 0.03, (variant, key, o[key], g[key])
        assert abs(o["rounds"] - g["rounds"]) <= 6
        assert abs(o["worst_ftf"] - g["worst_ftf"]) / g["worst_ftf"] < 0.10
        assert abs(o["unfair_frac"] - g["unfair_frac"]) < 0.03
    t = pin["product_placement_tight_counts"]       # near-optimal counts (what the GPU delivers): within 0.5 %
    assert abs(t["makespan"] - g["makespan"]) / g["makespan"] < 0.005 and abs(t["avg_jct"] - g["avg_jct"]) / g["avg_jct"] < 0.005


def test_live_placement_pin_is_what_the_current_rule_produces():
    """Live version of the placement pin (needs /root/reference, ~1 min): guards tacc32_placement_pin.json against going
    stale when tests/ref_placement.py (i.e. place.cu) changes."""
    from oracle import ref_harness as rh
    from tests.ref_placement import place
    if not rh.reference_available():
        pytest.skip("reference tree not present")
    pin = json.load(open(os.path.join(HERE, "golden", "tacc32_placement_pin.json")))["product_placement"]
    out = rh.simulate("shockwave", shockwave_scheduler_cls=rh.make_oracle_scheduler_cls(placement=place))
    assert abs(out["makespan"] - pin["makespan"]) <= 1e-9 * pin["makespan"], (out["makespan"], pin["makespan"])
    assert abs(out["avg_jct"] - pin["avg_jct"]) <= 1e-9 * pin["avg_jct"]
