"""CPU tests of the dense-relaxation oracle (oracle/market_lp.py) and of the numpy restatement of the device iteration
(tests/ref_market.py): the LP anchors on the homogeneous relaxation of oracle/shockwave_milp.py (itself pinned on the
reference's golden pickles), the restatement converges to that LP."""
import numpy as np
import pytest

from oracle import market_lp as ml
from oracle import shockwave_milp as om
from tests import fixtures as fx
from tests.ref_market import RefMarket, proj_budget
from tests.synth import synth_problem

LOGV = om.pwl_log_values(fx.BASES, fx.ORIGIN)


@pytest.mark.parametrize("J,G,T,k", [(40, 16, 12, 1e-3), (96, 32, 20, 1e-9), (64, 32, 16, 1e1)])
def test_w1_equals_homogeneous_relaxation(J, G, T, k):
    D = 120.0
    pb = synth_problem(J, G, T, D, seed=J, tight=3.0)
    rate = (D / pb["dbar"])[:, None]
    cap = np.full((1, T), float(G))
    sol = ml.solve(pb["g"], pb["E"], pb["c"], pb["dbar"], pb["rem"], rate, cap, k, fx.BASES, LOGV)
    lp = om.dynamic_eisenberg_gale(pb["g"], pb["E"], pb["c"], pb["dbar"], pb["rem"], 1e30 * np.ones(J), G, T, D,
                                   pb["round_ptr"], k, 12.0, 1.0, fx.BASES, LOGV, relax=True)
    assert lp["status"] == om.STATUS_FTF_FEASIBLE
    assert abs(sol["objective"] - lp["objective"]) <= 1e-7 * abs(lp["objective"])
    ev = ml.evaluate(sol["x"], pb["g"], pb["E"].astype(float), pb["c"].astype(float), pb["dbar"], pb["rem"], rate, cap,
                     k, fx.BASES, LOGV)
    assert abs(ev[0] - sol["objective"]) <= 1e-7 * abs(sol["objective"]) and ev[1] <= 1e-7 and ev[2] <= 1e-7


def test_heterogeneous_dominates_homogeneous_and_respects_rows():
    """More worker types can only help; budget rows sum_w x_jwt <= 1 and per-round capacities hold."""
    J, G, T, D, k = 60, 16, 12, 120.0, 1e-9
    pb = synth_problem(J, G, T, D, seed=4, tight=3.0)
    base = D / pb["dbar"]
    rate3 = np.stack([base, 0.6 * base, 0.35 * base], axis=1)
    cap3 = np.array([[G] * T, [G // 2] * T, [G // 4] * T], float)
    cap3[1, ::3] = 2.0                                            # capacities that differ between rounds
    s1 = ml.solve(pb["g"], pb["E"], pb["c"], pb["dbar"], pb["rem"], rate3[:, :1], cap3[:1], k, fx.BASES, LOGV)
    s3 = ml.solve(pb["g"], pb["E"], pb["c"], pb["dbar"], pb["rem"], rate3, cap3, k, fx.BASES, LOGV)
    assert s3["objective"] >= s1["objective"] - 1e-12
    x = s3["x"]
    assert x.sum(axis=1).max() <= 1 + 1e-7
    assert np.all((pb["g"][:, None, None] * x).sum(axis=0) <= cap3 * (1 + 1e-7))


def test_budget_projection_is_euclidean():
    rng = np.random.default_rng(0)
    y = rng.normal(0.3, 0.6, (50, 3, 8)).astype(np.float32)
    p = proj_budget(y)
    assert p.min() >= 0 and p.sum(axis=1).max() <= 1 + 1e-6
    # optimality: no feasible point of a random cloud is closer
    for _ in range(200):
        q = rng.dirichlet(np.ones(4), (50, 8))[:, :, :3].transpose(0, 2, 1) * rng.uniform(0, 1, (50, 1, 8))
        assert np.all(((y - p) ** 2).sum(axis=1) <= ((y - q) ** 2).sum(axis=1) + 1e-5)


@pytest.mark.parametrize("W,nonuniform", [(1, 0.0), (3, 0.3)])
def test_restatement_converges_to_the_lp(W, nonuniform):
    J, G, T, D, k = 96, 32, 16, 120.0, 1e-9
    pb = synth_problem(J, G, T, D, seed=11, tight=3.0)
    base = D / pb["dbar"]
    rate = np.stack([base * f for f in [1.0, 0.6, 0.35][:W]], axis=1)
    cap = np.repeat(np.array([G, G // 2, G // 4][:W], float)[:, None], T, axis=1)
    if nonuniform:
        cap = np.maximum(1.0, np.round(cap * (1 + nonuniform * np.random.default_rng(1).uniform(-1, 1, cap.shape))))
    lp = ml.solve(pb["g"], pb["E"], pb["c"], pb["dbar"], pb["rem"], rate, cap, k, fx.BASES, LOGV)["objective"]
    ref = RefMarket(pb["g"], pb["E"], pb["c"], pb["dbar"], pb["rem"], rate, cap, k, D, fx.BASES, LOGV, T)
    X, obj = ref.run(np.zeros((J, W, T), np.float32), 600, coarse_iters=800)
    ev = ml.evaluate(X, pb["g"], pb["E"].astype(float), pb["c"].astype(float), pb["dbar"], pb["rem"], rate, cap, k,
                     fx.BASES, LOGV)
    assert ev[1] <= 2e-5 and ev[2] <= 1e-5
    assert abs(ev[0] - obj[0]) <= 1e-5 * abs(ev[0])
    assert 0 <= (lp - obj[0]) / abs(lp) + 1e-6 < 5e-4
