"""place.cu against its plain numpy restatement (tests/ref_placement.py): given the round counts the kernel placed,
the x matrix (water-filling, round order, fallback priority sweep with back-off) and the back-fill matrix must be
identical bit for bit — on every recorded re-solve of the canonical simulation (128 of 129 take the fallback path) and
on synthetic non-fallback instances."""
import numpy as np
import pytest

from shockwave_b200 import make_params
from tests import fixtures as fx
from tests.ref_placement import place
from tests.synth import synth_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _sweep_only(engine):
    """the numpy restatement covers the constructive part of place.cu (sweep, packer, round order, back-fill); the
    local search that follows on the fallback path (rerank.cuh) has its own test against the exact re-rank MILP"""
    engine.set_option(6, 0)          # SWB_OPT_RERANK_ITERS
    yield
    engine.set_option(6, 400)


def _compare(out, g, G, T, bfkey, tag):
    res, x, bf, n, w = out["results"][0], out["x"][0], out["backfill"][0], out["nrounds"][0], out["weights"][0]
    if res["shortfall"] != 0:
        return None
    ref = place(n, g, G, T, bfkey, fallback=(res["status"] == 1), w=w)
    if ref["shortfall"] != 0:
        return None          # the kernel's improvement pass added rounds beyond the plan: not restated
    same_x = np.array_equal(ref["x"], x.astype(bool))
    same_bf = np.array_equal(ref["backfill"], bf.astype(bool))
    assert ref["swept_rounds"] == res["placement"], (tag, ref["swept_rounds"], res["placement"])
    assert same_x, (tag, "x differs in", int((ref["x"] != x.astype(bool)).sum()), "cells")
    assert same_bf, (tag, "back-fill differs in", int((ref["backfill"] != bf.astype(bool)).sum()), "cells")
    return True


def test_recorded_canonical_solves_bit_for_bit(engine):
    T, G, D = fx.TACC["T"], fx.TACC["G"], fx.TACC["D"]
    done = 0
    for i in range(fx.n_solves()):
        s = fx.solve(i)
        prm = make_params(G, T, D, fx.TACC["k"], fx.TACC["lam"], fx.TACC["rhomax"], fx.BASES, fx.ORIGIN,
                          round_ptr=s["round_ptr"])
        out = engine.solve(prm, s["g"], s["E"], s["c"], s["dbar"], s["rem"], s["ftobj"], bfkey=s["rem"])
        done += 1 if _compare(out, s["g"], G, T, s["rem"], ("canonical", i)) else 0
    print("placements compared bit for bit:", done, "of", fx.n_solves())
    assert done >= fx.n_solves() - 10


@pytest.mark.parametrize("J,G,T,tight", [(40, 32, 20, 3.0), (200, 64, 32, 3.0), (300, 96, 24, 1.0), (1024, 128, 32, 3.0),
                                         (4096, 512, 64, 3.0), (150, 24, 16, 0.5)])
def test_synthetic_bit_for_bit(engine, J, G, T, tight):
    done = 0
    for seed in range(2):
        pb = synth_problem(J, G, T, 120.0, seed=300 + seed, tight=tight)
        prm = make_params(G, T, 120.0, 1e-3, 12.0, 1.0, fx.BASES, fx.ORIGIN, round_ptr=pb["round_ptr"])
        out = engine.solve(prm, pb["g"], pb["E"], pb["c"], pb["dbar"], pb["rem"], pb["ftobj"])
        done += 1 if _compare(out, pb["g"], G, T, pb["rem"], (J, G, T, seed)) else 0
    assert done >= 1
