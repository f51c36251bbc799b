"""Small driver for compute-sanitizer (memcheck / racecheck) over the kernels touched in round 2's second half: the dense
market iteration (all W, generic and compile-time kernels, two levels) and the fallback re-rank search (cluster
multi-start + perturb-and-continue rounds).  compute-sanitizer --tool memcheck python profiles/sanitize_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from shockwave_b200 import Engine, make_params
from shockwave_b200.engine import market_pgd
from tests import fixtures as fx
from tests.synth import synth_problem

eng = Engine(0)
D = 120.0
for (J, G, T, W) in [(64, 32, 20, 1), (130, 64, 32, 3), (97, 32, 64, 2), (70, 64, 8, 4), (50, 16, 128, 1)]:
    pb = synth_problem(J, G, T, D, seed=J, tight=3.0)
    rate = np.stack([D / pb["dbar"] * f for f in [1.0, 0.6, 0.35, 0.2][:W]], axis=1).astype(np.float32)
    cap = np.repeat(np.array([G, G // 2, G // 4, G // 4][:W], float)[:, None], T, axis=1)
    prm = [make_params(G, T, D, k, 12.0, 1.0, fx.BASES, fx.ORIGIN, round_ptr=pb["round_ptr"]) for k in (1e-9, 1e1)]
    X = np.zeros((2, J, W, T), dtype=np.float32)
    obj, _ = market_pgd(eng, prm, pb["g"], pb["E"], pb["c"], pb["dbar"], pb["rem"], rate, None, X, 5, coarse_iters=6, cap=cap)
    print("market", J, G, T, W, obj[:, 0])
for seed in range(3):                     # canonical size, fallback path: cluster multi-start re-rank search
    pb = synth_problem(109, 32, 20, D, seed=5 + seed, tight=0.5)
    prm = make_params(32, 20, D, 1e-3, 12.0, 1.0, fx.BASES, fx.ORIGIN, round_ptr=pb["round_ptr"])
    out = eng.solve(prm, pb["g"], pb["E"], pb["c"], pb["dbar"], pb["rem"], pb["ftobj"], packed=True)
    print("rerank", out["results"][0]["status"], out["results"][0]["flags"] >> 8)
eng.set_option(2, 1)                      # single-CTA placement with the search
pb = synth_problem(60, 32, 20, D, seed=9, tight=0.5)
prm = make_params(32, 20, D, 1e-3, 12.0, 1.0, fx.BASES, fx.ORIGIN, round_ptr=pb["round_ptr"])
out = eng.solve(prm, pb["g"], pb["E"], pb["c"], pb["dbar"], pb["rem"], pb["ftobj"], packed=True)
print("rerank one CTA", out["results"][0]["status"], out["results"][0]["flags"] >> 8)
print("done")
