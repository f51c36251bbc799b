"""Probe: tiny instances (8 jobs on 8 GPUs, gangs as wide as half the cluster) — device counts vs the exact MILP."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import shockwave_milp as om
from shockwave_b200 import Engine, make_params
from tests.synth import synth_problem
from tests import fixtures as fx
LOGV = om.pwl_log_values(fx.BASES, fx.ORIGIN)
eng = Engine(0)
J, G, T, D, k = 8, 8, 6, 120.0, 1e-3
for seed in range(6):
    pb = synth_problem(J, G, T, D, seed=seed, tight=1.0)
    prm = make_params(G, T, D, k, 12.0, 1.0, fx.BASES, fx.ORIGIN, round_ptr=pb["round_ptr"])
    out = eng.solve(prm, pb["g"], pb["E"], pb["c"], pb["dbar"], pb["rem"], pb["ftobj"])
    ora = om.dynamic_eisenberg_gale(pb["g"], pb["E"], pb["c"], pb["dbar"], pb["rem"], pb["ftobj"], G, T, D, pb["round_ptr"], k, 12.0, 1.0, fx.BASES, LOGV, rel_gap=1e-6, time_limit=60.0, do_rank=False)
    res = out["results"][0]
    no = ora["x"].sum(axis=1).astype(int)
    ev = om.evaluate(out["x"][0], pb["g"], pb["E"].astype(float), pb["c"].astype(float), pb["dbar"], pb["rem"], out["weights"][0], G, T, D, k, fx.BASES, LOGV)
    eo = om.evaluate(ora["x"], pb["g"], pb["E"].astype(float), pb["c"].astype(float), pb["dbar"], pb["rem"], ora["weights"], G, T, D, k, fx.BASES, LOGV)
    print("seed", seed, "status", res["status"], ora["status"], "passes", eng.last_timings()["passes"], "shortfall", res["shortfall"])
    print("  g   ", pb["g"]); print("  ours", out["nrounds"][0], "welf %.6f M %.3f obj %.6f" % (ev[1], ev[2], ev[0]))
    print("  ora ", no, "welf %.6f M %.3f obj %.6f" % (eo[1], eo[2], eo[0]))
    print("  load ours", (out["x"][0] * pb["g"][:, None]).sum(axis=0), "ora", (ora["x"] * pb["g"][:, None]).sum(axis=0).astype(int))
