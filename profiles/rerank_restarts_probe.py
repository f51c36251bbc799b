"""Quality / latency of the fallback re-rank search against the number of perturb-and-continue rounds
(SWB_OPT_RERANK_RESTARTS) on the 128 recorded fallback solves; yardstick: exact re-rank MILP (HiGHS, gap 1e-6).
Run on the GPU box: python profiles/rerank_restarts_probe.py > gpurun_out/rerank_restarts.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import shockwave_milp as om
from shockwave_b200 import Engine, make_params
from tests import fixtures as fx

eng = Engine(0)
T, G, D = fx.TACC["T"], fx.TACC["G"], fx.TACC["D"]
inst = []
for i in range(fx.n_solves()):
    s = fx.solve(i)
    if s["status"] == om.STATUS_FALLBACK:
        inst.append(s)
exact = {}
out = {}
for r in (0, 1, 2, 3, 5):
    eng.set_option(7, r)
    exc, ms = [], []
    for n, s in enumerate(inst):
        prm = make_params(G, T, D, fx.TACC["k"], fx.TACC["lam"], fx.TACC["rhomax"], fx.BASES, fx.ORIGIN, round_ptr=s["round_ptr"])
        o = eng.solve(prm, s["g"], s["E"], s["c"], s["dbar"], s["rem"], s["ftobj"], bfkey=s["rem"])
        x, w = o["x"][0], o["weights"][0]
        if n not in exact:
            y = om.rank_in_schedule(x.astype(float), w, s["g"].astype(np.int64), G, 1e-6, 30.0)
            exact[n] = om.rank_objective(y, w)
        exc.append((om.rank_objective(x, w) - exact[n]) / max(1e-12, abs(exact[n])))
        ms.append(eng.last_timings()["ms_place"])
    exc = np.array(exc)
    out[r] = dict(median=float(np.median(exc)), p90=float(np.percentile(exc, 90)), max=float(exc.max()),
                  above_1e3=int((exc > 1e-3).sum()), n=len(exc), place_ms_mean=float(np.mean(ms)), place_ms_max=float(np.max(ms)))
print(json.dumps(out, indent=1))
