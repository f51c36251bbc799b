"""Per-instance quality of the fallback re-rank on the recorded canonical solves: GPU (sweep + local search) and sweep
alone vs the exact re-rank MILP (HiGHS) on the same counts.  Writes gpurun_out/rerank_probe.json."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import shockwave_milp as om
from shockwave_b200 import Engine, make_params
from tests import fixtures as fx

eng = Engine(0)
T, G, D = fx.TACC["T"], fx.TACC["G"], fx.TACC["D"]
rows = []
for i in range(fx.n_solves()):
    s = fx.solve(i)
    if s["status"] != om.STATUS_FALLBACK:
        continue
    prm = make_params(G, T, D, fx.TACC["k"], fx.TACC["lam"], fx.TACC["rhomax"], fx.BASES, fx.ORIGIN, round_ptr=s["round_ptr"])
    out = eng.solve(prm, s["g"], s["E"], s["c"], s["dbar"], s["rem"], s["ftobj"], bfkey=s["rem"])
    x, w = out["x"][0], out["weights"][0]
    tm = eng.last_timings()
    y = om.rank_in_schedule(x.astype(float), w, s["g"].astype(np.int64), G, 1e-6, 30.0)
    ry, rg = om.rank_objective(y, w), om.rank_objective(x, w)
    eng.set_option(6, 0)
    o0 = eng.solve(prm, s["g"], s["E"], s["c"], s["dbar"], s["rem"], s["ftobj"], bfkey=s["rem"])
    tm0 = eng.last_timings()
    eng.set_option(6, 400)
    r0 = om.rank_objective(o0["x"][0], w)
    rows.append(dict(i=i, J=int(s["J"]), milp=ry, exc=(rg - ry) / max(1e-12, abs(ry)), exc_sweep=(r0 - ry) / max(1e-12, abs(ry)),
                     cycles=(out["results"][0]["flags"] >> 8) & 0xfff, winner=out["results"][0]["flags"] >> 20, swept=out["results"][0]["placement"],
                     place_ms=tm["ms_place"], place_ms_sweep=tm0["ms_place"],
                     widths=sorted(set(int(v) for v in s["g"][x.sum(axis=1) > 0]))))
    if rows[-1]["exc"] > 1e-3:
        print(rows[-1], flush=True)
e = np.array([r["exc"] for r in rows]); e0 = np.array([r["exc_sweep"] for r in rows])
print("LS: median %.2e p90 %.2e max %.2e >1e-3: %d/%d | sweep: median %.2e p90 %.2e max %.2e >1e-3: %d | place ms mean %.3f vs %.3f"
      % (np.median(e), np.percentile(e, 90), e.max(), (e > 1e-3).sum(), len(e), np.median(e0), np.percentile(e0, 90), e0.max(),
         (e0 > 1e-3).sum(), np.mean([r["place_ms"] for r in rows]), np.mean([r["place_ms_sweep"] for r in rows])))
json.dump(rows, open("gpurun_out/rerank_probe.json", "w"), indent=1)
