"""Closed-loop pin of the Gavel-policy oracle: the UNMODIFIED reference simulator (scheduler/scheduler.py) driven by
shockwave_b200/policies.py with its two device calls routed to the HiGHS oracle (oracle/gavel_backend.py), on the
canonical 120-job trace, against the golden pickles the reference ships (scheduler/reproduce/pickles/tacc_32gpus/).
Run here (needs /root/reference; no GPU):  python tests/golden/make_policy_pins.py  -> tests/golden/tacc32_policy_pins.json
"""
import json
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh                    # noqa: E402
from oracle import gavel_backend as gb                  # noqa: E402

POLICIES = ["max_min_fairness", "finish_time_fairness", "min_total_duration", "max_sum_throughput_perf", "allox",
            "gandiva_fair"]
SUFFIX = "_120_0.2_5_100_40_25_0,0.5,0.5_0.6,0.3,0.09,0.01_multigpu_dynamic_simulation.pickle"


def summary(r):
    ftf = np.asarray(r["finish_time_fairness_list"], dtype=float)
    return dict(makespan=float(r["makespan"]), avg_jct=float(r["avg_jct"]), cluster_util=float(r["cluster_util"]),
                rounds=len(r["per_round_schedule"]), worst_ftf=float(ftf.max()), unfair_frac=float((ftf > 1.05).mean()))


def main():
    out = {}
    for select in ("centre", "vertex"):
        for name in POLICIES:
            if select == "vertex" and name in ("allox", "gandiva_fair", "max_min_fairness"):
                continue        # no LP degeneracy involved / unique optimum on this trace
            with gb.cpu_backend() as P:
                if select == "vertex":
                    P._pooled = lambda *a, **k: gb.pooled_cpu(*a, select="vertex", **k)
                pol = P.get_policy(name, solver="ECOS", seed=0)
                res = rh.simulate(name, policy_obj=pol)
            gold = pickle.load(open(os.path.join(rh.REF, "reproduce/pickles/tacc_32gpus", name + SUFFIX), "rb"))
            out.setdefault(name, {"golden": summary(gold)})[select] = summary(res)
            print(name, select, out[name][select], "golden", out[name]["golden"], flush=True)
    out["_note"] = ("centre = interior-point selection (analytic centre of the optimal face), vertex = HiGHS vertex; "
                    "same LP objective, different closed loop")
    json.dump(out, open(os.path.join(ROOT, "tests/golden/tacc32_policy_pins.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
