"""Generator of tests/golden/scale64_oracle_pin.json: the UNMODIFIED reference simulator on the 220-job trace of
reproduce/scale_64gpus.sh (64 GPUs, configurations/scale_64gpus.json: k = 10, lambda = 5, wisr_throughputs.json) with
the HiGHS oracle in place of the Gurobi call (oracle/ref_harness.py).  The reference ships no golden pickle for this run;
the pin is what the closed-loop GPU test compares the kernels-in-the-loop run against (tests/test_closed_loop.py).
Needs /root/reference; about an hour of CPU.  python tests/golden/make_scale64_pin.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import ref_harness as rh  # noqa: E402

TRACE = "traces/reproduce/220_0.2_5_100_25_4_0,0.5,0.5_0.6,0.3,0.09,0.01_multigpu_dynamic.trace"

if __name__ == "__main__":
    t0 = time.time()
    rec = []
    cls = rh.make_oracle_scheduler_cls(record=rec)
    out = rh.simulate("shockwave", shockwave_scheduler_cls=cls, config="configurations/scale_64gpus.json",
                      cluster="64:0:0", trace=TRACE, throughputs="wisr_throughputs.json")
    ftf = np.array(out["finish_time_fairness_list"])
    pin = dict(trace=TRACE, cluster="64:0:0", config="configurations/scale_64gpus.json", throughputs="wisr_throughputs.json",
               makespan=out["makespan"], avg_jct=out["avg_jct"], cluster_util=out["cluster_util"],
               worst_ftf=float(ftf.max()), unfair_fraction=float((ftf > 1.05).mean()),
               rounds=len(out["per_round_schedule"]), jct_list=[float(v) for v in out["jct_list"]],
               resolves=len(rec), oracle="HiGHS mip_rel_gap 1e-3, time_limit 15 s (the reference's Gurobi settings)",
               seconds=time.time() - t0)
    json.dump(pin, open(os.path.join(ROOT, "tests", "golden", "scale64_oracle_pin.json"), "w"), indent=1)
    print({k: v for k, v in pin.items() if k != "jct_list"})
