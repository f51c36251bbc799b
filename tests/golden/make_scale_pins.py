"""Generator of tests/golden/scale{64,128,256}_oracle_pin.json: the UNMODIFIED reference simulator on the traces of
reproduce/scale_{64,128,256}gpus.sh (220 / 460 / 900 jobs, configurations/scale_*gpus.json, wisr_throughputs.json) with
the HiGHS oracle in place of the Gurobi call (oracle/ref_harness.py).  The reference ships no golden pickles for these
runs; the pins are what the closed-loop GPU test compares the kernels-in-the-loop runs against
(tests/test_closed_loop.py).  Needs /root/reference.  python tests/golden/make_scale_pins.py 64|128|256
(64: 80 s of CPU; the larger ones up to hours — every MILP may use the reference's 15 s time limit)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import ref_harness as rh  # noqa: E402

TRACES = {64: "traces/reproduce/220_0.2_5_100_25_4_0,0.5,0.5_0.6,0.3,0.09,0.01_multigpu_dynamic.trace",
          128: "traces/reproduce/460_0.2_5_100_10_1_0,0.5,0.5_0.6,0.3,0.09,0.01_multigpu_dynamic.trace",
          256: "traces/reproduce/900_0.2_5_1000_5_15_0,0.5,0.5_0.6,0.3,0.09,0.01_multigpu_dynamic.trace"}

if __name__ == "__main__":
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    TRACE = TRACES[G]
    t0 = time.time()
    rec = []
    cls = rh.make_oracle_scheduler_cls(record=rec)
    out = rh.simulate("shockwave", shockwave_scheduler_cls=cls, config=f"configurations/scale_{G}gpus.json",
                      cluster=f"{G}:0:0", trace=TRACE, throughputs="wisr_throughputs.json")
    ftf = np.array(out["finish_time_fairness_list"])
    pin = dict(trace=TRACE, cluster=f"{G}:0:0", config=f"configurations/scale_{G}gpus.json", throughputs="wisr_throughputs.json",
               makespan=float(out["makespan"]), avg_jct=float(out["avg_jct"]), cluster_util=float(out["cluster_util"]),
               worst_ftf=float(ftf.max()), unfair_fraction=float((ftf > 1.05).mean()),
               rounds=len(out["per_round_schedule"]), jct_list=[float(v) for v in out["jct_list"]],
               resolves=len(rec), oracle="HiGHS mip_rel_gap 1e-3, time_limit 15 s (the reference's Gurobi settings)",
               seconds=time.time() - t0)
    json.dump(pin, open(os.path.join(ROOT, "tests", "golden", f"scale{G}_oracle_pin.json"), "w"), indent=1)
    print({k: v for k, v in pin.items() if k != "jct_list"})
