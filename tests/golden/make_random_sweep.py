"""Oracle results for a wide randomised parity sweep of the market solve, computed HERE on CPU (HiGHS, gap 1e-6) so that
the GPU test (tests/test_gpu_random_sweep.py::test_recorded_random_sweep) only has to run the kernels and compare.
Instances are regenerated from their seeds by tests/synth.py; only the oracle's verdict and objective are stored.
  python tests/golden/make_random_sweep.py [n_instances]  ->  tests/golden/random_sweep_oracle.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import shockwave_milp as om          # noqa: E402
from tests import fixtures as fx                 # noqa: E402
from tests.synth import synth_problem            # noqa: E402

LOGV = om.pwl_log_values(fx.BASES, fx.ORIGIN)


def instance(it):
    rng = np.random.default_rng(50_000 + it)
    J = int(rng.integers(4, 320))
    G = int(rng.choice([8, 12, 16, 24, 32, 48, 64, 96, 128]))
    T = int(rng.integers(3, 33))
    k = float(rng.choice([1e-6, 1e-3, 1e-1, 1e1, 1e5]))
    lam = float(rng.choice([5.0, 12.0, 15.0]))
    tight = float(rng.choice([0.3, 0.6, 1.0, 2.0, 4.0]))
    D = float(rng.choice([60.0, 120.0, 360.0]))
    return dict(it=it, J=J, G=G, T=T, k=k, lam=lam, tight=tight, D=D, seed=70_000 + it)


def main(n):
    out = []
    t0 = time.time()
    for it in range(n):
        p = instance(it)
        pb = synth_problem(p["J"], p["G"], p["T"], p["D"], seed=p["seed"], tight=p["tight"])
        ora = om.dynamic_eisenberg_gale(pb["g"], pb["E"], pb["c"], pb["dbar"], pb["rem"], pb["ftobj"], p["G"], p["T"],
                                        p["D"], pb["round_ptr"], p["k"], p["lam"], 1.0, fx.BASES, LOGV, rel_gap=1e-6,
                                        time_limit=120.0, do_rank=False)
        w = ora["weights"]
        obj = om.evaluate(ora["x"], pb["g"], pb["E"].astype(float), pb["c"].astype(float), pb["dbar"], pb["rem"], w,
                          p["G"], p["T"], p["D"], p["k"], fx.BASES, LOGV)[0]
        p.update(status=int(ora["status"]), objective=float(obj))
        out.append(p)
        print(it, p["J"], p["G"], p["T"], "status", p["status"], "obj", p["objective"], "%.0fs" % (time.time() - t0), flush=True)
        json.dump(out, open(os.path.join(ROOT, "tests/golden/random_sweep_oracle.json"), "w"), indent=0)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 240)
