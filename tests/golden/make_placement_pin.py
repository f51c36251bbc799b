"""Closed-loop check of the PRODUCT's placement rule (P5 of the parity definition, as far as it can be taken without a
GPU next to the reference): the unmodified reference simulator, HiGHS for the round counts, and tests/ref_placement.py
— the numpy restatement of place.cu that the GPU kernel matches bit for bit (tests/test_gpu_placement_ref.py) — for
which rounds each job gets (water-filling, round order, fallback priority sweep) in place of Gurobi's own x and of the
rank_in_schedule_jobs MILP.   python tests/golden/make_placement_pin.py -> tests/golden/tacc32_placement_pin.json
"""
import glob
import json
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh          # noqa: E402
from tests.ref_placement import place         # noqa: E402


def summary(r):
    ftf = np.asarray(r["finish_time_fairness_list"], dtype=float)
    return dict(makespan=float(r["makespan"]), avg_jct=float(r["avg_jct"]), cluster_util=float(r["cluster_util"]),
                rounds=len(r["per_round_schedule"]), worst_ftf=float(ftf.max()), unfair_frac=float((ftf > 1.05).mean()))


def main():
    out = rh.simulate("shockwave", shockwave_scheduler_cls=rh.make_oracle_scheduler_cls(placement=place))
    gold = pickle.load(open(glob.glob("/root/reference/scheduler/reproduce/pickles/tacc_32gpus/shockwave_*")[0], "rb"))
    tight = rh.simulate("shockwave", shockwave_scheduler_cls=rh.make_oracle_scheduler_cls(placement=place, rel_gap=1e-6,
                                                                                       time_limit=60.0))
    pin = dict(product_placement=summary(out), product_placement_tight_counts=summary(tight), golden=summary(gold),
               note="HiGHS round counts + the product's placement rule (tests/ref_placement.py), same harness as "
                    "tacc32_oracle_pin.json; product_placement: counts at the reference's mip_rel_gap 1e-3; "
                    "product_placement_tight_counts: counts at gap 1e-6 (time limit 60 s), which is what the GPU solve "
                    "delivers (objective within 3e-6 of the optimum on these states)")
    json.dump(pin, open(os.path.join(ROOT, "tests/golden/tacc32_placement_pin.json"), "w"), indent=1)
    print(json.dumps(pin, indent=1))


if __name__ == "__main__":
    main()
