"""Generates tests/golden/tacc32_solves.npz + tacc32_oracle_pin.json  (run in the build container,
where /root/reference exists):

    python tests/golden/make_solve_fixtures.py

It runs the UNMODIFIED reference simulator (scheduler/scheduler.py) on the canonical trace
(reproduce/tacc_32gpus.sh) with the Gurobi call replaced by the HiGHS oracle, through
oracle/ref_harness.py, and records (a) the end-to-end metrics next to the reference's golden pickle
and (b) every re-solve's inputs/outputs as seen at the ShockwaveScheduler boundary: the job profiles,
the per-solve host inputs (epoch progress, throughput-timeline summary, reestimate flag), the
reference-computed forecast (dbar, rem, ftobj) and the oracle's status / objective / x.
"""
import glob
import json
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rec = []
    cls = rh.make_oracle_scheduler_cls(record=rec)
    out = rh.simulate("shockwave", shockwave_scheduler_cls=cls)
    gold = pickle.load(open(glob.glob(
        "/root/reference/scheduler/reproduce/pickles/tacc_32gpus/shockwave_*")[0], "rb"))
    ftf, gftf = np.array(out["finish_time_fairness_list"]), np.array(gold["finish_time_fairness_list"])
    pin = dict(
        oracle=dict(makespan=out["makespan"], avg_jct=out["avg_jct"], cluster_util=out["cluster_util"],
                    rounds=len(out["per_round_schedule"]), worst_ftf=float(ftf.max()),
                    unfair_frac=float((ftf > 1.05).mean()), solves=len(rec)),
        golden=dict(makespan=gold["makespan"], avg_jct=gold["avg_jct"], cluster_util=gold["cluster_util"],
                    rounds=len(gold["per_round_schedule"]), worst_ftf=float(gftf.max()),
                    unfair_frac=float((gftf > 1.05).mean())),
        golden_file="scheduler/reproduce/pickles/tacc_32gpus/shockwave_120_..._simulation.pickle",
        note="HiGHS stand-in for Gurobi, mip_rel_gap=1e-3, time_limit=15 s")
    json.dump(pin, open(os.path.join(HERE, "tacc32_oracle_pin.json"), "w"), indent=1)
    st = cls.job_statics
    arrs = {}
    jids = sorted(st.keys())
    arrs["job_ids"] = np.array(jids, dtype=np.int32)
    arrs["job_nworkers"] = np.array([st[j]["nworkers"] for j in jids], dtype=np.int32)
    arrs["job_epochs"] = np.array([st[j]["epochs"] for j in jids], dtype=np.int32)
    arrs["job_nsamples"] = np.array([st[j]["epoch_nsamples"] for j in jids], dtype=np.float64)
    arrs["job_tsubmit"] = np.array([st[j]["timestamp_submit"] for j in jids], dtype=np.float64)
    arrs["job_grd"] = np.array([st[j]["grd"] for j in jids], dtype=np.float64)
    arrs["job_off"] = np.cumsum([0] + [st[j]["epochs"] for j in jids]).astype(np.int64)
    arrs["job_pre"] = np.concatenate([st[j]["pre"] for j in jids])
    arrs["job_bs"] = np.concatenate([st[j]["bs"] for j in jids]).astype(np.int32)
    arrs["n_solves"] = np.array(len(rec))
    for i, r in enumerate(rec):
        k = f"s{i:03d}_"
        arrs[k + "meta"] = np.array([r["round_ptr"], r["J"], r["status"], int(r["reestimate"])], dtype=np.int64)
        arrs[k + "scal"] = np.array([r["objective"], r["welfare"], r["makespan"]])
        for name in ("g", "E", "c", "dbar", "rem", "ftobj", "bfkey0", "x", "weights", "meas_ns", "meas_end"):
            arrs[k + name] = np.asarray(r[name])
        arrs[k + "rem_fb"] = np.asarray(r["rem_fb"] if r["rem_fb"] is not None else r["rem"])
        arrs[k + "jobids"] = np.array(r["jobids"], dtype=np.int32)
        arrs[k + "round0"] = np.array(r["round0"], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "tacc32_solves.npz"), **arrs)
    print(json.dumps(pin, indent=1))


if __name__ == "__main__":
    main()
