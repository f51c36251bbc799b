"""Loader for tests/golden/tacc32_solves.npz (made by tests/golden/make_solve_fixtures.py)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BASES = [0.0, 0.2, 0.4, 0.6, 0.8, 1.0]
ORIGIN = {0.0: 1e-6}
# configurations/tacc_32gpus.json + reproduce/tacc_32gpus.sh
TACC = dict(G=32, T=20, D=120.0, k=1e-3, lam=12.0, rhomax=1.0)

_cache = {}


def load():
    if "z" not in _cache:
        _cache["z"] = np.load(os.path.join(HERE, "golden", "tacc32_solves.npz"))
    return _cache["z"]


def n_solves():
    return int(load()["n_solves"])


def solve(i):
    z = load()
    k = f"s{i:03d}_"
    meta, scal = z[k + "meta"], z[k + "scal"]
    d = dict(round_ptr=int(meta[0]), J=int(meta[1]), status=int(meta[2]), reestimate=bool(meta[3]),
             objective=float(scal[0]), welfare=float(scal[1]), makespan=float(scal[2]))
    for name in ("g", "E", "c", "dbar", "rem", "ftobj", "bfkey0", "x", "weights", "meas_ns", "meas_end",
                 "rem_fb", "jobids", "round0"):
        d[name] = z[k + name]
    return d


def job_statics():
    z = load()
    out = {}
    off = z["job_off"]
    for i, jid in enumerate(z["job_ids"]):
        out[int(jid)] = dict(nworkers=int(z["job_nworkers"][i]), epochs=int(z["job_epochs"][i]),
                             epoch_nsamples=float(z["job_nsamples"][i]), timestamp_submit=float(z["job_tsubmit"][i]),
                             grd=float(z["job_grd"][i]), pre=z["job_pre"][off[i]:off[i + 1]],
                             bs=z["job_bs"][off[i]:off[i + 1]])
    return out
