"""Driver for the ncu captures of round 2: launches every kernel of the hot path a few times on BASELINE-size inputs
(config D for the Shockwave kernels, config E for the Gavel ones) so that `ncu -k regex:<kernel> -s 2 -c 1` finds a
warmed-up launch.  Not a benchmark: numbers printed by a run under ncu are never bench values."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from shockwave_b200 import Engine, make_params
from shockwave_b200.engine import market_pgd
from tests.synth import synth_problem

BASES, ORIGIN = [0.0, 0.2, 0.4, 0.6, 0.8, 1.0], {0.0: 1e-6}
J, G, T, D = 4096, 512, 64, 120.0
eng = Engine(0)
dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
REP = 4


def dev_scn(S, tight):
    pbs = [synth_problem(J, G, T, D, seed=1000 + s, tight=tight(s)) for s in range(S)]
    st = lambda k, dt: torch.from_numpy(np.ascontiguousarray(np.stack([p[k] for p in pbs]).astype(dt))).to(dev)
    t = dict(g=st("g", np.int32), E=st("E", np.int32), c=st("c", np.int32), dbar=st("dbar", np.float64),
             rem=st("rem", np.float64), ftobj=st("ftobj", np.float64))
    prm = [make_params(G, T, D, [1e-3, 1e1, 1e5][s % 3], 12.0, 1.0, BASES, ORIGIN, round_ptr=pbs[s]["round_ptr"]) for s in range(S)]
    return pbs, t, prm


if which in ("all", "single"):          # S = 1: 8-CTA cluster solve + place (plain and fallback)
    for tight in (3.0, 0.5):
        pbs, t, prm = dev_scn(1, lambda s: tight)
        xm = torch.zeros((1, J, 2), dtype=torch.int64, device=dev); bm = torch.zeros_like(xm)
        for _ in range(REP):
            eng.solve_device(prm, J, {k: v.data_ptr() for k, v in t.items()}, dict(xmask=xm.data_ptr(), bfmask=bm.data_ptr()))
if which in ("all", "batched"):         # S = 296: one CTA per scenario
    S = 296
    pbs, t, prm = dev_scn(S, lambda s: 3.0 if s % 4 else 0.5)
    xm = torch.zeros((S, J, 2), dtype=torch.int64, device=dev); bm = torch.zeros_like(xm)
    for _ in range(REP):
        eng.solve_device(prm, J, {k: v.data_ptr() for k, v in t.items()}, dict(xmask=xm.data_ptr(), bfmask=bm.data_ptr()))
if which in ("all", "gbm"):
    pb = synth_problem(J, G, T, D, seed=1000, tight=3.0)
    rng = np.random.default_rng(0)
    for _ in range(REP):
        eng.gbm_forecast(pb["rem"], np.minimum(pb["E"] - pb["c"], 256), rng.uniform(-1e-3, 1e-3, J), rng.uniform(0, 0.05, J), 8192, 0, 7)
if which in ("all", "market"):
    pb = synth_problem(J, G, T, D, seed=1000, tight=3.0)
    Sd = 512
    Xd = torch.zeros((Sd, J, 1, T), dtype=torch.float32, device=dev)
    dj = {k: torch.from_numpy(np.ascontiguousarray(pb[k]).astype(np.float64 if k != "g" else np.int32)).to(dev) for k in ("g", "E", "c", "dbar", "rem")}
    drate = torch.from_numpy((D / pb["dbar"])[:, None].astype(np.float32)).to(dev)
    mptr = dict(shape=(Sd, J, 1, T), per_scenario_jobs=0, g=dj["g"].data_ptr(), E=dj["E"].data_ptr(), c=dj["c"].data_ptr(),
                dbar=dj["dbar"].data_ptr(), rem=dj["rem"].data_ptr(), rate=drate.data_ptr(), X=Xd.data_ptr())
    mprm = [make_params(G, T, D, 1e-9, 12.0, 1.0, BASES, ORIGIN) for _ in range(Sd)]
    market_pgd(eng, mprm, None, None, None, None, None, None, [G], None, 6, coarse_iters=100, device_ptrs=mptr)
if which in ("all", "gavel"):           # config E: 2048 jobs, 3 worker types
    rng = np.random.default_rng(3)
    JE, W = 2048, 3
    alloc = np.round(rng.uniform(0, 1, (JE, W)), 3); jt = rng.uniform(0, 500, (JE, W)); wt = rng.uniform(1e3, 5e3, W)
    thr = rng.uniform(0.1, 10, (JE, W)); dfc = np.zeros((JE, W)); sf = rng.choice([1, 2, 4, 8], JE, p=[0.6, 0.3, 0.09, 0.01])
    cap = [256, 128, 128]
    lists = [list(range(0, 256)), list(range(256, 384)), list(range(384, 512))]
    for _ in range(REP):
        eng.gavel_round(alloc, jt, wt, thr, dfc, sf, cap, [0, 1, 2], lists, {})
if which in ("all", "canonical"):       # canonical size, fallback path: cluster multi-start re-rank search
    pb = synth_problem(109, 32, 20, D, seed=5, tight=0.5)
    prm = make_params(32, 20, D, 1e-3, 12.0, 1.0, BASES, ORIGIN, round_ptr=pb["round_ptr"])
    for _ in range(REP):
        eng.solve(prm, pb["g"], pb["E"], pb["c"], pb["dbar"], pb["rem"], pb["ftobj"], packed=True)
print("done", which)
