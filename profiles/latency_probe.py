"""Single re-solve latency probe (S = 1, inputs resident) at BASELINE config D and at the canonical size:
solve / place kernel times from the library's own CUDA events, with the cluster path on and off."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from shockwave_b200 import Engine, make_params
from tests.synth import synth_problem

BASES, ORIGIN = [0.0, 0.2, 0.4, 0.6, 0.8, 1.0], {0.0: 1e-6}
eng = Engine(0)
dev = torch.device("cuda", 0)
rows = []
for (J, G, T) in ((4096, 512, 64), (1024, 128, 32), (109, 32, 20)):
    for k in (1e-3, 1e1, 1e5):
        for tight in (3.0, 0.5):
            pb = synth_problem(J, G, T, 120.0, seed=7, tight=tight)
            prm = make_params(G, T, 120.0, k, 12.0, 1.0, BASES, ORIGIN, round_ptr=pb["round_ptr"])
            t = {n: torch.from_numpy(np.ascontiguousarray(pb[n].astype(dt))).to(dev)
                 for n, dt in (("g", np.int32), ("E", np.int32), ("c", np.int32), ("dbar", np.float64),
                               ("rem", np.float64), ("ftobj", np.float64))}
            ptrs = {n: v.data_ptr() for n, v in t.items()}
            xm = torch.zeros((1, J, 2), dtype=torch.int64, device=dev)
            bm = torch.zeros((1, J, 2), dtype=torch.int64, device=dev)
            outp = dict(xmask=xm.data_ptr(), bfmask=bm.data_ptr())
            for cl in (1, 8):
                eng.set_option(2, cl)
                wall, ks, kp = [], [], []
                for i in range(12):
                    t0 = time.perf_counter()
                    res = eng.solve_device([prm], J, ptrs, outp)
                    wall.append((time.perf_counter() - t0) * 1e3)
                    tm = eng.last_timings()
                    ks.append(tm["ms_solve"]); kp.append(tm["ms_place"])
                rows.append(dict(J=J, G=G, T=T, k=k, tight=tight, cluster=cl, status=res[0]["status"],
                                 evals=res[0]["mu_iters"], m_evals=res[0]["m_evals"], passes=tm["passes"],
                                 wall_ms=float(np.median(wall[4:])), solve_ms=float(np.median(ks[4:])),
                                 place_ms=float(np.median(kp[4:]))))
                print(rows[-1], flush=True)
eng.set_option(2, 8)
json.dump(rows, open("gpurun_out/latency_probe.json", "w"), indent=1)
