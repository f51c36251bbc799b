import sys, time; sys.path.insert(0,'.')
import numpy as np
from shockwave_b200 import Engine, make_params
from tests import fixtures as fx
eng=Engine(0)
G,T,D=32,20,120.0
for i in (60,97,5):
    s=fx.solve(i)
    prm=make_params(G,T,D,1e-3,12.0,1.0,fx.BASES,fx.ORIGIN,round_ptr=s["round_ptr"])
    for rep in range(5):
        t=time.perf_counter(); out=eng.solve(prm,s["g"],s["E"],s["c"],s["dbar"],s["rem"],s["ftobj"]); dt=time.perf_counter()-t
    tm=eng.last_timings()
    print('canonical solve',i,'J',s["J"],'e2e ms',dt*1e3,'kernels',tm, out["results"][0]["m_evals"], out["results"][0]["mu_iters"])
