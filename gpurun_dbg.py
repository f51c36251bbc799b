import sys, time; sys.path.insert(0,'.')
import numpy as np
from oracle import shockwave_milp as om
from shockwave_b200 import Engine, make_params
from tests import fixtures as fx
from tests.synth import synth_problem
LOGV=om.pwl_log_values(fx.BASES,fx.ORIGIN)
eng=Engine(0)
G,T,D=32,20,120.0
bad=0
for i in range(fx.n_solves()):
    s=fx.solve(i)
    prm=make_params(G,T,D,1e-3,12.0,1.0,fx.BASES,fx.ORIGIN,round_ptr=s["round_ptr"])
    out=eng.solve(prm,s["g"],s["E"],s["c"],s["dbar"],s["rem"],s["ftobj"],bfkey=s["rem"])
    res=out["results"][0]; x=out["x"][0]; w=out["weights"][0]
    ev=om.evaluate(x,s["g"],s["E"].astype(float),s["c"].astype(float),s["dbar"],s["rem"],w,G,T,D,1e-3,fx.BASES,LOGV)
    ora=om.evaluate(s["x"],s["g"],s["E"].astype(float),s["c"].astype(float),s["dbar"],s["rem"],w,G,T,D,1e-3,fx.BASES,LOGV)[0]
    d=abs(ev[0]-res["objective"])
    gap=(ora-ev[0])/max(1e-300,abs(ora))
    if d>1e-9 or gap>1e-3 or res["shortfall"]>0:
        bad+=1
        if bad<=12: print(i,'J',s["J"],'st',res["status"],'obj_gpu',res["objective"],'eval',ev[0],'welf',res["welfare"],ev[1],'M',res["makespan"],ev[2],'short',res["shortfall"],'oracle',ora,'gap',gap, 'wmax',w.max())
print('bad',bad)
# timing at config D
J,G,T=4096,512,64
pb=synth_problem(J,G,T,D,seed=7,tight=3.0)
prm=make_params(G,T,D,1e-3,12.0,1.0,fx.BASES,fx.ORIGIN,round_ptr=pb["round_ptr"])
for rep in range(3):
    t=time.perf_counter(); out=eng.solve(prm,pb["g"],pb["E"],pb["c"],pb["dbar"],pb["rem"],pb["ftobj"]); dt=time.perf_counter()-t
    print('D-config e2e solve ms',dt*1e3,out["results"][0])
pb=synth_problem(J,G,T,D,seed=8,tight=0.5)
t=time.perf_counter(); out=eng.solve(prm,pb["g"],pb["E"],pb["c"],pb["dbar"],pb["rem"],pb["ftobj"]); dt=time.perf_counter()-t
print('D-config fallback e2e solve ms',dt*1e3,out["results"][0])
