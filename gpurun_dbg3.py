import sys; sys.path.insert(0,'.')
import numpy as np
from oracle import shockwave_milp as om
from shockwave_b200 import Engine, make_params
from tests import fixtures as fx
from tests.synth import synth_problem
LOGV=om.pwl_log_values(fx.BASES,fx.ORIGIN)
eng=Engine(0)
rng=np.random.default_rng(2024)
for it in range(70):
    J=int(rng.integers(6,90)); G=int(rng.choice([8,12,16,24,32,48,64])); T=int(rng.integers(4,25))
    k=float(rng.choice([1e-6,1e-3,1e-1,1e1,1e5])); tight=float(rng.choice([0.4,1.0,3.0])); D=float(rng.choice([60.0,120.0,360.0]))
    pb=synth_problem(J,G,T,D,seed=1000+it,tight=tight)
    lam=float(rng.choice([5.0,12.0,15.0]))
    prm=make_params(G,T,D,k,lam,1.0,fx.BASES,fx.ORIGIN,round_ptr=pb["round_ptr"])
    out=eng.solve(prm,pb["g"],pb["E"],pb["c"],pb["dbar"],pb["rem"],pb["ftobj"])
    res,x,w=out["results"][0],out["x"][0],out["weights"][0]
    ora=om.dynamic_eisenberg_gale(pb["g"],pb["E"],pb["c"],pb["dbar"],pb["rem"],pb["ftobj"],G,T,D,pb["round_ptr"],k,lam,1.0,fx.BASES,LOGV,rel_gap=1e-6,time_limit=30.0,do_rank=False)
    ev=om.evaluate(x,pb["g"],pb["E"].astype(float),pb["c"].astype(float),pb["dbar"],pb["rem"],w,G,T,D,k,fx.BASES,LOGV)
    eo=om.evaluate(ora["x"],pb["g"],pb["E"].astype(float),pb["c"].astype(float),pb["dbar"],pb["rem"],w,G,T,D,k,fx.BASES,LOGV)
    gap=(eo[0]-ev[0])/max(1e-12,abs(eo[0]))
    if gap>2e-4:
        used=(x.T.astype(int)@pb["g"]); uo=(ora["x"].T.astype(int)@pb["g"])
        print(it,'J',J,'G',G,'T',T,'k',k,'tight',tight,'st',res["status"],'gap %.2e'%gap,'short',res["shortfall"],'welf',ev[1],eo[1],'M',ev[2],eo[2],'usedGPUrounds',used.sum(),uo.sum(),'GT',G*T,'gmax',pb["g"].max(), 'relaxed',res["relaxed_objective"])
