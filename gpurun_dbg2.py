import sys; sys.path.insert(0,'.')
import numpy as np
from oracle import shockwave_milp as om
from shockwave_b200 import Engine, make_params
from tests import fixtures as fx
eng=Engine(0)
G,T,D=32,20,120.0
s=fx.solve(97)
prm=make_params(G,T,D,1e-3,12.0,1.0,fx.BASES,fx.ORIGIN,round_ptr=s["round_ptr"])
out=eng.solve(prm,s["g"],s["E"],s["c"],s["dbar"],s["rem"],s["ftobj"],bfkey=s["rem"])
print(out["results"][0])
print('g     ',s["g"].tolist())
print('gpu n ',out["nrounds"][0].tolist())
print('ora n ',s["x"].sum(axis=1).astype(int).tolist())
nfin=np.ceil(s["dbar"]*(s["E"]-s["c"])/D-1e-9).astype(int)
print('nfin  ',nfin.tolist())
print('rem/D ',np.round(s["rem"]/D,2).tolist())
print('w     ',['%.1e'%v for v in out["weights"][0]])
print('load gpu',(out["x"][0].T.astype(int)@s["g"]).tolist())
print('load ora',(s["x"].T.astype(int)@s["g"]).tolist())
