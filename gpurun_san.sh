cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
S=/usr/local/cuda/bin/compute-sanitizer
echo "== memcheck: edge cases + small solves + policies"
timeout 900 $S --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_solve.py -m gpu -q -x -k "edge_cases or packed_masks or batched" 2>&1 | tail -15
echo rc=$?
timeout 900 $S --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_policies_hetero.py tests/test_gpu_policies.py -m gpu -q -x -k "hetero-7 or hetero-40 or hetero-60 or closed_forms or allox or (analytic_centre and 12-32)" 2>&1 | tail -15
echo "== racecheck: hetero small, pooled policy, one canonical-size solve"
cat > /tmp/race.py <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np
from shockwave_b200 import policies as P, Engine, make_params
from tests import fixtures as fx
rng=np.random.default_rng(3)
J=40; a=rng.uniform(0.5,10,(J,1))*rng.uniform(0.1,1,(J,3)); sf=rng.choice([1.0,2.0,4.0],J)
print(P._hetero(P.POL_MAXMIN,[8.0,4.0,4.0],a,sf)[1], P._hetero(P.POL_MAXSUM,[8.0,4.0,4.0],a,sf)[1])
print(P._pooled(P.POL_MTD,16.0,a[:,2],sf,n=rng.uniform(1e4,1e6,J))[1], P._pooled(P.POL_MAXSUM,16.0,a[:,2],sf)[1])
eng=P._engine()
for i in (3, 60):
    s=fx.solve(i)
    prm=make_params(32,20,120.0,1e-3,12.0,1.0,fx.BASES,fx.ORIGIN,round_ptr=s["round_ptr"])
    out=eng.solve(prm,s["g"],s["E"],s["c"],s["dbar"],s["rem"],s["ftobj"],bfkey=s["rem"])
    print(i, out["results"][0]["objective"], out["results"][0]["placement"])
PY
timeout 1200 $S --tool racecheck --error-exitcode 7 --print-limit 20 python /tmp/race.py 2>&1 | tail -25
