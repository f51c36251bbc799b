cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
S=/usr/local/cuda/bin/compute-sanitizer
cat > /tmp/race.py <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np
from shockwave_b200 import policies as P
rng=np.random.default_rng(3)
J=40; a=rng.uniform(0.5,10,(J,1))*rng.uniform(0.1,1,(J,3)); sf=rng.choice([1.0,2.0,4.0],J)
print(P._hetero(P.POL_MAXMIN,[8.0,4.0,4.0],a,sf)[1], P._hetero(P.POL_MAXSUM,[8.0,4.0,4.0],a,sf)[1])
print(P._hetero(P.POL_MAXMIN,[8.0,4.0],a[:,:2],sf)[1], P._hetero(P.POL_MTD,[8.0,4.0,4.0],a,sf,n=rng.uniform(1e4,1e6,J))[1])
PY
timeout 1200 $S --tool racecheck --error-exitcode 7 --print-limit 20 python /tmp/race.py 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_policies_hetero.py -m gpu -q -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
