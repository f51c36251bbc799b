"""Launches the Gavel-policy kernels once (for ncu): heterogeneous max-min at J=2048 x 3 worker types (hetero_kernel),
pooled min-total-duration at J=2048 (policy_kernel incl. the analytic-centre selection), AlloX assignment."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shockwave_b200 import policies as P        # noqa: E402

rng = np.random.default_rng(7)
J = 2048
m = rng.uniform(0.5, 20.0, size=(J, 1)) * np.sort(rng.uniform(0.1, 1.0, size=(J, 3)), axis=1)
sf = rng.choice([1.0, 2.0, 4.0, 8.0], J, p=[0.6, 0.3, 0.09, 0.01])
x, obj, rc = P._hetero(P.POL_MAXMIN, [256.0, 128.0, 128.0], m * (sf / (m @ np.array([0.5, 0.25, 0.25])))[:, None], sf)
print("hetero max-min", obj, P._hetero.last_stats)
x, obj, rc = P._pooled(P.POL_MTD, 512.0, m[:, 2], sf, n=rng.uniform(1e4, 1e6, J))
print("pooled MTD", obj)
x, obj, rc = P._pooled(P.POL_MAXSUM, 512.0, m[:, 2], sf)
print("pooled max-sum", obj)
