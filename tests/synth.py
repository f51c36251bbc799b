"""Seeded synthetic solver inputs shaped like the reference's workloads:
scale factors {1,2,4,8} w.p. (0.6,0.3,0.09,0.01) (scripts/utils/generate_trace.py:406-418), epoch counts
and per-epoch durations spanning the small/medium/large/XL split (generate_trace.py:371-403)."""
import numpy as np


def synth_problem(J, G, T, D=120.0, seed=0, round_ptr=None, tight=1.0):
    rng = np.random.default_rng(seed)
    g = rng.choice([1, 2, 4, 8], size=J, p=[0.6, 0.3, 0.09, 0.01]).astype(np.int32)
    g = np.minimum(g, G).astype(np.int32)
    cls = rng.choice(4, size=J, p=[0.72, 0.2, 0.05, 0.03])
    hours = np.array([0.6, 3.0, 12.0, 40.0])[cls] * rng.uniform(0.3, 1.7, size=J)
    E = rng.integers(4, 120, size=J).astype(np.int32)
    dur = hours * 3600.0 / E                      # mean seconds per epoch
    c = np.floor(rng.random(J) * 0.8 * E).astype(np.int32)
    c[rng.random(J) < 0.25] = 0                   # fresh arrivals: first PWL segment (slope ~61)
    dbar = dur * rng.uniform(0.85, 1.15, size=J)
    rem = dur * (E - c) * rng.uniform(0.9, 1.1, size=J)
    r = int(rng.integers(0, 150)) if round_ptr is None else int(round_ptr)
    share = min(1.0, G / J)
    # finish-time objective around the fair-share finish time; `tight` < 1 makes FTF rows infeasible
    ftobj = (D * r + (rem + D * T * rng.uniform(0.0, 1.5, size=J)) / share) * rng.uniform(0.9, 1.6, size=J) * tight
    return dict(g=g, E=E, c=c, dbar=dbar, rem=rem, ftobj=ftobj, round_ptr=r)
