"""Dense volatile-Fisher-market relaxation of Shockwave's schedule program (market.cu behind swb_market_pgd).

The reference plans on ONE worker type with the same capacity in every round ("we assume homogeneous hardware",
scripts/drivers/simulate_scheduler_with_trace.py:76-78; capacity row shockwave.py:303-319).  This module is the general
form BASELINE.json's north_star names: an allocation tensor x[job][worker type][round] with per-(job, type) progress
rates and per-(type, round) capacities (workers joining / leaving / reserved inside the planning window), solved by
primal-dual price-response iterations on the device.  `ShockwaveScheduler.heterogeneous_plan()` feeds it from the
drop-in scheduler's live state; nothing here falls back to a CPU solver.
"""
import numpy as np

from . import engine as _eng


def solve_relaxation(eng, params, g, E, c, dbar, rem, speed, cap, full_iters=400, coarse_iters=1000,
                     primal_weight=0.0, x0=None):
    """Fractional plan of J jobs on W worker types over T rounds.

    speed [W] or [J][W]: progress on type w relative to the type `dbar` was measured on (epochs per round =
    speed * D / dbar_j); cap [W][T] workers of type w available in round t.
    Returns dict(x [J][W][T] fp32 feasible, objective, makespan, progress [J] epochs, share [J][W] = mean over rounds).
    """
    g = np.ascontiguousarray(g, dtype=np.int32)
    J = len(g)
    cap = np.ascontiguousarray(cap, dtype=np.float64)
    W, T = cap.shape
    speed = np.asarray(speed, dtype=np.float64)
    if speed.ndim == 1:
        speed = np.broadcast_to(speed[None, :], (J, W))
    D = float(params.round_duration)
    rate = np.ascontiguousarray(speed * (D / np.asarray(dbar, float))[:, None], dtype=np.float32)
    X = np.zeros((1, J, W, T), dtype=np.float32) if x0 is None else np.ascontiguousarray(x0, np.float32).reshape(1, J, W, T)
    obj, ms = _eng.market_pgd(eng, params, g, E, c, dbar, rem, rate, None, X, full_iters, coarse_iters=coarse_iters,
                              primal_weight=primal_weight, cap=cap, warm_start=x0 is not None)
    x = X[0]
    progress = np.minimum((rate[:, :, None] * x).sum(axis=(1, 2)), np.asarray(E, float) - np.asarray(c, float))
    return dict(x=x, objective=float(obj[0, 0]), makespan=float(obj[0, 1]), progress=progress, share=x.mean(axis=2),
                dense_pass_ms=ms)
