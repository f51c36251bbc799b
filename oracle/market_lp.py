"""CPU ORACLE (test infrastructure only — never imported by the product path).

HiGHS LP of the heterogeneous *relaxation* of Shockwave's schedule program over the dense allocation tensor
x[j][w][t] (job x worker type x planning round).  Pieces and where they come from in the reference:

* per-round, per-type capacity  sum_j g_j x_jwt <= G_wt   — scheduler/shockwave.py:303-319 generalised by the
  Gavel base constraints scheduler/policies/policy.py:58-65 (one row per worker type, `sum_w x_jw <= 1` per job);
* progress  P_j <= sum_wt r_jw x_jwt  (epochs; r_jw = D / dbar_j on the reference type, shockwave.py:371-377) and
  P_j <= E_j - c_j;
* utility u_j = (c_j + P_j)/E_j and its piece-wise linear log  plog(u_j) = sum_b lam_jb log(base_b) with
  sum_b lam_jb base_b = u_j, sum_b lam_jb = 1 (shockwave.py:379-402; the SOS2 binaries :403-419 are redundant for a
  concave maximise, tests/test_oracle_milp.py::test_sos2_binaries_are_redundant);
* objective  sum_j plog_j / (J T) - k max_j max(0, R_j - dbar_j P_j)   (shockwave.py:555-568).

With W = 1 and r_j = D / dbar_j this is `dynamic_eisenberg_gale(relax=True)` of oracle/shockwave_milp.py (checked in
tests/test_oracle_market_lp.py); that function is itself pinned end to end on the reference's golden pickles.  For
W > 1 the reference has no counterpart ("we assume homogeneous hardware"): parity unpinned beyond that W = 1 anchor.
"""
import numpy as np
import scipy.sparse as sp
from scipy.optimize import linprog


def solve(g, E, c, dbar, R, rate, cap, k, bases, logv, weights=None):
    """rate [J][W] epochs per round, cap [W][T] GPUs.  Returns dict(objective, x [J][W][T], P [J], M, price [W][T])."""
    g = np.asarray(g, float)
    E = np.asarray(E, float)
    c = np.asarray(c, float)
    dbar = np.asarray(dbar, float)
    R = np.asarray(R, float)
    rate = np.asarray(rate, float)
    cap = np.asarray(cap, float)
    J, W = rate.shape
    T = cap.shape[1]
    B = len(bases)
    wj = np.ones(J) if weights is None else np.asarray(weights, float)
    nx = J * W * T
    oP, oL, oM = nx, nx + J, nx + J + J * B
    nv = oM + 1
    cost = np.zeros(nv)
    cost[oL:oM] = -(wj[:, None] * np.asarray(logv, float)[None, :]).ravel() / (J * T)
    cost[oM] = k
    lo = np.zeros(nv)
    hi = np.full(nv, np.inf)
    hi[:nx] = 1.0
    hi[oP:oL] = np.maximum(E - c, 0.0)
    jj, ww, tt = np.meshgrid(np.arange(J), np.arange(W), np.arange(T), indexing="ij")
    xi = (jj * W + ww) * T + tt
    rows, cols, vals, ub = [], [], [], []
    nr = 0
    # capacity per (w, t)
    rows.append((ww * T + tt).ravel()); cols.append(xi.ravel()); vals.append(g[jj].ravel())
    ub.append(cap.ravel()); nr += W * T
    # budget per (j, t): sum_w x <= 1
    if W > 1:
        rows.append(nr + (jj * T + tt).ravel()); cols.append(xi.ravel()); vals.append(np.ones(nx))
        ub.append(np.ones(J * T)); nr += J * T
    # P_j - sum r x <= 0
    rows.append(nr + jj.ravel()); cols.append(xi.ravel()); vals.append(-rate[jj, ww].ravel())
    rows.append(nr + np.arange(J)); cols.append(oP + np.arange(J)); vals.append(np.ones(J))
    ub.append(np.zeros(J)); nr += J
    # R_j - dbar_j P_j - M <= 0
    rows.append(nr + np.arange(J)); cols.append(oP + np.arange(J)); vals.append(-dbar)
    rows.append(nr + np.arange(J)); cols.append(np.full(J, oM)); vals.append(-np.ones(J))
    ub.append(-R); nr += J
    A_ub = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nr, nv))
    b_ub = np.concatenate(ub)
    # equalities: sum_b lam base_b - P/E = c/E ; sum_b lam = 1
    jb = np.repeat(np.arange(J), B)
    er = np.concatenate([jb, np.arange(J), J + jb])
    ec = np.concatenate([oL + np.arange(J * B), oP + np.arange(J), oL + np.arange(J * B)])
    ev = np.concatenate([np.tile(np.asarray(bases, float), J), -1.0 / E, np.ones(J * B)])
    A_eq = sp.csr_matrix((ev, (er, ec)), shape=(2 * J, nv))
    b_eq = np.concatenate([c / E, np.ones(J)])
    res = linprog(cost, A_ub=A_ub, b_ub=b_ub, A_eq=A_eq, b_eq=b_eq, bounds=np.stack([lo, hi], axis=1),
                  method="highs", options=dict(primal_feasibility_tolerance=1e-9, dual_feasibility_tolerance=1e-9))
    if res.status != 0:
        raise RuntimeError(f"HiGHS: {res.message}")
    x = res.x[:nx].reshape(J, W, T)
    return dict(objective=-res.fun, x=x, P=res.x[oP:oL], M=res.x[oM],
                price=-res.ineqlin.marginals[:W * T].reshape(W, T))


def evaluate(x, g, E, c, dbar, R, rate, cap, k, bases, logv, weights=None):
    """Objective of a dense allocation and its worst relative capacity / budget violations."""
    from oracle.shockwave_milp import plog
    x = np.asarray(x, float)
    J, W, T = x.shape
    P = np.minimum((np.asarray(rate, float)[:, :, None] * x).sum(axis=(1, 2)), np.asarray(E, float) - c)
    u = (c + P) / E
    wj = np.ones(J) if weights is None else np.asarray(weights, float)
    welfare = float((wj * np.array([plog(v, bases, logv) for v in u])).sum() / (J * T))
    M = float(np.maximum(0.0, R - dbar * P).max())
    load = (np.asarray(g, float)[:, None, None] * x).sum(axis=0)
    return welfare - k * M, float((load / cap - 1.0).max()), float(x.sum(axis=1).max() - 1.0)
