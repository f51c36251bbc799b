"""Plain numpy restatement of the dense market iteration (market.cu + the orchestration of swb_market_pgd in api.cu) —
the same operation sequence (fp32 tensor pass, fp64 dual pass), used as the numerics oracle of that floating-point
kernel.  The LP it converges to is pinned separately against HiGHS (oracle/market_lp.py)."""
import numpy as np


def plog(u, bases, logv):
    return np.interp(u, bases, logv)


def proj_budget(y):
    """Euclidean projection of every (job, round) W-vector on {x >= 0, sum_w x <= 1}; y [J][W][T] fp32."""
    J, W, T = y.shape
    if W == 1:
        return np.clip(y, 0.0, 1.0).astype(np.float32)
    pos = np.maximum(y, 0.0).sum(axis=1)
    out = np.maximum(y, 0.0).astype(np.float32)
    bad = pos > 1.0
    if bad.any():
        v = np.moveaxis(y, 1, 2)[bad].astype(np.float32)          # [n][W]
        act = np.ones_like(v, bool)
        th = np.zeros(len(v), np.float32)
        for _ in range(W):
            n = act.sum(axis=1).astype(np.float32)
            th = ((np.where(act, v, 0.0).sum(axis=1) - np.float32(1.0)) / n).astype(np.float32)
            act &= v > th[:, None]
        res = np.maximum(v - th[:, None], 0.0).astype(np.float32)
        tmp = np.moveaxis(out, 1, 2).copy()
        tmp[bad] = res
        out = np.moveaxis(tmp, 2, 1)
    return np.ascontiguousarray(out, dtype=np.float32)


class Level:
    """One level (full or time-coarsened) of the iteration on a single scenario."""

    def __init__(self, g, E, c, dbar, rem, rate, cap, k, D, bases, logv, Tfull, rscale, pw):
        self.g = np.asarray(g, np.float32); self.E = np.asarray(E, float); self.c = np.asarray(c, float)
        self.dbar = np.asarray(dbar, float); self.rem = np.asarray(rem, float)
        self.rate = np.asarray(rate, np.float32)
        self.icap = (1.0 / np.asarray(cap, float)).astype(np.float32)           # [W][T]
        self.k, self.D, self.bases, self.logv = k, D, np.asarray(bases, float), np.asarray(logv, float)
        self.Tfull, self.rs, self.pw = float(Tfull), float(rscale), np.float32(pw)
        self.pw0 = np.float32(pw)
        self.ipw = np.float32(1.0) / self.pw
        J, W = self.rate.shape
        self.T = self.icap.shape[1]
        self.mj = np.zeros(J); self.om = np.zeros(J); self.pi = np.zeros((W, self.T))

    # ---- dense pass, MODE 1 / MODE 0 (fp32)
    def scale(self, X, cs):
        Y = np.clip(X * cs[None], 0.0, 1.0).astype(np.float32)
        self._reduce(Y)
        return Y

    def step(self, X):
        g, r = self.g[:, None, None], self.rate[:, :, None]
        rb = (self.rate * self.beta[:, None])[:, :, None]
        den = (g * self.icap[None] + rb).max(axis=1, keepdims=True).astype(np.float32)
        tau = (self.ipw / den).astype(np.float32)
        gain = (self.theta[:, None] * self.rate)[:, :, None]
        Y = (X + tau * (gain - self.price[None] * g)).astype(np.float32)
        Y = proj_budget(Y)
        self._reduce(Y)
        return Y

    def _reduce(self, Y):
        self.rowp = (self.rate[:, :, None] * Y).sum(axis=(1, 2)).astype(np.float32)
        self.colload = (self.g[:, None, None] * Y).sum(axis=0).astype(np.float32)

    # ---- dual pass (fp64)
    def score(self, load):
        J = len(self.E)
        P = np.minimum(self.rs * self.rowp.astype(float), self.E - self.c)
        welfare = plog((self.c + P) / self.E, self.bases, self.logv).sum() / (J * self.Tfull)
        mx = np.maximum(0.0, self.rem - self.dbar * P).max()
        viol = float((load.astype(float) * self.icap - 1.0).max())
        return welfare - self.k * mx, mx, viol

    def dual(self, phase):
        J, T = len(self.E), self.T
        DT = self.D * self.Tfull
        kk = self.k * DT
        aE, fD = self.rs / self.E, self.rs * self.dbar / DT
        sumr = self.rate.astype(float).sum(axis=1)
        obj = self.score(self.colload)
        pw = None
        if phase == 0:
            self.beta = (aE + fD).astype(np.float32)
            mkmax = (self.rem / DT).max()
            self.pw = np.float32(max(float(self.pw0), kk * (T * fD * sumr).max() / (32.0 * mkmax) if mkmax > 0 else 0.0))
            self.ipw = np.float32(1.0) / self.pw
        else:
            pw = float(self.pw)
            sig_m = pw / (T * fD * sumr).max()
            Pbar = 2.0 * self.rowp.astype(float) - self.rowprev.astype(float)
            du, sig_u = aE * Pbar, pw / (T * aE * sumr)
            capu, u0 = (self.E - self.c) / self.E, self.c / self.E
            slopes = np.diff(self.logv) / np.diff(self.bases) / (J * self.Tfull)
            v = du - self.mj / sig_u
            ps = capu.copy()
            done = np.zeros(J, bool)
            for b in range(len(slopes)):
                hi = np.minimum(np.maximum(self.bases[b + 1] - u0, 0.0), capu)
                lo = np.minimum(np.maximum(self.bases[b] - u0, 0.0), capu)
                cand = v + slopes[b] / sig_u
                sel = (~done) & (cand <= hi)
                ps[sel] = np.maximum(cand[sel], lo[sel])
                done |= sel
            self.mj = self.mj - sig_u * (du - ps)
            z = np.maximum(0.0, self.om + sig_m * (self.rem / DT - fD * Pbar))
            if kk <= 0:
                z[:] = 0.0
            elif z.sum() > kk:
                th = -1.0
                for _ in range(64):
                    a = z > th
                    nt = (z[a].sum() - kk) / max(a.sum(), 1)
                    if not nt > th:
                        break
                    th = nt
                z = np.maximum(z - th, 0.0)
            self.om = z
            lbar = (2.0 * self.colload.astype(float) - self.colprev.astype(float)) * self.icap
            self.pi = np.maximum(0.0, self.pi + pw / (self.g.astype(float).sum() * self.icap) * (lbar - 1.0))
        self.rowprev = self.rowp.copy()
        self.colprev = self.colload.copy()
        self.theta = (self.mj * aE + self.om * fD).astype(np.float32)
        self.price = (self.pi * self.icap).astype(np.float32)
        return obj

    def run(self, X, iters):
        X = self.scale(X.astype(np.float32), np.ones_like(self.icap))
        self.dual(0)
        obj = None
        for _ in range(iters):
            X = self.step(X)
            obj = self.dual(1)
        return X

    def finish(self, X):
        load = self.colprev
        obj = self.score(load)
        cs = np.where(load * self.icap > 1.0, 1.0 / np.maximum(load.astype(float) * self.icap, 1e-30), 1.0).astype(np.float32)
        X = self.scale(X, cs)
        objf = self.score(self.colload)
        return X, obj, objf


class RefMarket:
    def __init__(self, g, E, c, dbar, rem, rate, cap, k, D, bases, logv, T):
        self.args = (g, E, c, dbar, rem, rate)
        self.cap = np.asarray(cap, float)
        if self.cap.ndim == 1:
            self.cap = np.repeat(self.cap[:, None], T, axis=1)
        self.k, self.D, self.bases, self.logv, self.T = k, D, bases, logv, T

    def run(self, X0, iters, coarse_iters=0, primal_weight=60.0):
        """X0 [J][W][T] starting point (zeros when the kernel is called without warm start).
        Returns (X, (objective, makespan, violation) of the repaired X)."""
        J, W, T = X0.shape
        pw = np.float32(primal_weight / (J * T))
        fine = Level(*self.args, self.cap, self.k, self.D, self.bases, self.logv, T, 1.0, pw)
        X = X0.astype(np.float32)
        if coarse_iters > 0 and T >= 16:
            grp = T // 4
            capc = self.cap.reshape(W, 4, grp).sum(axis=2)
            icapc = (grp / capc)
            lc = Level(*self.args, 1.0 / icapc.astype(np.float32).astype(float), self.k, self.D, self.bases, self.logv,
                       T, float(grp), pw)
            lc.icap = icapc.astype(np.float32)
            Xc = (X.reshape(J, W, 4, grp).sum(axis=3) / np.float32(grp)).astype(np.float32)
            Xc = lc.run(Xc, coarse_iters)
            X = np.repeat(Xc, grp, axis=2)
            fine.mj, fine.om = lc.mj, lc.om
            fine.pi = np.repeat(lc.pi, grp, axis=1) / grp
        X = fine.run(X, iters)
        X, obj, objf = fine.finish(X)
        return X, objf
