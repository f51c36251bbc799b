"""Plain numpy fp32 reference of the dense market iteration (market.cu) — the same operation sequence,
used as the numerics oracle of that floating-point kernel."""
import numpy as np


def plog_slope(u, bases, logv):
    bases = np.asarray(bases); logv = np.asarray(logv)
    slope = np.diff(logv) / np.diff(bases)
    b = np.clip(np.searchsorted(bases, u, side="right") - 1, 0, len(bases) - 2)
    return slope[b]


def plog(u, bases, logv):
    return np.interp(u, bases, logv)


class RefMarket:
    def __init__(self, g, E, c, dbar, rem, rate, Gw, k, D, bases, logv, T):
        self.g = np.asarray(g, np.float32); self.E = np.asarray(E, float); self.c = np.asarray(c, float)
        self.dbar = np.asarray(dbar, float); self.rem = np.asarray(rem, float)
        self.rate = np.asarray(rate, np.float32); self.Gw = np.asarray(Gw, np.float32)
        self.k, self.D, self.bases, self.logv, self.T = k, D, np.asarray(bases, float), np.asarray(logv, float), T
        J, W = self.rate.shape
        self.theta = np.zeros(J, np.float32)
        self.colscale = np.ones((W, T), np.float32)
        self.price = np.zeros((W, T), np.float32)
        self.rowp = np.zeros(J, np.float32)
        self.colload = np.zeros((W, T), np.float32)

    def dense(self, X, eta):
        eta = np.float32(eta)
        gain = self.theta[:, None, None] * self.rate[:, :, None]
        cost = self.price[None] * self.g[:, None, None]
        Y = X * self.colscale[None] + eta * ((gain - cost) / (gain + cost + np.float32(1e-30)))
        Y = np.clip(Y, 0.0, 1.0).astype(np.float32)
        tot = Y.sum(axis=1, keepdims=True)
        Y = np.where(tot > 1.0, Y / np.maximum(tot, 1e-30), Y).astype(np.float32)
        self.rowp = (self.rate[:, :, None] * Y).sum(axis=(1, 2)).astype(np.float32)
        self.colload = (self.g[:, None, None] * Y).sum(axis=0).astype(np.float32)
        return Y

    def dual(self, sigma, theta_scale=1.0, init_price=False):
        J = len(self.E)
        P = np.minimum(self.rowp.astype(float), self.E - self.c)
        u = (self.c + P) / self.E
        welfare = plog(u, self.bases, self.logv).sum() / (J * self.T)
        remj = np.maximum(0.0, self.rem - self.dbar * P)
        mx = remj.max()
        band = 1e-3 * self.D
        unfinished = self.rowp.astype(float) < self.E - self.c
        crit = (mx > 0) & (remj >= mx - band) & (P < self.E - self.c)
        cnt = crit.sum()
        th = np.where(unfinished, plog_slope(u, self.bases, self.logv) / self.E / (J * self.T), 0.0)
        if cnt > 0:
            th = th + np.where(unfinished & (mx > 0) & (remj >= mx - band), self.k * self.dbar / cnt, 0.0)
        self.theta = (th * theta_scale).astype(np.float32)
        cap = self.Gw[:, None]
        viol = float((self.colload / cap - 1.0).max())
        self.colscale = np.where((self.colload > cap) & (self.colload > 0), cap / np.maximum(self.colload, 1e-30),
                                 1.0).astype(np.float32)
        if init_price:
            dens = (self.theta.astype(float)[:, None] * self.rate.astype(float)).sum(axis=0) / self.g.astype(float).sum()
            self.price = np.broadcast_to(dens[:, None], self.price.shape).astype(np.float32)
        self.price = np.maximum(np.float32(1e-30), self.price * np.exp(np.float32(sigma) * (self.colload / cap - 1.0))
                                ).astype(np.float32)
        return welfare - self.k * mx, mx, viol

    def run(self, X, iters, eta, sigma, theta_scale=1.0, eta_decay=0.0):
        X = self.dense(X.astype(np.float32), 0.0)
        for it in range(iters):
            self.dual(sigma, theta_scale, init_price=(it == 0))
            X = self.dense(X, eta / (1.0 + it / eta_decay) if eta_decay > 0 else eta)
        self.dual(sigma, theta_scale)
        X = self.dense(X, 0.0)
        obj = self.dual(sigma, theta_scale)
        return X, obj
