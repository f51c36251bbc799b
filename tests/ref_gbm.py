"""Plain numpy restatement of gbm.cu for ONE sample path (test infrastructure): the same stream keying
(splitmix64 of seed ^ job ^ global PAIR id -> xorshift128+; global paths 2q and 2q+1 are an antithetic pair: the same
normals, W and -W), the same Box-Muller pairing (one 64-bit draw ->
two 24-bit uniforms -> two normals used for consecutive epochs), the same running sum along the path.  The kernel uses
the fast fp32 intrinsics (__logf, __sincosf, __expf), so values agree to ~1e-5 relative, not bit for bit."""
import numpy as np

M64 = (1 << 64) - 1


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return x, (z ^ (z >> 31)) & M64


def path_value(seed, j, gp, R0, H, mu, sigma):
    """R^(p) of job j on global path gp (gbm.cu:44-71)."""
    sign = np.float32(-1.0 if (gp & 1) else 1.0)
    sm = (seed ^ ((0xD1B54A32D192ED03 * (j + 1)) & M64) ^ (((gp >> 1) * 0x9E3779B97F4A7C15) & M64)) & M64
    sm, s0 = _splitmix64(sm)
    sm, s1 = _splitmix64(sm)
    s1 |= 1
    f = np.float32
    mu32, sg = f(mu), f(sigma)
    drift = f(mu32 - f(0.5) * sg * sg)
    W, acc = f(0.0), f(0.0)
    h = 1
    while h <= H:
        x, y = s0, s1                               # xorshift128+
        s0 = y
        x ^= (x << 23) & M64
        s1 = (x ^ y ^ (x >> 17) ^ (y >> 26)) & M64
        r = (s1 + y) & M64
        u1 = f((f(r >> 40) + f(1.0)) * f(1.0 / 16777217.0))
        u2 = f(f((r >> 8) & 0xFFFFFF) * f(1.0 / 16777216.0))
        rad = f(np.sqrt(f(-2.0) * f(np.log(u1))))
        ang = f(6.283185307179586) * u2
        W = f(W + rad * f(np.cos(ang)))
        acc = f(acc + f(np.exp(f(f(drift * f(h)) + sign * f(sg * W)))))
        if h + 1 <= H:
            W = f(W + rad * f(np.sin(ang)))
            acc = f(acc + f(np.exp(f(f(drift * f(h + 1)) + sign * f(sg * W)))))
        h += 2
    return float(R0) * float(f(acc * f(1.0 / H))) if H > 0 else float(R0)
