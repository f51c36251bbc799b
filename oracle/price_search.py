"""ctypes loader of oracle/libprice_search.so (plain-C restatement of solve.cu's ALGORITHM on host cores; see the header
of oracle/price_search.c).  Test / baseline infrastructure: only tests/ and bench.py's CPU legs may import it."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libprice_search.so")


def build():
    src = os.path.join(HERE, "price_search.c")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", src, "-o", SO, "-lm"], check=True)
    return SO


def price_search(k, round_ptr, g, E, c, dbar, rem, ftobj, G, T, D, lam, rhomax, bases, logv, nthreads=0):
    """S scenarios ([S, J] arrays, k [S], round_ptr [S]) -> dict(n [S, J], objective [S], status [S], evals [S])."""
    lib = C.CDLL(build())
    i32 = lambda v: np.ascontiguousarray(v, dtype=np.int32)
    f64 = lambda v: np.ascontiguousarray(v, dtype=np.float64)
    g, E, c, dbar, rem, ftobj = i32(g), i32(E), i32(c), f64(dbar), f64(rem), f64(ftobj)
    if g.ndim == 1:
        g, E, c, dbar, rem, ftobj = (a[None] for a in (g, E, c, dbar, rem, ftobj))
    S, J = g.shape
    k = f64(np.broadcast_to(k, (S,))); rp = i32(np.broadcast_to(round_ptr, (S,)))
    bases, logv = f64(bases), f64(logv)
    n = np.zeros((S, J), dtype=np.int32); obj = np.zeros(S); st = np.zeros(S, dtype=np.int32); ev = np.zeros(S, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.sw_price_search.restype = C.c_int
    lib.sw_price_search(C.c_int(S), C.c_int(J), C.c_int(T), C.c_int(G), C.c_int(len(bases)), C.c_double(D), p(k),
                        C.c_double(lam), C.c_double(rhomax), p(rp), p(bases), p(logv), p(g), p(E), p(c), p(dbar),
                        p(rem), p(ftobj), p(n), p(obj), p(st), p(ev), C.c_int(nthreads or (os.cpu_count() or 1)))
    return dict(n=n, objective=obj, status=st, evals=ev)
