"""CPU sanity of tests/ref_placement.py (the numpy restatement of place.cu): capacity per round, counts honoured or a
shortfall reported, back-fill disjoint from the schedule and work-conserving."""
import numpy as np

from tests.ref_placement import place


def test_invariants_random():
    rng = np.random.default_rng(0)
    for t in range(40):
        J = int(rng.integers(3, 80))
        G = int(rng.choice([8, 12, 16, 32, 64]))
        T = int(rng.integers(3, 25))
        g = np.minimum(rng.choice([1, 2, 4, 8], J, p=[.6, .3, .09, .01]), G)
        n = np.zeros(J, dtype=int)
        cap = G * T
        for j in rng.permutation(J):
            m = min(int(rng.integers(0, T + 1)), cap // g[j])
            n[j] = m
            cap -= m * g[j]
        w = rng.uniform(1, 100, J)
        key = rng.uniform(0, 1e5, J)
        for fb in (False, True):
            r = place(n, g, G, T, key, fallback=fb, w=w)
            x, bf = r["x"], r["backfill"]
            assert ((x * g[:, None]).sum(0) <= G).all()
            assert (((x | bf) * g[:, None]).sum(0) <= G).all() and not (x & bf).any()
            assert r["shortfall"] == int((n - x.sum(1)).sum()) and r["shortfall"] >= 0
            # work-conserving: no job left out of a round could still fit into what stays idle there
            load = ((x | bf) * g[:, None]).sum(0)
            for tt in range(T):
                out = ~(x[:, tt] | bf[:, tt])
                assert not np.any(g[out] <= G - load[tt])


def test_unit_widths_always_pack_and_sweep_orders_by_priority():
    rng = np.random.default_rng(1)
    J, G, T = 30, 8, 10
    g = np.ones(J, dtype=int)
    n = rng.integers(0, 4, J)
    n[:G] = np.minimum(n[:G] + 2, T)
    while (n * g).sum() > G * T:
        n[n.argmax()] -= 1
    w = rng.uniform(1, 1e6, J)
    r = place(n, g, G, T, np.zeros(J), fallback=True, w=w)
    assert r["shortfall"] == 0 and r["swept_rounds"] == T      # McNaughton: unit widths never fragment
    # the job with the highest priority density sits in the earliest rounds
    dens = np.where(n > 0, w / np.maximum(n, 1), -1)
    top = int(dens.argmax())
    assert r["x"][top, :n[top]].all()
