"""Instances for the packing-policy tests: a duck-typed stand-in for the reference's JobIdPair (job_id_pair.py:4-107:
ordering singles-before-pairs, hashing, singletons(), as_tuple(), overlaps_with()) and a seeded generator of
throughput dicts with co-location slowdowns."""
import numpy as np

WT = ["k80", "p100", "v100"]


class JobId:
    def __init__(self, a, b=None):
        if b is not None:
            a, b = min(a, b), max(a, b)
        self._a, self._b = a, b

    def __getitem__(self, k):
        return (self._a, self._b)[k]

    def __lt__(self, o):
        if o._b is not None:
            if self._b is None:
                return True
            if self._a == o._a:
                return self._b < o._b
        elif self._b is not None:
            return False
        return self._a < o._a

    def __eq__(self, o):
        if type(o) is int:
            return self._a == o
        return self._a == o._a and self._b == o._b

    def __hash__(self):
        return self._a if self._b is None else self._a + self._b * self._b

    def __repr__(self):
        return f"{self._a}" if self._b is None else f"({self._a}, {self._b})"

    def is_pair(self):
        return self._b is not None

    def as_tuple(self):
        return (self._a, self._b)

    def singletons(self):
        return (self,) if self._b is None else (JobId(self._a), JobId(self._b))

    def overlaps_with(self, o):
        return self._a in (o._a, o._b)


def instance(ns, spec, seed, pair_fraction=1.0, make_id=JobId):
    """ns single jobs + a random subset of their pairs.  Returns (throughputs, scale_factors, priority_weights,
    times_since_start, num_steps_remaining, cluster_spec, singles)."""
    rng = np.random.default_rng(seed)
    singles = [make_id(i, None) for i in range(ns)]
    speed = {"k80": 1.0, "p100": 2.2, "v100": 3.5}
    base = rng.uniform(0.5, 20.0, ns)
    sf = {s: int(rng.choice([1, 2, 4], p=[0.7, 0.2, 0.1])) for s in singles}
    thr = {}
    for i, s in enumerate(singles):
        thr[s] = {w: float(base[i] * speed[w] * rng.uniform(0.7, 1.0)) for w in WT}
    for i in range(ns):
        for j in range(i + 1, ns):
            if rng.uniform() > pair_fraction:
                continue
            p = make_id(i, j)
            slow = rng.uniform(0.35, 0.75, (len(WT), 2))       # each member keeps 35-75 % of its solo throughput
            thr[p] = {w: [float(thr[singles[i]][w] * slow[k, 0]), float(thr[singles[j]][w] * slow[k, 1])]
                      for k, w in enumerate(WT)}
    prio = {s: float(rng.choice([1.0, 2.0, 5.0])) for s in singles}
    t0 = {s: float(rng.uniform(0.0, 5000.0)) for s in singles}
    steps = {s: float(rng.uniform(2e3, 4e5)) for s in singles}
    return thr, sf, prio, t0, steps, dict(spec), singles


def job_type_instance(n, ntypes, spec, seed):
    """n jobs of `ntypes` job types (name, scale_factor); throughputs[type][worker_type][other type or None] as the
    reference's job-type formulation takes them (max_min_fairness.py:122-180)."""
    rng = np.random.default_rng(seed)
    speed = {"k80": 1.0, "p100": 2.2, "v100": 3.5}
    keys = [(f"model{t}", int(rng.choice([1, 1, 2]))) for t in range(ntypes)]
    base = rng.uniform(0.5, 20.0, ntypes)
    thr = {}
    slow = rng.uniform(0.35, 0.95, (ntypes, ntypes))
    for a, ka in enumerate(keys):
        thr[ka] = {}
        for w in spec:
            row = {None: float(base[a] * speed.get(w, 1.5))}
            for b, kb in enumerate(keys):
                row[kb] = float(base[a] * speed.get(w, 1.5) * slow[a, b])
            thr[ka][w] = row
    job_ids = [JobId(i, None) for i in range(n)]
    tof = list(rng.integers(0, ntypes, n))
    for t in range(min(ntypes, n)):            # every type present once at least, one type possibly a singleton
        tof[t] = t
    j2k = {j: keys[t] for j, t in zip(job_ids, tof)}
    sf = {j: j2k[j][1] for j in job_ids}
    prio = {j: float(rng.choice([1.0, 2.0])) for j in job_ids}
    return thr, j2k, sf, prio
