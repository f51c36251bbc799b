"""Closed-loop pins of the Gavel-policy oracle at the reference's scale-out sizes: the UNMODIFIED reference simulator
driven by shockwave_b200/policies.py with its device calls routed to the HiGHS oracle (oracle/gavel_backend.py) on the
traces of reproduce/scale_{64,128,256}gpus.sh (220 / 460 / 900 jobs, wisr_throughputs.json).  The reference ships no golden pickles
for these runs; the pins are what the closed-loop GPU test compares the kernels-in-the-loop runs against.
Needs /root/reference; no GPU.  python tests/golden/make_scale_policy_pins.py  -> tests/golden/scale_policy_pins.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gavel_backend as gb                  # noqa: E402
from oracle import ref_harness as rh                    # noqa: E402
from tests.golden.make_policy_pins import summary       # noqa: E402
from tests.golden.make_scale_pins import TRACES         # noqa: E402

POLICIES = ["max_min_fairness", "finish_time_fairness", "min_total_duration"]

if __name__ == "__main__":
    path = os.path.join(ROOT, "tests/golden/scale_policy_pins.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for G in (64, 128, 256):
        for name in POLICIES:
            key = f"{name}@{G}"
            if key in out:
                continue
            t0 = time.time()
            with gb.cpu_backend() as P:
                pol = P.get_policy(name, solver="ECOS", seed=0)
                res = rh.simulate(name, policy_obj=pol, cluster=f"{G}:0:0", trace=TRACES[G], throughputs="wisr_throughputs.json")
            out[key] = dict(summary(res), trace=TRACES[G], cluster=f"{G}:0:0", throughputs="wisr_throughputs.json",
                            seconds=time.time() - t0)
            print(key, out[key], flush=True)
            json.dump(out, open(path, "w"), indent=1)
