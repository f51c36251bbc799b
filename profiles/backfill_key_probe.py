"""How often do the reference's PER-ROUND back-fill sort keys differ from the first key of the same re-solve?
(construct_schedules calls dirichlet_posterior_remaining_runtime() again in every round with idle GPUs and every call
recalibrates, shockwave.py:261-267; place.cu sorts every round on the first key and replays only the STATE change of the
later calls.)  Runs the UNMODIFIED reference simulator on the canonical trace with the HiGHS oracle in the loop (CPU,
~1-2 min, needs /root/reference or the staged copy) and writes profiles/backfill_key_probe_r02.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as rh  # noqa: E402

out = {}
for tag, kw in (("oracle_x", {}), ("product_placement_rule", {"placement": "place"})):
    if kw.get("placement") == "place":
        from tests.ref_placement import place
        kw = {"placement": place}
    st = {}
    res = rh.simulate("shockwave", shockwave_scheduler_cls=rh.make_oracle_scheduler_cls(key_stats=st, **kw))
    st.pop("_first", None)
    st.update(makespan=res["makespan"], rounds=len(res["per_round_schedule"]))
    out[tag] = st
    print(tag, st, flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "backfill_key_probe_r02.json"), "w"), indent=1)
