"""shockwave_b200.sweep (command-line what-if sweeps) on CPU stand-ins: the reference's own parsing feeds the ensembles,
one result pickle per point; the points that coincide with a run of the unmodified reference loop reproduce its pickle
fields (makespan, completion times, finish-time fairness, utilisation, lease extensions)."""
import os
import pickle
import tempfile

import numpy as np
import pytest

from oracle import gavel_backend as gb
from oracle import ref_harness as rh
from oracle.gavel_round_backend import OracleBackend
from tests import sim_fixtures as sf_

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="staged reference not present")


def test_sweep_cli_policies_and_shockwave(monkeypatch):
    if sf_.host_sim_lib() is None:
        pytest.skip("g++ not available")
    from shockwave_b200 import simulate as sim
    from shockwave_b200 import sweep
    from tests.golden import make_sim_pins as pins
    scratch = tempfile.mkdtemp(prefix="swcli_")
    dst = pins.stage_static_trace(scratch, keep=40, static=False)
    trace = os.path.join(dst, pins.REL)
    thr = os.path.join(dst, "tacc_throughputs.json")
    out = os.path.join(scratch, "results")
    monkeypatch.setattr(sim, "DeviceSim", sf_.HostDeviceSim)
    orig = sim.PolicyEnsemble.__init__
    monkeypatch.setattr(sim.PolicyEnsemble, "__init__",
                        lambda self, *a, **k: orig(self, *a, **{**k, "round_backend": OracleBackend()}))
    names = ["max_min_fairness", "finish_time_fairness"]
    with gb.cpu_backend() as P:
        refs = {n: rh.simulate(n, policy_obj=P.get_policy(n, solver="ECOS", seed=0), trace=pins.REL, scratch=scratch,
                               cluster="8:0:0") for n in names}
        cwd = os.getcwd()
        paths = sweep.main(["--reference-dir", dst, "-t", trace, "--throughputs_file", thr, "-c", "8:0:0",
                            "-p", ",".join(names), "--output_dir", out])
        assert os.getcwd() == cwd
    for n, p in zip(names, paths):
        d = pickle.load(open(p, "rb"))
        r = refs[n]
        assert d["policy"] == n and d["makespan"] == r["makespan"]
        assert d["jct_list"] == list(r["jct_list"])
        assert d["finish_time_fairness_list"] == list(r["finish_time_fairness_list"])
        assert d["cluster_util"] == float(r["cluster_util"])
        assert [dict(x) for x in d["per_round_schedule"]] == [{int(k): tuple(v) for k, v in rnd.items()}
                                                              for rnd in r["per_round_schedule"]]
    # ---- shockwave grid with the rule scheduler standing in for the device solves
    Rule = sf_.make_rule_scheduler_cls()
    monkeypatch.setattr(sim, "ShockwaveScheduler", Rule)
    ref = rh.simulate("shockwave", shockwave_scheduler_cls=Rule, trace=pins.REL, scratch=scratch, cluster="8:0:0")
    cfg = os.path.join(dst, "configurations", "tacc_32gpus.json")
    paths = sweep.main(["--reference-dir", dst, "-t", trace, "--throughputs_file", thr, "-c", "8:0:0", "-p", "shockwave",
                        "--config", cfg, "--set", "future_rounds=20,4", "--set", "k=0.001,0.1", "--output_dir", out])
    assert len(paths) == 4
    ds = [pickle.load(open(p, "rb")) for p in paths]
    assert [d["hyperparameters"] for d in ds] == [{"future_rounds": 20, "k": 0.001}, {"future_rounds": 20, "k": 0.1},
                                                  {"future_rounds": 4, "k": 0.001}, {"future_rounds": 4, "k": 0.1}]
    assert ds[0]["makespan"] == ref["makespan"] and ds[0]["jct_list"] == list(ref["jct_list"])
    assert ds[0]["finish_time_fairness_list"] == list(ref["finish_time_fairness_list"])
    assert ds[2]["per_round_schedule"] != ds[0]["per_round_schedule"]


def test_sweep_cli_gavel_policies_on_a_mixed_cluster(monkeypatch):
    """--cluster_spec 4:3:2 (v100:p100:k80), static trace: the pickles of the heterogeneity-aware policies carry the
    reference loop's makespan, completion times, utilisation and `{job: worker ids}` schedule."""
    if sf_.host_sim_lib() is None:
        pytest.skip("g++ not available")
    from shockwave_b200 import simulate as sim
    from shockwave_b200 import sweep
    from tests.golden import make_sim_pins as pins
    scratch = tempfile.mkdtemp(prefix="swclih_")
    dst = pins.stage_static_trace(scratch, keep=40, static=True)
    trace = os.path.join(dst, pins.REL)
    thr = os.path.join(dst, "tacc_throughputs.json")
    out = os.path.join(scratch, "results")
    monkeypatch.setattr(sim, "DeviceSim", sf_.HostDeviceSim)
    orig = sim.PolicyEnsemble.__init__
    monkeypatch.setattr(sim.PolicyEnsemble, "__init__",
                        lambda self, *a, **k: orig(self, *a, **{**k, "round_backend": OracleBackend()}))
    names = ["max_min_fairness_perf", "max_min_fairness"]
    with gb.cpu_backend() as P:
        refs = {n: rh.simulate(n, policy_obj=P.get_policy(n, solver="ECOS", seed=0), trace=pins.REL, scratch=scratch,
                               cluster="4:3:2") for n in names}
        paths = sweep.main(["--reference-dir", dst, "-t", trace, "--throughputs_file", thr, "-c", "4:3:2",
                            "-p", ",".join(names), "--output_dir", out])
        with pytest.raises(SystemExit):
            sweep.main(["--reference-dir", dst, "-t", trace, "--throughputs_file", thr, "-c", "4:3:2", "-p", "shockwave",
                        "--config", os.path.join(dst, "configurations", "tacc_32gpus.json"), "--output_dir", out])
    for n, p in zip(names, paths):
        d = pickle.load(open(p, "rb"))
        r = refs[n]
        assert d["policy"] == n and d["makespan"] == r["makespan"]
        assert d["jct_list"] == list(r["jct_list"])
        assert d["cluster_util"] == float(r["cluster_util"])
        assert d["finish_time_fairness_list"] == list(r["finish_time_fairness_list"])
        assert [dict(x) for x in d["per_round_schedule"]] == [{int(k): tuple(v) for k, v in rnd.items()}
                                                              for rnd in r["per_round_schedule"]]
