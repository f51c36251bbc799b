"""What-if sweeps of a trace from the command line: the counterpart of the reference's
`scripts/drivers/simulate_scheduler_with_trace.py` (one run, one policy, one hyper-parameter point, minutes of Python per
run) for MANY points at once on the device round loop (simulate.py):

    python -m shockwave_b200.sweep --reference-dir /path/to/shockwave/scheduler \\
        --trace_file traces/reproduce/120_..._dynamic.trace --throughputs_file tacc_throughputs.json \\
        --cluster_spec 32:0:0 --policy shockwave --config configurations/tacc_32gpus.json \\
        --set k=1e-3,1e-2,1e-1 --set lambda=12,24 --output_dir results/

    python -m shockwave_b200.sweep ... --policy max_min_fairness,finish_time_fairness,min_total_duration

Trace parsing, profile generation and the throughput file stay the reference's (`utils.generate_pickle_file`,
`utils.read_all_throughputs_json_v2`, `utils.get_gns_bs_pattern`: imported from --reference-dir, exactly the calls its own
driver makes, simulate_scheduler_with_trace.py:28-43); everything after that runs here.  One result pickle per point with
the keys of the reference's pickle (see `ShockwaveEnsemble.result_dicts` / `PolicyEnsemble.result_dicts`), readable by
`aggregate_result.py`.  Single jobs.  Shockwave runs on one worker type (the first count of --cluster_spec, as in the
reference: "we assume homogeneous hardware"); the Gavel policies also run MIXED clusters (`--cluster_spec 8:4:4` =
v100:p100:k80) of static traces — per-type throughputs from the same throughput file, worker ids handed out like the
reference does (PolicyEnsemble(worker_types=...), swb_sim_set_worker_types)."""
import argparse
import importlib
import itertools
import json
import os
import pickle
import sys

import numpy as np

from . import simulate as _sim


def load_reference_inputs(reference_dir, trace_file, throughputs_file, worker_type="v100"):
    """The reference's own parsing, as its driver calls it.  Returns (trace arrays, profiles, models, modes,
    throughput_of, gns_pattern, isolated durations)."""
    for p in (reference_dir, os.path.join(reference_dir, "policies")):
        if p not in sys.path:
            sys.path.append(p)
    utils = importlib.import_module("utils")
    jobs, arrival_times = utils.generate_pickle_file(trace_file, throughputs_file)
    with open(os.path.splitext(trace_file)[0] + ".pickle", "rb") as f:
        profiles = pickle.load(f)
    table = utils.read_all_throughputs_json_v2(throughputs_file)[worker_type]

    def throughput_of(model, bs, sf):
        row = table.get((f"{model} (batch size {bs})", sf))
        return None if row is None else float(row["null"])
    iso = [sum(p["duration_every_epoch"]) for p in profiles[:len(jobs)]]
    trace = dict(arrival=np.asarray(arrival_times, np.float64),
                 total_steps=np.asarray([j.total_steps for j in jobs], np.int64),
                 scale_factor=np.asarray([j.scale_factor for j in jobs], np.int32),
                 batch_size=np.asarray([j.batch_size for j in jobs], np.int32),
                 duration=np.asarray(iso, np.float64),                      # simulate_scheduler_with_trace.py:36-41
                 dataset_len=np.asarray([utils.dataset_len[utils.model_dataset_mapping[j.model]] for j in jobs], np.int64),
                 throughput=np.asarray([throughput_of(j.model, j.batch_size, j.scale_factor) for j in jobs], np.float64),
                 priority_weight=np.asarray([j.priority_weight for j in jobs], np.float64))
    return trace, profiles, [j.model for j in jobs], [j.mode for j in jobs], throughput_of, utils.get_gns_bs_pattern, iso


def per_type_throughputs(reference_dir, throughputs_file, models, trace, names):
    """[J][W] throughputs of the trace's jobs on the worker types `names` (`Scheduler._set_initial_throughput`,
    scheduler.py:579-590: the "null" co-location entry of the job type at its scale factor)."""
    utils = importlib.import_module("utils")
    tables = utils.read_all_throughputs_json_v2(throughputs_file)
    out = np.zeros((len(models), len(names)))
    for j, (m, bs, sf) in enumerate(zip(models, trace["batch_size"], trace["scale_factor"])):
        for i, w in enumerate(names):
            out[j, i] = float(tables[w][(f"{m} (batch size {int(bs)})", int(sf))]["null"])
    return out


def grid(sets):
    """--set k=1e-3,1e-2 --set lambda=12,24 -> [{k: 1e-3, lambda: 12}, ...] (cartesian product, first key slowest)."""
    keys, vals = [], []
    for item in sets or []:
        k, v = item.split("=", 1)
        keys.append(k)
        vals.append([json.loads(x) for x in v.split(",")])
    return [dict(zip(keys, combo)) for combo in itertools.product(*vals)] or [{}]


def main(argv=None):
    ap = argparse.ArgumentParser(description="what-if sweeps of a trace on the device round loop")
    ap.add_argument("--reference-dir", required=True, help="the reference's scheduler/ directory (utils.py, job.py ...)")
    ap.add_argument("-t", "--trace_file", required=True)
    ap.add_argument("--throughputs_file", required=True)
    ap.add_argument("-c", "--cluster_spec", default="32:0:0", help="v100:p100:k80 like the reference (shockwave: v100 only)")
    ap.add_argument("-p", "--policy", default="shockwave", help="shockwave, or a comma-separated list of Gavel policies")
    ap.add_argument("--config", help="shockwave_config json (reference: configurations/*.json)")
    ap.add_argument("--set", action="append", help="shockwave hyper-parameter values, e.g. k=1e-3,1e-2 (repeatable)")
    ap.add_argument("--time_per_iteration", type=int, default=120)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--output_dir", required=True)
    args = ap.parse_args(argv)
    counts = [int(v) for v in args.cluster_spec.split(":")]
    mixed = any(counts[1:])
    if mixed and args.policy == "shockwave":
        raise SystemExit("shockwave assumes homogeneous hardware (as the reference does): use a v100-only --cluster_spec")
    trace, profiles, models, modes, thr_of, gns, iso = load_reference_inputs(args.reference_dir, args.trace_file,
                                                                             args.throughputs_file)
    dyn = None
    if any(m != "static" for m in modes):
        if mixed:
            raise SystemExit("mixed clusters run static traces only (the reference keeps a rescaled job's progress on "
                             "v100 alone, scheduler.py:4896-4925)")
        dyn = _sim.build_dynamic_tables(models, modes, trace, thr_of, gns)
    worker_types = None
    if mixed:
        names = sorted(n for n, c in zip(("v100", "p100", "k80"), counts) if c > 0)
        worker_types = dict(names=names, ngpus=[counts[("v100", "p100", "k80").index(n)] for n in names],
                            throughput=per_type_throughputs(args.reference_dir, args.throughputs_file, models, trace, names))
    os.makedirs(args.output_dir, exist_ok=True)
    stem = os.path.splitext(os.path.basename(args.trace_file))[0]
    paths = []
    if args.policy == "shockwave":
        if not args.config:
            raise SystemExit("--policy shockwave needs --config")
        cfg = json.load(open(args.config))
        points = grid(args.set)
        ens = _sim.ShockwaveEnsemble(trace, profiles, cfg, points, ngpus=counts[0], time_per_iteration=args.time_per_iteration,
                                     device=args.device, dynamic=dyn)
        ens.run()
        for d, pt in zip(ens.result_dicts(trace_file=args.trace_file), points):
            d["hyperparameters"] = pt
            tag = "_".join(f"{k}={v}" for k, v in pt.items()) or "default"
            paths.append(os.path.join(args.output_dir, f"shockwave_{tag}_{stem}_simulation.pickle"))
            pickle.dump(d, open(paths[-1], "wb"))
    else:
        from . import policies as P
        names = args.policy.split(",")
        pols = [P.get_policy(n, solver="ECOS", seed=args.seed) for n in names]
        ens = _sim.PolicyEnsemble(trace, pols, counts[0], time_per_iteration=args.time_per_iteration, device=args.device,
                                  dynamic=dyn, priority_weights=trace["priority_weight"], worker_types=worker_types,
                                  seed=args.seed)
        ens.run()
        for d, n in zip(ens.result_dicts(iso, trace_file=args.trace_file), names):
            d["policy"] = n
            paths.append(os.path.join(args.output_dir, f"{n}_{stem}_simulation.pickle"))
            pickle.dump(d, open(paths[-1], "wb"))
    for p in paths:
        print(p)
    return paths


if __name__ == "__main__":
    main()
