"""Turns the .ncu-rep captures under gpurun_out/ into the text summaries committed under profiles/."""
import csv
import glob
import io
import os
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "launch__shared_mem_per_block_dynamic"]


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
        "msecond": 1e3, "nsecond": 1e-3, "second": 1e6, "s": 1e6}


def _num(v, u):
    try:
        return float(v.replace(",", "")) * UNIT.get(u, 1.0)
    except ValueError:
        return None


def main(tag):
    out = []
    kernels = {}
    for rep in sorted(glob.glob(f"gpurun_out/*_{tag}.ncu-rep")):
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else rep
        out.append(f"== {os.path.basename(rep)} :: {name}")
        row = {}
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                out.append(f"   {h} = {v} {u}")
                row[h] = _num(v, u)
        rd, wr = row.get("dram__bytes_read.sum"), row.get("dram__bytes_write.sum")
        kernels[os.path.basename(rep).replace(f"_{tag}.ncu-rep", "") + " :: " + name] = {"capture": os.path.basename(rep), "duration_us": row.get("gpu__time_duration.sum"),
                         "dram_bytes": (rd + wr) if rd is not None and wr is not None else None,
                         "dram_bytes_read": rd, "dram_bytes_write": wr,
                         "issue_active_pct": row.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                         "sm_throughput_pct": row.get("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                         "warps_active_pct": row.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
                         "dram_throughput_pct": row.get("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
                         "grid": row.get("launch__grid_size"), "block": row.get("launch__block_size"),
                         "registers": row.get("launch__registers_per_thread")}
    open(f"profiles/ncu_summary_{tag}.txt", "w").write("\n".join(out) + "\n")
    import json
    json.dump(kernels, open(f"profiles/ncu_kernels_{tag}.json", "w"), indent=1)    # bench.py reads roofline.traffic here
    print("\n".join(out))




def shares(tag, csv_path=None):
    """Per (kernel, grid size) totals of the ncu launch list (gpu__time_duration.sum per launch) of the bench command.
    The headline step launches gbm_kernel (grid 4096), gbm_ensemble_kernel, solve_kernel<1,512,8> (grid 8: the 8-CTA
    cluster) and place_kernel (grid 1 or 8); the batched leg launches solve / place with grid 296; the other kernels
    belong to the side legs (dense-market roofline leg, Gavel-policy latency, plug-in calls)."""
    csv_path = csv_path or f"gpurun_out/launches_{tag}.csv"
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, mi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Grid Size")
    tot = {}
    for r in rows:
        if r is hdr or len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ki].split("(swb::")[0].split("(")[0] if "<" not in r[ki] else r[ki].split(">(")[0] + ">"
        grid = int(r[gi].strip("()").split(",")[0])
        ms = float(r[vi].replace(",", "")) / 1e6          # ns -> ms
        t = tot.setdefault((name, grid), [0, 0.0])
        t[0] += 1
        t[1] += ms
    total = sum(v[1] for v in tot.values())
    lines = ["ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (cold-cache, serialised: compare SHARES)", "",
             f"{'kernel':<64s} {'grid':>6s} {'n':>4s} {'mean_ms':>9s} {'total_ms':>9s} {'share':>6s}"]
    for (name, grid), (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{name:<64s} {grid:6d} {n:4d} {ms / n:9.3f} {ms:9.3f} {ms / total:6.3f}")
    # the headline step: gbm (grid 4096) + cluster solve (grid 8) + place launched right after it
    step = {k: v for k, v in tot.items() if ("gbm_kernel" in k[0] and k[1] >= 1024) or ("solve_kernel" in k[0] and k[1] == 8)
            or ("place_kernel" in k[0] and k[1] in (1, 8)) or "gbm_ensemble" in k[0]}
    ssum = sum(v[1] / v[0] for v in step.values())
    lines += ["", "-- kernels of the headline step (one re-solve): mean per launch, share of their sum --"]
    for (name, grid), (n, ms) in sorted(step.items(), key=lambda kv: -kv[1][1] / kv[1][0]):
        lines.append(f"{name:<64s} {grid:6d} {n:4d} {ms / n:9.3f} share={ms / n / ssum:6.3f}")
    open(f"profiles/launch_shares_{tag}.txt", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    main(tag)
    if os.path.exists(f"gpurun_out/launches_{tag}.csv"):
        shares(tag)
