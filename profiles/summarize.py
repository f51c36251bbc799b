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


def main(tag):
    out = []
    for rep in sorted(glob.glob(f"gpurun_out/*_{tag}.ncu-rep")):
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else rep
        out.append(f"== {os.path.basename(rep)} :: {name}")
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                out.append(f"   {h} = {v} {u}")
    open(f"profiles/ncu_summary_{tag}.txt", "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
