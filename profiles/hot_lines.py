"""Aggregates an ncu source page (`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`) into the hottest
CUDA source lines: warp-stall samples (≈ time) and executed warp instructions per line.  Usage:
python profiles/hot_lines.py gpurun_out/place_single_r02.ncu-rep [top]"""
import csv
import io
import subprocess
import sys


def main(rep, top=40):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    h = rows[hdr]
    si, ii = h.index("# Samples"), h.index("Instructions Executed")
    cur, agg, src, fname = None, {}, {}, ""
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) <= max(si, ii) or r[0] == "Line No":
            continue
        if r[0]:
            cur = (fname, int(r[0]))
            src[cur] = r[1]
            continue
        if cur is None:
            continue
        a = agg.setdefault(cur, [0, 0])
        try:
            a[0] += int(r[si]); a[1] += int(r[ii])
        except ValueError:
            pass
    ts = sum(a[0] for a in agg.values()) or 1
    ti = sum(a[1] for a in agg.values()) or 1
    print(f"{rep}: {ts} samples, {ti} warp instructions")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{a[0] / ts:6.3f} smp {a[1] / ti:6.3f} ins  {k[0]}:{k[1]:<4d} {src[k].strip()[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
