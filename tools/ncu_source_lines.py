#!/usr/bin/env python
"""Aggregate the source page of an Nsight Compute report per CUDA source line:
  python tools/ncu_source_lines.py report.ncu-rep KERNEL_REGEX [launch_skip] [top]
prints, for the selected launch, the lines with the most stall samples and executed warp instructions."""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre,
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr = "", None
    agg = defaultdict(lambda: defaultdict(float))
    src = {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr) or not r[0].strip().isdigit():
            continue   # SASS rows carry no line number: the source-line rows already aggregate them
        d = dict(zip(hdr, r))   # duplicate 'Source' header: the last one (SASS) wins
        key = (cur_file, int(r[0]))
        src[key] = r[1]
        for k in ("# Samples", "Instructions Executed", "stall_barrier", "stall_long_sb", "stall_wait", "stall_no_inst", "stall_short_sb", "stall_lg", "stall_membar",
                  "stall_mio", "stall_math", "stall_branch_resolving", "stall_not_selected"):
            try:
                agg[key][k] += float(d.get(k, 0) or 0)
            except ValueError:
                pass
    tot_s = sum(v["# Samples"] for v in agg.values()) or 1
    tot_i = sum(v["Instructions Executed"] for v in agg.values()) or 1
    print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
    print(f"{'file:line':28s} {'samp%':>6s} {'inst%':>6s}  top stalls | source")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
        st = sorted(((k[6:], x) for k, x in v.items() if k.startswith("stall_") and x > 0), key=lambda t: -t[1])[:3]
        sts = " ".join(f"{k}:{100 * x / tot_s:.1f}" for k, x in st)
        print(f"{key[0] + ':' + str(key[1]):28s} {100 * v['# Samples'] / tot_s:6.1f} {100 * v['Instructions Executed'] / tot_i:6.1f}  {sts:40s} | {src[key].strip()[:90]}")


if __name__ == "__main__":
    main()
