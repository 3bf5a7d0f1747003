#!/usr/bin/env python
"""Condense an Nsight Compute report (.ncu-rep, captured on the GPU box with `ncu --set full`)
into the short text summary that is committed under profiles/.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_fused_cfg2.txt [--note "..."]

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU)."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__bytes_read.sum.per_second", "dram read rate"),
    ("dram__bytes_write.sum.per_second", "dram write rate"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of ncu peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_fp64.sum", "FP64 pipe instructions"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__cluster_size", "cluster size"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = ""
    if "--note" in sys.argv:
        note = sys.argv[sys.argv.index("--note") + 1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    lines = [f"# ncu summary of {rep.split('/')[-1]}", ""]
    if note:
        lines += [note, ""]
    for r in rows[hdr_i + 2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append(f"## {d['Kernel Name']}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k, name in KEYS:
            if k in d and d[k] != "":
                lines.append(f"  {name:34s} {d[k]} {u.get(k, '')}   [{k}]")
        st = [(float(d[k]), k) for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k and d[k] not in ("", "n/a")]
        st.sort(reverse=True)
        lines.append("  warp stall reasons (warps per issue-active cycle):")
        for v, k in st[:8]:
            lines.append(f"    {k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''):22s} {v:.2f}")
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
