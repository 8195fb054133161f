"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python scripts/summarize_ncu.py launches <launches.csv> <out.md>
    python scripts/summarize_ncu.py full <report.ncu-rep> <out.md>
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, start = r, i
            break
    ki, mi = h.index("Kernel Name"), h.index("Metric Value")
    d = defaultdict(list)
    for r in rows[start + 1:]:
        if len(r) > mi:
            d[r[ki]].append(float(r[mi].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list ({path})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are "
                "cold-cache and serialised: compare shares, not absolutes.\n\n| kernel | launches | mean us | share of captured time |\n|---|---:|---:|---:|\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k[:90]}` | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {100 * sum(v) / tot:.1f} % |\n")


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({path})\n\n")
        for r in rows[2:]:
            f.write(f"## {r[hdr.index('Kernel Name')][:100]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    f.write(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")
            f.write("\nwarp stall reasons (average warps stalled per issue-active cycle, > 0.3 only): ")
            st = []
            for i, k in enumerate(hdr):
                if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(r[i] or 0) > 0.3:
                    st.append(k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "") + f" {float(r[i]):.2f}")
            f.write(", ".join(st) + "\n\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
