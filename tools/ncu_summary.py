"""Summarise one `ncu --set full` report into the small CSV kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_ascent_ipm_kernel_metrics.csv

Adds the dynamic instruction mix by opcode (from the SASS source page of the same report).
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sass__inst_executed_global_loads", "sass__inst_executed_global_stores",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def run(cmd):
    return subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(run(f"ncu -i {rep} --page raw --csv").splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    lines = [("metric", "unit", "value"), ("kernel", "", d["Kernel Name"][1])]
    for k in KEEP:
        if k in d:
            lines.append((k, d[k][0], d[k][1]))
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            lines.append((h, d[h][0], d[h][1]))
    src = list(csv.reader(run(f"ncu -i {rep} --page source --csv --print-source sass").splitlines()))
    sh = src[1]
    ix = {h: i for i, h in enumerate(sh)}
    mix, tot = collections.Counter(), 0.0
    for r in src[2:]:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
        if not m:
            continue
        n = float(r[ix["Instructions Executed"]] or 0)
        mix[m.group(2)] += n
        tot += n
    for op, n in mix.most_common(16):
        lines.append((f"inst_mix.{op}", "% of warp instructions", f"{100 * n / tot:.2f}"))
    with open(out, "w", newline="") as f:
        csv.writer(f).writerows(lines)
    print(f"wrote {out}: {len(lines) - 1} rows")


if __name__ == "__main__":
    main()
