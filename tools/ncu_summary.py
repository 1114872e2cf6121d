#!/usr/bin/env python
"""Summarise an .ncu-rep: per kernel launch the headline metrics (raw page) and, from the source
page, the opcode mix and the SASS instructions with the most stall samples.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 20] [--kernel REGEX] > profiles/xxx.txt
"""
import argparse
import collections
import csv
import io
import re
import subprocess

RAW = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def run(args):
    return subprocess.run(["ncu", "-i", args.rep, *args.extra], capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--top", type=int, default=18)
    ap.add_argument("--kernel", default=".")
    a = ap.parse_args()
    a.extra = ["--page", "raw", "--csv"]
    rows = list(csv.reader(io.StringIO(run(a))))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    launches = rows[2:]
    a.extra = ["--page", "source", "--csv"]
    src = run(a)
    # split the source page per kernel
    blocks, cur = [], None
    for line in src.splitlines():
        if line.startswith('"Kernel Name"'):
            cur = [line]
            blocks.append(cur)
        elif cur is not None:
            cur.append(line)
    for li, r in enumerate(launches):
        name = r[idx["Kernel Name"]]
        if not re.search(a.kernel, name):
            continue
        print("=" * 100)
        print(f"launch {li}: {name[:140]}")
        for m in RAW:
            if m in idx:
                print(f"  {m:85s} {r[idx[m]]} {units[idx[m]]}")
        if 2 * li < len(blocks):
            rd = list(csv.reader(io.StringIO("\n".join(blocks[2 * li][1:]))))
            h = {k: i for i, k in enumerate(rd[0])}
            ins = rd[1:]
            ops = collections.Counter()
            tot = 0
            for x in ins:
                try:
                    n = int(x[h["Instructions Executed"]])
                except ValueError:
                    continue
                sass = x[h["Source"]].strip()
                sass = re.sub(r"^@!?U?P\d+\s+", "", sass)
                op = sass.split()[0].split(".")[0] if sass else "?"
                ops[op] += n
                tot += n
            print(f"  -- warp-instructions executed: {tot}; opcode mix (top 16):")
            print("     " + ", ".join(f"{k}:{v * 100 / max(tot, 1):.1f}%" for k, v in ops.most_common(16)))
            key = h["# Samples"]
            ins2 = [x for x in ins if x[key].isdigit()]
            tots = sum(int(x[key]) for x in ins2) or 1
            print(f"  -- top {a.top} SASS by stall samples (of {tots}):")
            stall_cols = [k for k in rd[0] if k.startswith("stall_") and "Not Issued" not in k]
            for x in sorted(ins2, key=lambda x: -int(x[key]))[: a.top]:
                st = sorted(((int(x[h[c]]), c[6:]) for c in stall_cols if x[h[c]].isdigit()), reverse=True)[:2]
                print(f"     {int(x[key]) * 100 / tots:5.1f}%  {x[h['Source']].strip()[:70]:70s} {st}")


if __name__ == "__main__":
    main()
