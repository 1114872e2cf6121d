#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name:
launch count, total and mean duration, share of the summed GPU time.

    python tools/launch_summary.py gpurun_out/launches_bench.csv > profiles/rN_launches_<what>.txt
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("tvdk::", "")
    return name[:110]


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    rd = csv.DictReader(lines)
    agg = collections.OrderedDict()
    total = 0.0
    n = 0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
        n += 1
    print(f"# {path}: {n} launches, {total / 1e6:.3f} ms summed GPU time (per-launch times under ncu are cold-cache and")
    print("# serialised: compare SHARES with the live CUDA-event numbers of bench.py, not absolutes)")
    print(f"{'launches':>9} {'total ms':>11} {'mean us':>10} {'share':>7}  kernel")
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{c:9d} {ns / 1e6:11.3f} {ns / c / 1e3:10.2f} {100 * ns / total:6.2f}%  {k}")


if __name__ == "__main__":
    main()
