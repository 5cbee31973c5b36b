#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel (and grid size) launches, total and mean time.
  python tools/launch_summary.py gpurun_out/launches.csv [--by-grid]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    by_grid = "--by-grid" in sys.argv
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("wb200::", "").replace("<unnamed>::", "")
        key = (name, r["Grid Size"]) if by_grid else (name,)
        rows.append((key, us))
    agg = collections.OrderedDict()
    for k, us in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values()) or 1.0
    print("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {' '.join(k)} | {a[0]} | {a[1] / 1e3:.3f} | {a[1] / a[0]:.2f} | {a[1] / tot:.3f} |")
    print(f"\ntotal kernel time {tot / 1e3:.2f} ms over {len(rows)} launches")


if __name__ == "__main__":
    main()
