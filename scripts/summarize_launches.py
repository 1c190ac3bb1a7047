#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python scripts/summarize_launches.py profiles/<file>.csv"""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[row["Metric Unit"]]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{sum(v[0] for v in agg.values())} launches, {tot / 1e3:.3f} ms total (cold-cache, serialised: compare shares)")
    print(f"{'us':>10} {'share':>6} {'n':>5}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:10.1f} {100 * v[1] / tot:5.1f}% {v[0]:5d}  {k[:100]}")


if __name__ == "__main__":
    main(sys.argv[1])
