#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys


def main(path, top=16):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e6 if unit == "ns" else v / 1e3 if unit == "us" else v
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:72]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"| `{k}` | {v[0]} | {v[1]:.3f} | {v[1]/v[0]:.4f} | {100*v[1]/tot:.1f}% |")
    print(f"\nTotal {tot:.1f} ms over {sum(v[0] for v in agg.values())} launches.")


if __name__ == "__main__":
    main(sys.argv[1])
