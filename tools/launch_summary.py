#!/usr/bin/env python
"""Aggregate one frame of an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.
usage: python tools/launch_summary.py LAUNCHES.csv"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    seq = [(r[ki].split("(")[0].replace("void ", "").replace("mssvt::", "")[:70],
            float(r[vi].replace(",", "")) / (1000 if r[ui] == "ns" else 1)) for r in rows[hdr + 1:] if len(r) > vi]
    first = [i for i, (n, _) in enumerate(seq) if "k_count_samples" in n]
    s, e = first[-2], first[-1]  # the last complete frame
    agg = collections.OrderedDict()
    for n, t in seq[s:e]:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"one frame: {e - s} launches, {tot:.1f} us (serialised, cold caches)")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:8.1f} us  {100 * t / tot:5.1f}%  x{c:<3d} {n}")


if __name__ == "__main__":
    main()
