#!/usr/bin/env python
"""Aggregate an ncu report's source page per CUDA source line.
usage: python tools/ncu_hot_lines.py REPORT.ncu-rep KERNEL_REGEX [TOP]"""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    path, hdr = None, None
    inst, smp, src = defaultdict(float), defaultdict(float), {}
    stall = defaultdict(lambda: defaultdict(float))
    cur = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            path, hdr = r[1].split("/")[-1], None
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 2:
            continue
        d = dict(zip(hdr, r))
        # two 'Source' columns: first = cuda line text (only on the first sass row of a line)
        if r[0]:
            cur = (path, int(r[0]))
            src[cur] = r[1].strip()
        if cur is None:
            continue
        def f(k):
            try:
                return float(d.get(k) or 0)
            except ValueError:
                return 0.0
        inst[cur] += f("Instructions Executed")
        smp[cur] += f("# Samples")
        for k in hdr:
            if k.startswith("stall_") and "Not Issued" not in k:
                stall[cur][k] += f(k)
    ti, ts = sum(inst.values()) or 1, sum(smp.values()) or 1
    print("total warp-instructions %.0f, samples %.0f" % (ti, ts))
    for k in sorted(inst, key=lambda k: -smp[k])[:top]:
        tops = sorted(stall[k].items(), key=lambda kv: -kv[1])[:3]
        print("%5.1f%% smp %5.1f%% inst  %s:%d  %-80s  %s" % (
            100 * smp[k] / ts, 100 * inst[k] / ti, k[0], k[1], src.get(k, "")[:80],
            " ".join("%s=%d" % (a[6:], b) for a, b in tops if b)))


if __name__ == "__main__":
    main()
