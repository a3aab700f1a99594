#!/usr/bin/env python
"""Per-kernel table of ONE training step from an ncu launch list: launches, time, DRAM bytes, achieved GB/s and the
fraction of the measured HBM peak (MEASURED_PEAKS.json), our kernels (mssvt::) and the library kernels side by side.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/train_launches.csv python benchmarks/train_step.py --steps 1 --warmup 2
    python tools/train_kernel_table.py gpurun_out/train_launches.csv > profiles/<tag>_train_kernels.md

The last step of the run is taken: from the last-but-one `k_count_samples` launch (first kernel of a forward) to the
last one would miss the final step, so the list is cut at the LAST `k_count_samples` and runs to the end (forward +
backward + AdamW of the last step).  ncu serialises kernels and flushes caches: times are cold-cache."""
import collections
import csv
import json
import os
import sys


def read(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ii, ki, mi, vi, ui = (H.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    launches = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        rec = launches.setdefault(r[ii], {"name": r[ki], "us": 0.0, "bytes": 0.0})
        v = float(r[vi].replace(",", ""))
        if r[mi] == "gpu__time_duration.sum":
            rec["us"] = v / {"ns": 1000.0, "us": 1.0, "ms": 1e-3}.get(r[ui], 1000.0)
        elif r[mi].startswith("dram__bytes"):
            rec["bytes"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1.0)
    seq = list(launches.values())
    starts = [i for i, l in enumerate(seq) if "k_count_samples" in l["name"]]
    return seq[starts[-1]:]


def short(name):
    n = name.replace("void ", "")
    ours = "mssvt::" in n
    n = n.replace("mssvt::", "")
    cut = n.find("(")
    return (n[:cut] if cut > 0 else n)[:72], ours


def main():
    step = read(sys.argv[1])
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peak = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"]
    agg = collections.OrderedDict()
    for l in step:
        name, ours = short(l["name"])
        a = agg.setdefault(name, [0, 0.0, 0.0, ours])
        a[0] += 1; a[1] += l["us"]; a[2] += l["bytes"]
    tot = sum(a[1] for a in agg.values())
    ours_t = sum(a[1] for a in agg.values() if a[3])
    print("one training step (forward + backward + AdamW), 150 k-voxel S0 frame: %d launches, %.2f ms serialised under ncu "
          "(cold caches); hand-written kernels %.2f ms = %.0f %%, library / torch kernels %.2f ms\n"
          % (len(step), tot / 1e3, ours_t / 1e3, 100 * ours_t / tot, (tot - ours_t) / 1e3))
    print("| kernel | ours | launches | us total | share | DRAM MB / launch | GB/s | of HBM peak (%.0f GB/s) |" % peak)
    print("|---|---|---:|---:|---:|---:|---:|---:|")
    for name, (c, us, by, ours) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        gbs = by / us / 1e3 if us > 0 else 0.0
        print("| `%s` | %s | %d | %.1f | %.1f %% | %.1f | %.0f | %.2f |"
              % (name, "yes" if ours else "", c, us, 100 * us / tot, by / c / 1e6, gbs, gbs / peak))


if __name__ == "__main__":
    main()
