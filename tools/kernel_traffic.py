#!/usr/bin/env python
"""Per-kernel table of one frame from an ncu launch list, and the DRAM traffic per launch of every C-ABI entry
point -- the numbers bench.py reports as `roofline.traffic` / `kernels[*].dram_traffic_per_launch`.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \\
        --csv --log-file gpurun_out/launches_<mode>.csv python tools/profile_forward.py --iters 3 --precision <mode>
    python tools/kernel_traffic.py MODE=LAUNCHES.csv [MODE=LAUNCHES.csv ...] --json profiles/r02_kernel_traffic.json \\
        --md profiles/r02_launches_summary.md

ncu serialises the kernels and flushes caches between them: times are cold-cache (compare SHARES), DRAM bytes
are the cold-cache traffic of each kernel."""
import collections
import csv
import json
import sys

# kernel-name prefix -> entry point whose launch it belongs to
ENTRY = [
    ("k_tca_tile", "mssvt_block_attention_tc"), ("k_tca_merge", "mssvt_block_attention_tc"),
    ("k_tca_plan", "mssvt_attention_tiles"),
    ("k_ffn_tc", "mssvt_ffn_tc"), ("k_ffn", "mssvt_ffn"),
    ("k_block_geometry", "mssvt_block_geometry"), ("k_geo_", "mssvt_block_geometry"),
    ("k_tcc_plan", "mssvt_compress_tiles"), ("k_tcc_", "mssvt_compress_attention_tc"),
    ("k_tc_linear", "mssvt_compress_attention_tc"),
    ("k_layernorm", "mssvt_layernorm"), ("k_window_rows", "mssvt_window_rows"),
    ("k_win_", "mssvt_window_partition"), ("k_fill_i32", "mssvt_window_partition"),
    ("k_grid_", "mssvt_grid_index_build"), ("k_count_samples", "mssvt_count_samples"),
    ("k_prefix_small", "mssvt_count_samples"), ("k_world_coords", "mssvt_voxel_world_coords"),
    ("k_scan_", "mssvt_exclusive_scan"), ("k_query_src", "mssvt_query_src"),
    ("k_dense_scatter", "mssvt_dense_scatter"), ("k_block_attention", "mssvt_block_attention"),
    ("k_compress_attention", "mssvt_compress_attention"),
]


def short(name):
    return name.split("(")[0].replace("void ", "").replace("mssvt::", "")[:60]


def entry_of(name):
    for prefix, entry in ENTRY:
        if name.startswith(prefix):
            return entry
    return None


def read(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ii, ki, mi, vi, ui = H.index("ID"), H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit")
    launches = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        rec = launches.setdefault(r[ii], {"name": short(r[ki]), "us": 0.0, "bytes": 0.0})
        v = float(r[vi].replace(",", ""))
        if r[mi] == "gpu__time_duration.sum":
            rec["us"] = v / {"ns": 1000.0, "us": 1.0, "ms": 1e-3}.get(r[ui], 1000.0)
        elif r[mi].startswith("dram__bytes"):
            rec["bytes"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1.0)
    seq = list(launches.values())
    first = [i for i, l in enumerate(seq) if l["name"].startswith("k_count_samples")]
    return seq[first[-2]:first[-1]]          # the last complete frame


def main():
    args = sys.argv[1:]
    out_json = args[args.index("--json") + 1] if "--json" in args else None
    out_md = args[args.index("--md") + 1] if "--md" in args else None
    traffic, md = {}, []
    for a in args:
        if "=" not in a:
            continue
        mode, path = a.split("=", 1)
        frame = read(path)
        per_kernel, per_entry = collections.OrderedDict(), {}
        for l in frame:
            k = per_kernel.setdefault(l["name"], [0, 0.0, 0.0])
            k[0] += 1; k[1] += l["us"]; k[2] += l["bytes"]
            e = entry_of(l["name"])
            if e:
                per_entry.setdefault(e, [0.0, set()])[0] += l["bytes"]
        # launches per frame of an entry point = count of its leading kernel
        lead = {}
        for l in frame:
            e = entry_of(l["name"])
            if e and e not in lead:
                lead[e] = l["name"]
        traffic[mode] = {e: int(v[0] / max(1, per_kernel[lead[e]][0])) for e, v in per_entry.items()}
        tot = sum(v[1] for v in per_kernel.values())
        md.append(f"## {mode}: one frame = {len(frame)} launches, {tot:.1f} us (serialised, cold caches); "
                  "DRAM bytes = dram__bytes_read.sum + dram__bytes_write.sum\n")
        md.append("| kernel | launches | us | share | DRAM MB | GB/s |\n|---|---:|---:|---:|---:|---:|")
        for n, (c, t, b) in sorted(per_kernel.items(), key=lambda kv: -kv[1][1]):
            md.append(f"| `{n}` | {c} | {t:.1f} | {100 * t / tot:.1f} % | {b / 1e6:.1f} | {b / t / 1e3 if t else 0:.0f} |")
        md.append("")
    if out_json:
        json.dump(traffic, open(out_json, "w"), indent=1, sort_keys=True)
    text = "\n".join(md)
    if out_md:
        open(out_md, "w").write("# ncu launch lists, one frame of 150 k voxels (tools/kernel_traffic.py)\n\n" + text + "\n")
    print(text)


if __name__ == "__main__":
    main()
