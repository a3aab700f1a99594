#!/usr/bin/env python
"""BASELINE config 3: hash build, hash query + chessboard sampling, K/V feature gather, swept over
N in {10k, 30k, 100k, 300k, 1M} voxels -- achieved GB/s (algorithmic bytes / CUDA-event time) against
the measured HBM peak.  Prints one JSON line per (kernel, N) and writes a markdown table.

    python benchmarks/microbench.py [--out profiles/r01_microbench.md]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mssvt_b200 import mssvt_ops  # noqa: E402
from mssvt_b200.mssvt_backbone import vox_query_table  # noqa: E402
from mssvt_b200.mssvt_utils import SparseTensor  # noqa: E402
from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL, synth_frame  # noqa: E402
from mssvt_b200 import pointnet2_utils  # noqa: E402
from mssvt_b200._lib import call, ptr, stream  # noqa: E402
# benchmark side only: the reference's own CUDA kernels (oracle/_ref, compiled unmodified from /root/reference),
# timed next to ours as "reference CUDA on B200" (BASELINE.md section 5).  Never used by the product.
from oracle import ref_kernels as ref  # noqa: E402

HAVE_REF = ref.available()


def timed(fn, iters=20, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()  # 256 MB write: evicts the 126 MB L2 between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_microbench.md"))
    ap.add_argument("--sizes", default="10000,30000,100000,300000,1000000")
    args = ap.parse_args()
    peak = 6558.4
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    dev = torch.device("cuda", 0)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in vox_query_table([3, 3, 3], [5, 5, 5]).items()}
    rows = []
    for n in [int(v) for v in args.sizes.split(",")]:
        crop = min(1.0, max(0.1, (n / 150000.0) ** 0.5)) if n < 150000 else 1.0
        feats, coords = synth_frame(7, n, crop=crop)
        feats, coords = torch.from_numpy(feats).to(dev), torch.from_numpy(coords).to(dev)
        H = 400000 if n <= 200000 else 1 << (2 * n - 1).bit_length()
        cnt = torch.tensor([n], dtype=torch.int32, device=dev)
        grid = [S0_GRID[i] // 3 for i in range(3)]
        # (i) hash build (reference contract) and grid-index build (fused path)
        dt = timed(lambda: mssvt_ops.build_hash_table(1, H, S0_GRID, coords, cnt), flush=flush)
        dr = timed(lambda: ref.build_hash_table(1, H, S0_GRID, coords, cnt), iters=5, flush=flush) if HAVE_REF else None
        rows.append(("hash build (k % H table)", n, 16 * n + 8 * n + 8 * H, dt, dr))
        sp = SparseTensor(feats, coords, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE), 1, H)

        def grid_build():
            sp._derived = {}
            sp.grid_index()
        dt = timed(grid_build, flush=flush)
        words = S0_GRID[0] * S0_GRID[1]
        rows.append(("grid index build (bitmap + rank)", n, 2 * 16 * n + 3 * 8 * words + 4 * n, dt, None))
        # (ii) hash query = chessboard gather (op level, padded outputs mandated by the API)
        table = mssvt_ops.build_hash_table(1, H, S0_GRID, coords, cnt)
        win, _ = mssvt_ops.get_non_empty_window_center([3, 3, 3], max(90000, n), 1, H, grid, coords)
        W = win.shape[0]
        dt = timed(lambda: mssvt_ops.gather_two_window_voxels(S0_GRID, [3, 3, 3], 12, 3, 27, 125, t["odd"], t["even"],
                                                              t["win1"], t["win2"], win, table), flush=flush)
        dr = timed(lambda: ref.gather_two_window_voxels(S0_GRID, [3, 3, 3], 12, 3, 27, 125, t["odd"], t["even"], t["win1"],
                                                        t["win2"], win, table), iters=5, flush=flush) if HAVE_REF else None
        rows.append(("chessboard gather (hash probes, padded lists)", n, 16 * W + 16 * W * (12 + 3 + 27 + 125) + 12 * 125, dt, dr))
        # window partition (op level: window hash + first-occurrence numbering)
        dt = timed(lambda: mssvt_ops.window_partition_device([3, 3, 3], max(90000, n), 1, H, grid, coords), flush=flush)
        dr = timed(lambda: ref.get_non_empty_window_center([3, 3, 3], max(90000, n), 1, H, grid, coords), iters=5,
                   flush=flush) if HAVE_REF else None
        rows.append(("window partition (window hash + list)", n, 16 * n + 16 * W + 8 * H + 4 * n, dt, dr))
        # FPS of 32 keys out of the 125-slot win2 lists (op level, float offsets like the reference passes them)
        off = torch.randint(-2, 3, (W, 125, 3), device=dev).float()
        off[:, 14:] *= (torch.rand((W, 111, 1), device=dev) < 0.12)     # ~14 real voxels per list, the rest padding
        dt = timed(lambda: pointnet2_utils.farthest_point_sample(off, 32), flush=flush)
        dr = timed(lambda: ref.farthest_point_sample(off, 32), iters=5, flush=flush) if HAVE_REF else None
        rows.append(("farthest point sampling 125 -> 32 per window", n, 12 * 125 * W + 4 * 32 * W, dt, dr))
        # (ii') the same sampling on the fused path: grid-index probes + FPS + key maps + three-NN in one kernel
        from mssvt_b200.config import block_cfg
        from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformerBlock
        cfgb = block_cfg()
        blk = MixedScaleSparseTransformerBlock(cfgb, 64, 128, 64, [2, 2], drop_path=0.0, window_size=cfgb.window_size).to(dev).eval()
        blk.max_num_wins = max(90000, n)
        sp.grid_index()
        blk._windows(sp)

        def fused_geometry():
            c = sp._cache()
            for k in [k for k in c if isinstance(k, tuple) and k[0] == "geo"]:
                del c[k]
            blk.geometry(sp)
        dt = timed(fused_geometry, flush=flush)
        rows.append(("chessboard sampling, fused (grid index + FPS + key maps + 3-NN)", n, 16 * W + W * 1100 + n, dt, None))
        # (iii) K/V feature gather: idx = 32 key rows per window, C = 64 and the 32-channel slice
        idx = torch.randint(0, n, (W, 32), dtype=torch.int32, device=dev)
        idx[torch.rand((W, 32), device=dev) < 0.1] = -1
        wcnt = torch.tensor([W], dtype=torch.int32, device=dev)
        valid = int((idx >= 0).sum())
        for C in (64, 32):
            f = feats[:, :C].contiguous()
            dt = timed(lambda: mssvt_ops.grouping_operation(f, cnt, idx, wcnt), flush=flush)
            dr = timed(lambda: ref.grouping_operation(f, cnt, idx, wcnt), iters=5, flush=flush) if HAVE_REF else None
            rows.append(("K/V feature gather C=%d ns=32" % C, n, 4 * W * 32 + 4 * C * valid + 4 * C * W * 32, dt, dr))
    # (iv) DynamicVFE: points -> voxels (bitmap + scan voxelisation, PFN, per-voxel max); N = points here
    from mssvt_b200.config import AttrDict
    from mssvt_b200.dynamic_vfe import DynamicVFE
    vfe = DynamicVFE(AttrDict(NUM_FILTERS=[64]), 5, list(S0_VOXEL), list(S0_GRID), list(S0_RANGE)).to(dev).eval()
    words = S0_GRID[0] * S0_GRID[1]
    for n in (180000, 1000000):
        g = torch.Generator().manual_seed(n)
        centres = torch.rand((n // 30, 3), generator=g) * torch.tensor([140.0, 140.0, 5.0]) + torch.tensor([-70.0, -70.0, -1.9])
        xyz = centres[torch.randint(0, n // 30, (n,), generator=g)] + (torch.rand((n, 3), generator=g) - 0.5) * torch.tensor([2.0, 2.0, 0.8])
        pts = torch.cat([torch.zeros(n, 1), xyz, torch.rand((n, 2), generator=g)], 1).to(dev)
        batch = {"points": pts, "batch_size": 1}
        v = vfe(dict(batch))["voxel_features"].shape[0]
        dt = timed(lambda: vfe(dict(batch)), flush=flush)
        # points read twice (voxelise, PFN), bitmap + counts + scan, per-voxel sums, per-voxel output rows
        rows.append(("DynamicVFE points -> %d voxels (voxelise + PFN 11->64 + max, incl. one host sync)" % v, n,
                     2 * 24 * n + 4 * n + 5 * 4 * words + 16 * v + 16 * v + 4 * 64 * v, dt, None))
    lines = ["| kernel | N voxels | algorithmic MB | time us | GB/s | of measured HBM peak (%.0f GB/s) | reference CUDA kernel on the "
             "same B200, us | speed-up |" % peak, "|---|---:|---:|---:|---:|---:|---:|---:|"]
    for name, n, b, dt, dr in rows:
        gbs = b / dt / 1e9
        print(json.dumps({"kernel": name, "n_voxels": n, "bytes": b, "us": dt * 1e6, "gbs": gbs, "frac": gbs / peak,
                          "reference_cuda_us": dr * 1e6 if dr else None}))
        lines.append("| %s | %d | %.1f | %.1f | %.0f | %.2f | %s | %s |" % (
            name, n, b / 1e6, dt * 1e6, gbs, gbs / peak, "%.1f" % (dr * 1e6) if dr else "-", "%.1fx" % (dr / dt) if dr else "-"))
    with open(args.out, "w") as f:
        f.write("# Round 2 micro-benchmarks (BASELINE config 3)\n\n`python benchmarks/microbench.py` on one B200; median of 20 "
                "(reference kernels: of 5), CUDA events, L2 flushed (256 MB write) between iterations; bytes are algorithmic "
                "(DESIGN.md section 4).  Reference column: the reference's own .cu files compiled unmodified for sm_100a "
                "(oracle/_ref), outputs allocated and pre-filled ON THE DEVICE (the reference's Python fills them on the "
                "CPU and copies them over PCIe, which is not included here).\n\n")
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
