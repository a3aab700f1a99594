#!/usr/bin/env python
"""BASELINE config 1: ONE MixedScaleSparseTransformerBlock forward, batch 1, 20 000 non-empty voxels (cropped
region, Waymo-like density), C = 64, S0 windows / heads, eval -- on one B200 in every precision mode, with the
CPU path (the oracle's restatement of the reference block) timed beside it on the host cores, and the two
outputs compared.  Prints ONE JSON line; --out also writes it to a file.

    python benchmarks/config1_block.py [--out profiles/r02_config1_block.json]
"""
import argparse
import json
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mssvt_b200.config import block_cfg  # noqa: E402
from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer, MixedScaleSparseTransformerBlock  # noqa: E402
from mssvt_b200.mssvt_utils import SparseTensor  # noqa: E402
from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL, synth_frame  # noqa: E402

N = 20000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--iters", type=int, default=50)
    args = ap.parse_args()
    cfg = block_cfg()
    torch.manual_seed(0)
    blk = MixedScaleSparseTransformerBlock(cfg, 64, 128, 64, [2, 2], drop_path=0.0, window_size=cfg.window_size,
                                           cbs_pattern=1).eval()
    P = {k: v.detach().clone() for k, v in blk.state_dict().items()}
    frames = [synth_frame(s, N, crop=0.38) for s in range(4)]       # rotated: each forward sees new coordinates

    # ---- CPU path: the oracle's block (PyTorch-CPU fp32 + OpenMP C kernels), all host threads
    from oracle import backbone as orc
    from oracle import ops as orc_ops
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    orc_ops.lib().orc_set_num_threads(cores)

    def cpu_block(f, c):
        sp = orc.Frame(torch.from_numpy(f), torch.from_numpy(c), list(S0_GRID), list(S0_VOXEL), list(S0_RANGE), 1, 400000)
        with torch.no_grad():
            return orc.block_forward(P, dict(cfg), sp).features

    want = cpu_block(*frames[0])
    cpu_t = []
    for i in range(5):
        t0 = time.perf_counter()
        cpu_block(*frames[i % 4])
        cpu_t.append(time.perf_counter() - t0)
    cpu_ms = statistics.median(cpu_t) * 1e3

    # ---- GPU: the block through the module API (geometry + LayerNorm + tile attention + FFN every call)
    dev = torch.device("cuda", 0)
    blk = blk.to(dev)
    gpu = [(torch.from_numpy(f).to(dev), torch.from_numpy(c).to(dev)) for f, c in frames]

    def gpu_block(i):
        f, c = gpu[i % 4]
        sp = SparseTensor(f, c, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE), 1, 400000)
        return blk(sp).features

    modes = {}
    for mode, tol in (("fp32", 1e-4), ("tf32x3", 1e-4), ("tf32", 2e-3)) + ((("bf16", 2e-2),) if "bf16" in MixedScaleSparseTransformer.PRECISIONS else ()):
        blk.precision = mode
        with torch.no_grad():
            got = gpu_block(0)
            err = (got.cpu() - want).abs().max().item() / want.abs().max().item()
            for i in range(5):
                gpu_block(i)
            torch.cuda.synchronize()
            ts = []
            for i in range(args.iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                gpu_block(i)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        modes[mode] = {"ms_per_block": ms, "voxels_per_s": N / (ms * 1e-3), "max_rel_err_vs_cpu": err, "tolerance": tol,
                       "ok": err <= tol}
    line = {"metric": "mssvt_block_fwd_voxels_per_s", "unit": "voxels/s", "config": {
                "workload": "single MixedScaleSparseTransformerBlock forward (geometry included), batch 1, %d voxels on a "
                            "0.38 crop of the S0 grid, C=64, ff=128, heads [2,2], windows 3^3/5^3, K=32, pattern 1" % N,
                "launch": "eager (module API), median of %d, coordinates change every call" % args.iters},
            "gpu": modes,
            "cpu_baseline": {"value": N / (cpu_ms * 1e-3), "unit": "voxels/s", "ms_per_block": cpu_ms, "cores": cores,
                             "kind": "port", "sample": "median of 5 block forwards of the CPU oracle"},
            "data": "synthetic"}
    text = json.dumps(line)
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
