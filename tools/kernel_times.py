#!/usr/bin/env python
"""Per-entry-point CUDA-event times of eager backbone forwards over one synthetic frame (the breakdown bench.py
prints, without the rest of the bench): the quick A/B tool for kernel experiments.
usage: [MSSVT_B200_LIB=variant.so] python tools/kernel_times.py [--precision tf32] [--n 150000] [--iters 10]"""
import argparse
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mssvt_b200 import _lib  # noqa: E402
from mssvt_b200.config import s0_model_cfg  # noqa: E402
from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer  # noqa: E402
from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL, synth_frame  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--n", type=int, default=150000)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--patterns", default="1,1,1")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    cfg = s0_model_cfg(cbs_patterns=tuple(int(v) for v in a.patterns.split(",")))
    cfg["PRECISION"] = a.precision
    torch.manual_seed(0)
    model = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE)).cuda().eval()
    frames = []
    for s in range(4):   # rotate frames: 4 x 41 MB of inputs + intermediates > L2
        f, c = synth_frame(s, a.n)
        frames.append((torch.from_numpy(f).cuda(), torch.from_numpy(c).cuda()))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    whole = 0.0
    with torch.no_grad():
        for it in range(a.iters + 3):
            f, c = frames[it % len(frames)]
            _lib.PROFILE = [] if it >= 3 else None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sp = model({"voxel_features": f, "voxel_coords": c, "batch_size": 1})["encoded_spconv_tensor"]
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                whole += e0.elapsed_time(e1)
                for name, s0, s1 in _lib.PROFILE:
                    tot[name] += s0.elapsed_time(s1)
                    cnt[name] += 1
    _lib.PROFILE = None
    lib = os.path.basename(_lib.LIB_PATH)
    parts = ["%s=%.1f" % (k.replace("mssvt_", ""), 1000 * v / cnt[k]) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]]
    print("[%s %s %s] eager frame %.3f ms | us per launch: %s" % (a.tag, lib, a.precision, whole / a.iters, " ".join(parts)))


if __name__ == "__main__":
    main()
