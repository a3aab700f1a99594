#!/usr/bin/env python
"""A few eager backbone forwards over one synthetic frame: the short command profilers wrap
(ncu --set full -k regex:<kernel> ..., or a -DMSSVT_TRACE build that prints per-phase clock timelines).
usage: python tools/profile_forward.py [--precision tf32] [--n 150000] [--iters 3] [--patterns 1,1,1]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mssvt_b200.config import s0_model_cfg  # noqa: E402
from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer  # noqa: E402
from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL, synth_frame  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--n", type=int, default=150000)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--patterns", default="1,1,1")
    a = ap.parse_args()
    cfg = s0_model_cfg(cbs_patterns=tuple(int(v) for v in a.patterns.split(",")))
    cfg["PRECISION"] = a.precision
    torch.manual_seed(0)
    model = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE)).cuda().eval()
    f, c = synth_frame(0, a.n)
    f, c = torch.from_numpy(f).cuda(), torch.from_numpy(c).cuda()
    with torch.no_grad():
        for _ in range(a.iters):
            sp = model({"voxel_features": f, "voxel_coords": c, "batch_size": 1})["encoded_spconv_tensor"]
            sp.dense()
    torch.cuda.synchronize()
    print("rows", sp.features.shape[0])


if __name__ == "__main__":
    main()
