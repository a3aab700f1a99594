#!/usr/bin/env python
"""Where do the ragged and the padded training paths differ?  Runs both on a batch with an empty sample (the configuration
of tests/test_gpu_training.py::test_ragged_training_path_batch_with_empty_sample_and_overflow) and prints the location and
size of the largest differences of the input gradient, with the samples' first voxels (the rows every padded key position
aliases, quirk Q1: they collect a gradient term from most windows of the sample) listed separately."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mssvt_b200 import mssvt_backbone  # noqa: E402
from mssvt_b200.config import s0_model_cfg  # noqa: E402
from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer  # noqa: E402
from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL, synth_frame  # noqa: E402


def run(path, cfg, state, feats, coords, batch):
    mssvt_backbone.TRAIN_PATH = path
    model = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE))
    model.load_state_dict(state)
    model = model.cuda().train()
    for m in model.modules():
        if hasattr(m, "drop_prob"):
            m.drop_prob = 0.0
    x = feats.cuda().requires_grad_(True)
    sp = model({"voxel_features": x, "voxel_coords": coords.cuda().float(), "batch_size": batch})["encoded_spconv_tensor"]
    (sp.features ** 2).sum().backward()
    return sp.features.detach(), x.grad.clone(), {n: p.grad.clone() for n, p in model.named_parameters()}


def main():
    feats, coords = synth_frame(5, 3000, batch_size=3, crop=0.3)
    keep = coords[:, 0] != 1
    feats, coords = torch.from_numpy(feats[keep]), torch.from_numpy(coords[keep])
    torch.manual_seed(5)
    cfg = s0_model_cfg(cbs_patterns=(2, 1, 0))
    state = {k: v.clone() for k, v in MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE)).state_dict().items()}
    fa, ga, pa = run("ragged", cfg, state, feats, coords, 3)
    fb, gb, pb = run("padded", cfg, state, feats, coords, 3)
    fc, gc, pc = run("padded", cfg, state, feats, coords, 3)
    first = [0, int((coords[:, 0] == 0).sum())]
    d = (ga - gb).abs()
    print("features: max diff %.3g of max %.3g" % ((fa - fb).abs().max().item(), fb.abs().max().item()))
    print("input gradient: max |ragged - padded| %.4g at row %d (max |grad| %.4g); padded run-to-run %.4g"
          % (d.max().item(), int(d.max(1)[0].argmax()), gb.abs().max().item(), (gb - gc).abs().max().item()))
    for r in first:
        print("  first voxel of a sample, row %d: max |grad| %.4g, max diff %.4g (run-to-run of the padded path %.4g)"
              % (r, gb[r].abs().max().item(), d[r].max().item(), (gb[r] - gc[r]).abs().max().item()))
    mask = torch.ones(d.shape[0], dtype=torch.bool, device=d.device)
    mask[first] = False
    print("  all other rows: max |grad| %.4g, max diff %.4g" % (gb[mask].abs().max().item(), d[mask].max().item()))
    worst = max(((pa[n] - pb[n]).abs().max().item() / max(pb[n].abs().max().item(), 1e-6), n) for n in pb)
    print("parameter gradients: worst relative difference %.3g (%s)" % worst)


if __name__ == "__main__":
    main()
