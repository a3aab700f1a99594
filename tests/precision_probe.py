"""Max / RMS error of both precision modes against the CPU oracle on two synthetic frames (test infrastructure:
imports oracle/).  usage: python tests/precision_probe.py"""
import sys, torch
sys.path.insert(0, ".")
from oracle import backbone as orc
from mssvt_b200.config import s0_model_cfg
from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer
from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL, synth_frame
for seed, n, crop in ((0, 4000, 0.17), (3, 20000, 0.38)):
    feats, coords = synth_frame(seed, n, crop=crop)
    feats, coords = torch.from_numpy(feats), torch.from_numpy(coords)
    cfg = s0_model_cfg()
    torch.manual_seed(0)
    model = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE))
    state = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        want = orc.backbone_forward(state, cfg, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE), feats, coords, 1)
    model = model.cuda().eval()
    for prec in ("fp32", "tf32x3", "tf32"):
        model.set_precision(prec)
        with torch.no_grad():
            sp = model({"voxel_features": feats.cuda(), "voxel_coords": coords.cuda().float(), "batch_size": 1})["encoded_spconv_tensor"]
        d = (sp.features.cpu() - want.features).abs()
        s = want.features.abs().max().item()
        print(f"N={n} {prec}: max|d|/max|ref| = {d.max().item() / s:.2e}, rms|d|/rms|ref| = {d.pow(2).mean().sqrt().item() / want.features.pow(2).mean().sqrt().item():.2e}")
