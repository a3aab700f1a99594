import sys, torch
sys.path.insert(0, ".")
from mssvt_b200.config import s0_model_cfg
from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer
from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL, synth_frame
cfg = s0_model_cfg(); cfg["PRECISION"] = "tf32"
torch.manual_seed(0)
model = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE)).cuda().eval()
f, c = synth_frame(100, 150000)
f, c = torch.from_numpy(f).cuda(), torch.from_numpy(c).cuda()
with torch.no_grad():
    for i in range(3):
        model({"voxel_features": f, "voxel_coords": c, "batch_size": 1})
        torch.cuda.synchronize()
