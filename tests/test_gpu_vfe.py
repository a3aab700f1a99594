"""GPU: DynamicVFE (csrc/vfe.cu through the module mirror) against golden vectors produced by the reference's
unmodified module (oracle/pin_vfe_against_reference.py) and against the CPU restatement on larger clouds.
Coordinates bit-exact; features within 1e-5 (cluster means are summed with float atomics)."""
import os

import numpy as np
import pytest
import torch

from oracle import vfe as orc_vfe
from mssvt_b200.config import AttrDict
from mssvt_b200.dynamic_vfe import DynamicVFE

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-5


def build(blob):
    cfg = AttrDict(NUM_FILTERS=[int(v) for v in blob["filters"]])
    vfe = DynamicVFE(cfg, 5, blob["voxel_size"].tolist(), blob["grid_size"].tolist(), blob["pc_range"].tolist())
    state = {k[6:]: torch.from_numpy(blob[k]) for k in blob.files if k.startswith("state/")}
    vfe.load_state_dict(state, strict=True)
    return vfe.cuda().eval(), state


@pytest.mark.parametrize("name", ["vfe_s0_b2_p6000", "vfe_two_layer_b3_p4000"])
def test_dynamic_vfe_matches_reference_golden(name):
    blob = np.load(os.path.join(GOLDEN, name + ".npz"))
    vfe, _ = build(blob)
    out = vfe({"points": torch.from_numpy(blob["points"]).cuda(), "batch_size": int(blob["batch_size"])})
    assert torch.equal(out["voxel_coords"].cpu(), torch.from_numpy(blob["voxel_coords"]))
    ref = torch.from_numpy(blob["voxel_features"])
    err = (out["voxel_features"].cpu() - ref).abs().max().item()
    assert err <= TOL * max(ref.abs().max().item(), 1.0), err
    assert vfe.get_output_feature_dim() == ref.shape[1]


def test_dynamic_vfe_vs_oracle_on_a_dense_cloud_and_feeds_the_backbone():
    """180 k points on the S0 grid (several points per voxel), then straight into the backbone"""
    from mssvt_b200.config import s0_model_cfg
    from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer
    from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL
    g = torch.Generator().manual_seed(5)
    n = 180000
    centres = torch.rand((6000, 3), generator=g) * torch.tensor([120.0, 120.0, 4.0]) + torch.tensor([-60.0, -60.0, -1.8])
    xyz = centres[torch.randint(0, 6000, (n,), generator=g)] + (torch.rand((n, 3), generator=g) - 0.5) * torch.tensor([2.0, 2.0, 0.8])
    points = torch.cat([torch.zeros(n, 1), xyz, torch.rand((n, 2), generator=g)], 1)
    torch.manual_seed(0)
    vfe = DynamicVFE(AttrDict(NUM_FILTERS=[64]), 5, list(S0_VOXEL), list(S0_GRID), list(S0_RANGE)).eval()
    with torch.no_grad():
        for m in vfe.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.3); m.running_var.uniform_(0.5, 1.5)
    state = {k: v.clone() for k, v in vfe.state_dict().items()}
    want_f, want_c = orc_vfe.dynamic_vfe_forward(state, points, 1, list(S0_VOXEL), list(S0_GRID), list(S0_RANGE), 5)
    out = vfe.cuda()({"points": points.cuda(), "batch_size": 1})
    assert torch.equal(out["voxel_coords"].cpu(), want_c)
    err = (out["voxel_features"].cpu() - want_f).abs().max().item()
    assert err <= TOL * want_f.abs().max().item(), err
    pv = out["point_voxel"].cpu()
    assert int((pv >= 0).sum()) > 0.9 * n and int(pv.max()) == want_c.shape[0] - 1
    cfg = s0_model_cfg()
    cfg["PRECISION"] = "tf32"
    backbone = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE)).cuda().eval()
    with torch.no_grad():
        sp = backbone(out)["encoded_spconv_tensor"]
    assert torch.isfinite(sp.features).all() and sp.features.shape[1] == 64


def test_dynamic_vfe_edge_cases():
    from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL
    vfe = DynamicVFE(AttrDict(NUM_FILTERS=[64]), 5, list(S0_VOXEL), list(S0_GRID), list(S0_RANGE)).cuda().eval()
    far = torch.tensor([[0, 500.0, 0.0, 0.0, 0.1, 0.2], [0, 0.0, 0.0, 99.0, 0.1, 0.2]]).cuda()   # all outside
    out = vfe({"points": far, "batch_size": 1})
    assert out["voxel_features"].shape == (0, 64) and out["voxel_coords"].shape == (0, 4)
    same = torch.tensor([[0, 1.0, 1.0, 0.0, 0.3, 0.4]] * 7).cuda()                                  # one voxel, 7 points
    out = vfe({"points": same, "batch_size": 1})
    assert out["voxel_coords"].shape == (1, 4) and bool((out["point_voxel"] == 0).all())
    with pytest.raises(RuntimeError):
        vfe.train()({"points": same, "batch_size": 1})
