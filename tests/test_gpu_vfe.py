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
    out_t = vfe.train()({"points": same, "batch_size": 1})       # training mode: batch statistics, differentiable
    assert out_t["voxel_features"].requires_grad and out_t["voxel_coords"].shape == (1, 4)


def test_dynamic_vfe_training_mode_batch_statistics_and_gradients():
    """train(): BatchNorm1d uses the statistics of the in-range points (dynamic_vfe.py:57-66, 124-130) and the
    module is differentiable; checked against the same mathematics in plain PyTorch on the CPU (torch.unique +
    scatter), coordinates bit-exact, features and parameter gradients within 1e-4"""
    g = torch.Generator().manual_seed(3)
    n = 20000
    xyz = (torch.rand((n, 3), generator=g) - 0.5) * torch.tensor([40.0, 40.0, 5.0]) + torch.tensor([0.0, 0.0, 1.0])
    xyz[:200] += 500.0                                                       # some points outside the range
    points = torch.cat([torch.randint(0, 2, (n, 1), generator=g).float(), xyz, torch.rand((n, 2), generator=g)], 1)
    from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL
    torch.manual_seed(0)
    vfe = DynamicVFE(AttrDict(NUM_FILTERS=[32, 64]), 5, list(S0_VOXEL), list(S0_GRID), list(S0_RANGE))
    ref = DynamicVFE(AttrDict(NUM_FILTERS=[32, 64]), 5, list(S0_VOXEL), list(S0_GRID), list(S0_RANGE))
    ref.load_state_dict(vfe.state_dict())
    vfe = vfe.cuda().train()
    out = vfe({"points": points.cuda(), "batch_size": 2})
    (out["voxel_features"] ** 2).mean().backward()
    # the reference mathematics on the CPU
    vs, lo = torch.tensor(S0_VOXEL), torch.tensor(S0_RANGE[:3])
    pc = torch.floor((points[:, 1:4] - lo) / vs).int()
    m = ((pc >= 0) & (pc < torch.tensor(S0_GRID))).all(1)
    p, pc = points[m], pc[m]
    key = ((p[:, 0].long() * S0_GRID[0] + pc[:, 0]) * S0_GRID[1] + pc[:, 1]) * S0_GRID[2] + pc[:, 2]
    unq, inv = torch.unique(key, return_inverse=True)
    V = unq.shape[0]
    smax = lambda t: t.new_zeros((V, t.shape[1])).scatter_reduce(0, inv.unsqueeze(1).expand_as(t), t, "amax", include_self=False)
    mean = torch.zeros(V, 3).index_add_(0, inv, p[:, 1:4]) / torch.bincount(inv, minlength=V).unsqueeze(1)
    off = torch.tensor([S0_VOXEL[i] / 2 + S0_RANGE[i] for i in range(3)])
    x = torch.cat([p[:, 1:6], p[:, 1:4] - mean[inv], p[:, 1:4] - (pc * vs + off)], 1)
    ref.train()
    for i, blk in enumerate(ref.pfn):
        x = blk(x)
        if i == 0:
            x = torch.cat((x, smax(x)[inv]), 1)
    want = smax(x)
    (want ** 2).mean().backward()
    coords = torch.stack([unq // (S0_GRID[0] * S0_GRID[1] * S0_GRID[2]), unq % S0_GRID[2], (unq // S0_GRID[2]) % S0_GRID[1],
                          (unq // (S0_GRID[1] * S0_GRID[2])) % S0_GRID[0]], 1).int()
    assert torch.equal(out["voxel_coords"].cpu(), coords)
    assert (out["voxel_features"].detach().cpu() - want.detach()).abs().max().item() <= 1e-4 * want.abs().max().item()
    gscale = max(b.grad.abs().max().item() for b in ref.parameters())   # (a Linear bias in front of a BatchNorm has
    for (n1, a), (_, b) in zip(vfe.named_parameters(), ref.named_parameters()):   #  a mathematically zero gradient)
        assert (a.grad.cpu() - b.grad).abs().max().item() <= 1e-4 * gscale, n1
    # the running statistics were updated like nn.BatchNorm1d does
    assert torch.allclose(vfe.pfn[0][1].running_mean.cpu(), ref.pfn[0][1].running_mean, atol=1e-5)
