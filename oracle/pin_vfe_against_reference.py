"""Pin oracle/vfe.py against the reference's own, unmodified DynamicVFE and write golden vectors.

TEST INFRASTRUCTURE, build container only (needs /root/reference).  The reference module
(pcdet/models/backbones_3d/vfe/dynamic_vfe.py) is loaded as it is; `torch_scatter` (not installed) is
replaced by a stand-in with scatter_mean / scatter_max built on index_add_ / index_reduce_, `.cuda()` maps
to the CPU.  Usage:  python -m oracle.pin_vfe_against_reference [--write]
"""
import argparse
import importlib.util
import os
import sys
import types

import numpy as np
import torch

from . import vfe as orc_vfe

REF = os.environ.get("MSSVT_REFERENCE", "/root/reference")
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _load_reference():
    torch.Tensor.cuda = lambda self, *a, **k: self
    ts = types.ModuleType("torch_scatter")

    def scatter_mean(src, index, dim=0):
        return orc_vfe.scatter_mean(src, index, int(index.max()) + 1)

    def scatter_max(src, index, dim=0):
        return orc_vfe.scatter_max(src, index, int(index.max()) + 1), None

    ts.scatter_mean, ts.scatter_max = scatter_mean, scatter_max
    sys.modules["torch_scatter"] = ts
    base = os.path.join(REF, "pcdet/models/backbones_3d/vfe")
    pkg = types.ModuleType("refvfe")
    pkg.__path__ = [base]
    sys.modules["refvfe"] = pkg
    for name in ("vfe_template", "dynamic_vfe"):
        spec = importlib.util.spec_from_file_location("refvfe." + name, os.path.join(base, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["refvfe." + name] = mod
        spec.loader.exec_module(mod)
    return sys.modules["refvfe.dynamic_vfe"].DynamicVFE


class _Cfg(dict):
    def __getattr__(self, k):
        return self[k]


def synth_points(seed, n, batch_size, pc_range):
    g = torch.Generator().manual_seed(seed)
    lo, hi = torch.tensor(pc_range[0:3]), torch.tensor(pc_range[3:6])
    r = torch.rand((n, 3), generator=g)
    # clustered: a few hundred centres with many points each (so that voxels hold several points) + 5 % outside
    centres = lo + (hi - lo) * torch.rand((max(n // 20, 1), 3), generator=g)
    xyz = centres[torch.randint(0, centres.shape[0], (n,), generator=g)] + (r - 0.5) * torch.tensor([1.2, 1.2, 0.6])
    out = torch.rand(n, generator=g) < 0.05
    xyz[out] = lo + (hi - lo) * (torch.rand((int(out.sum()), 3), generator=g) * 1.4 - 0.2)
    b = torch.randint(0, batch_size, (n, 1), generator=g).float()
    extra = torch.rand((n, 2), generator=g)
    return torch.cat([b, xyz, extra], 1)


def case(name, seed, n, batch_size, filters, voxel_size, grid_size, pc_range, write):
    DynamicVFE = _load_reference()
    torch.manual_seed(seed)
    cfg = _Cfg(NUM_FILTERS=filters, WITH_CLUSTER_CENTER=True, WITH_VOXEL_CENTER=True, WITH_DISTANCE=False)
    cfg.get = lambda k, d=None: dict.get(cfg, k, d)
    ref = DynamicVFE(cfg, 5, voxel_size, grid_size, pc_range).eval()
    with torch.no_grad():
        for m in ref.modules():
            if isinstance(m, torch.nn.BatchNorm1d):   # non-trivial running statistics / affine
                m.running_mean.normal_(0, 0.3); m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.7, 1.3); m.bias.normal_(0, 0.2)
    points = synth_points(seed, n, batch_size, pc_range)
    with torch.no_grad():
        out = ref({"batch_size": batch_size, "points": points.clone()})
    state = {k: v.clone() for k, v in ref.state_dict().items()}
    got_f, got_c = orc_vfe.dynamic_vfe_forward(state, points, batch_size, voxel_size, grid_size, pc_range, 5)
    assert torch.equal(got_c, out["voxel_coords"].int()), "voxel coordinates differ"
    err = (got_f - out["voxel_features"]).abs().max().item()
    print("%s: %d points -> %d voxels, max|oracle - reference| = %.2e" % (name, n, got_c.shape[0], err))
    assert err <= 1e-6 * max(out["voxel_features"].abs().max().item(), 1.0)
    if write:
        blob = {"points": points.numpy(), "voxel_features": out["voxel_features"].numpy(),
                "voxel_coords": out["voxel_coords"].int().numpy(), "batch_size": np.int32(batch_size),
                "voxel_size": np.float64(voxel_size), "grid_size": np.int64(grid_size),
                "pc_range": np.float64(pc_range), "filters": np.int64(filters)}
        blob.update({"state/" + k: v.numpy() for k, v in state.items()})
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **blob)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--write", action="store_true")
    a = ap.parse_args()
    S0 = ([0.32, 0.32, 0.1875], [468, 468, 32], [-74.88, -74.88, -2, 74.88, 74.88, 4])
    case("vfe_s0_b2_p6000", 1, 6000, 2, [64], *S0, a.write)
    case("vfe_two_layer_b3_p4000", 2, 4000, 3, [32, 64], [0.4, 0.4, 0.25], [60, 50, 16], [-12, -10, -2, 12, 10, 2], a.write)


if __name__ == "__main__":
    main()
