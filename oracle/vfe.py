"""CPU restatement of DynamicVFE.forward in eval mode (pcdet/models/backbones_3d/vfe/dynamic_vfe.py:71-130).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this.
Third-party arithmetic restated: torch.unique(sorted, return_inverse, return_counts) and torch_scatter's
scatter_mean / scatter_max along dim 0 (torch_scatter is not installed; dynamic_vfe.py:5-9 imports it
lazily).  Pinned by oracle/pin_vfe_against_reference.py, which runs the reference's unmodified module with
a torch_scatter stand-in built on torch.Tensor.index_reduce_ and compares.
"""
import torch
import torch.nn.functional as F


def scatter_mean(src, index, num):
    out = torch.zeros((num, src.shape[1]), dtype=src.dtype)
    out.index_add_(0, index, src)
    cnt = torch.zeros(num, dtype=src.dtype).index_add_(0, index, torch.ones(index.shape[0], dtype=src.dtype))
    return out / cnt.clamp(min=1).unsqueeze(1)


def scatter_max(src, index, num):
    out = torch.full((num, src.shape[1]), float("-inf"), dtype=src.dtype)
    return out.index_reduce_(0, index, src, "amax", include_self=True)


def dynamic_vfe_forward(state, points, batch_size, voxel_size, grid_size, point_cloud_range,
                        num_point_features, with_cluster_center=True, with_voxel_center=True,
                        with_distance=False, eps=1e-5):
    """state: {'pfn.{i}.0.weight', 'pfn.{i}.0.bias', 'pfn.{i}.1.{weight,bias,running_mean,running_var}'}.
    points (P, 1 + num_point_features) [batch_idx, x, y, z, ...] -> (voxel_features (V, C), voxel_coords (V, 4)
    int32 [b, z, y, x]) in ascending (b, x, y, z) key order (dynamic_vfe.py:85-119)."""
    points = points.float()
    vs = torch.tensor(voxel_size, dtype=torch.float32)
    lo = torch.tensor(point_cloud_range[0:3], dtype=torch.float32)
    gs = torch.tensor(grid_size, dtype=torch.int32)
    pc = torch.floor((points[:, 1:4] - lo) / vs).int()                                   # :85
    mask = ((pc >= 0) & (pc < gs)).all(dim=1)                                            # :86
    points, pc = points[mask], pc[mask]
    sxyz, syz, sz = grid_size[0] * grid_size[1] * grid_size[2], grid_size[1] * grid_size[2], grid_size[2]
    merge = points[:, 0].int() * sxyz + pc[:, 0] * syz + pc[:, 1] * sz + pc[:, 2]        # :89-92
    unq, inv, _ = torch.unique(merge, return_inverse=True, return_counts=True)           # :93
    V = unq.shape[0]
    xyz = points[:, 1:4]
    feats = [points[:, 1:num_point_features + 1]]                                        # :96
    if with_cluster_center:
        feats.append(xyz - scatter_mean(xyz, inv, V)[inv])                               # :98-100
    if with_voxel_center:
        offset = torch.tensor([voxel_size[i] / 2 + point_cloud_range[i] for i in range(3)], dtype=torch.float32)
        feats.append(xyz - (pc * vs + offset.view(1, 3)))                                # :102-104
    if with_distance:
        feats.append(torch.norm(xyz, p=2, dim=1, keepdim=True))
    x = torch.cat(feats, dim=-1)
    n_layers = len({k.split(".")[1] for k in state if k.startswith("pfn.")})
    for i in range(n_layers):                                                            # :124-130
        x = F.linear(x, state["pfn.%d.0.weight" % i], state["pfn.%d.0.bias" % i])
        x = F.batch_norm(x, state["pfn.%d.1.running_mean" % i], state["pfn.%d.1.running_var" % i],
                         state["pfn.%d.1.weight" % i], state["pfn.%d.1.bias" % i], False, 0.0, eps)
        x = F.relu(x)
        if i < n_layers - 1:
            x = torch.cat((x, scatter_max(x, inv, V)[inv]), dim=-1)
    voxel_fea = scatter_max(x, inv, V)                                                    # :108
    unq = unq.int()
    coords = torch.stack((unq // sxyz, (unq % sxyz) // syz, (unq % syz) // sz, unq % sz), dim=1)
    return voxel_fea.contiguous(), coords[:, [0, 3, 2, 1]].contiguous().int()             # :111-116
