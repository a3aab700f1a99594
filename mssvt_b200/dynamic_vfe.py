"""DynamicVFE: points -> (voxel_features, voxel_coords), the producer of the backbone's inputs.

Mirror of pcdet/models/backbones_3d/vfe/dynamic_vfe.py:12-130 (same constructor, config keys, state-dict
names `pfn.{i}.0` Linear / `pfn.{i}.1` BatchNorm1d, batch_dict keys), running on mssvt_vfe_voxelize /
mssvt_vfe_features (csrc/vfe.cu): occupancy bitmap + popcount scan instead of torch.unique, atomics instead of
torch_scatter, BatchNorm (eval) folded into the linear layers.  Training mode (BatchNorm batch statistics,
gradients) keeps the CUDA voxelisation and runs the PFN as differentiable torch ops.  One or two PFN layers.
"""
import torch
import torch.nn as nn

from ._lib import call, host_floats, ptr, stream


class DynamicVFE(nn.Module):
    def __init__(self, model_cfg, num_point_features, voxel_size, grid_size, point_cloud_range, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_point_features_in = num_point_features
        self.voxel_size = [float(v) for v in voxel_size]
        self.grid_size = [int(v) for v in grid_size]
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.with_cluster_center = model_cfg.get('WITH_CLUSTER_CENTER', True)
        self.with_voxel_center = model_cfg.get('WITH_VOXEL_CENTER', True)
        self.with_distance = model_cfg.get('WITH_DISTANCE', False)
        in_c = num_point_features + 3 * bool(self.with_cluster_center) + 3 * bool(self.with_voxel_center) \
            + bool(self.with_distance)
        self.in_channels = in_c
        filters = list(model_cfg.get('NUM_FILTERS', [64, 128]))
        if len(filters) not in (1, 2):
            raise NotImplementedError("DynamicVFE: one or two PFN layers")
        self.num_point_features = filters[-1]
        self.pfn = nn.ModuleList([])
        for out_c in filters:
            self.pfn.append(nn.Sequential(nn.Linear(in_c, out_c), nn.BatchNorm1d(out_c), nn.ReLU(inplace=True)))
            in_c = out_c * 2
        # voxel-centre offset as the reference builds it: python doubles, then fp32 (dynamic_vfe.py:30-33)
        self.centre_offset = [self.voxel_size[i] / 2 + self.point_cloud_range[i] for i in range(3)]

    def get_output_feature_dim(self):
        return self.num_point_features

    def _folded(self):
        """(w', b') per layer with the eval-mode BatchNorm folded in; cached until a parameter changes"""
        tag = tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))
        hit = self.__dict__.get("_fold_cache")
        if hit is None or hit[0] != tag:
            out = []
            for lin, bn, _ in self.pfn:
                s = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
                out.append(((lin.weight.detach() * s.unsqueeze(1)).float().contiguous(),
                            ((lin.bias.detach() - bn.running_mean) * s + bn.bias.detach()).float().contiguous()))
            hit = self.__dict__["_fold_cache"] = (tag, out)
        return hit[1]

    def _voxelize(self, points, B):
        """mssvt_vfe_voxelize: -> (point_voxel (P), coords (V, 4), xyz_sum (V, 4) = sum of xyz | #points, V)"""
        P, stride, dev = points.shape[0], points.shape[1], points.device
        gx, gy, gz = self.grid_size
        words = int(call("mssvt_vfe_bitmap_words", B, gx, gy, gz))
        i32 = dict(dtype=torch.int32, device=dev)
        bitmap, counts, base = torch.empty(words, **i32), torch.empty(words, **i32), torch.empty(words + 1, **i32)
        work = torch.empty((words + 1) // 1024 + 2, **i32)
        point_voxel, coords = torch.empty(max(P, 1), **i32), torch.empty((max(P, 1), 4), **i32)
        xyz_sum = torch.empty((max(P, 1), 4), dtype=torch.float32, device=dev)
        call("mssvt_vfe_voxelize", P, ptr(points), stride, B, gx, gy, gz, host_floats(self.voxel_size),
             host_floats(self.point_cloud_range[0:3]), ptr(bitmap), ptr(counts), ptr(base), ptr(work), ptr(point_voxel),
             ptr(coords), ptr(xyz_sum), stream())
        V = int(base[words].item())
        return point_voxel[:P], coords[:V], xyz_sum[:V], V

    def _forward_train(self, batch_dict):
        """Training mode (dynamic_vfe.py:71-130 with BatchNorm1d batch statistics over the in-range points): the
        voxelisation -- the sort-free replacement of torch.unique -- is the CUDA kernel, the per-point features,
        Linear + BatchNorm + ReLU and the per-voxel max run as differentiable torch ops (scatter_reduce amax)."""
        points = batch_dict['points'].float().contiguous()
        B = int(batch_dict['batch_size'])
        with torch.no_grad():
            point_voxel, coords, xyz_sum, V = self._voxelize(points, B)
        keep = point_voxel >= 0                                                    # :86 (points outside the grid drop out)
        inv = point_voxel[keep].long()
        pts = points[keep]
        xyz = pts[:, 1:4]
        feats = [pts[:, 1:self.num_point_features_in + 1]]
        if self.with_cluster_center:
            mean = xyz_sum[:, 0:3] / xyz_sum[:, 3:4].clamp(min=1.0)
            feats.append(xyz - mean[inv])
        if self.with_voxel_center:
            cell = coords[inv][:, [3, 2, 1]].float()                               # [b, z, y, x] -> x, y, z cell indices
            centre = cell * xyz.new_tensor(self.voxel_size) + xyz.new_tensor(self.centre_offset)
            feats.append(xyz - centre)
        if self.with_distance:
            feats.append(torch.norm(xyz, p=2, dim=1, keepdim=True))
        x = torch.cat(feats, dim=-1)
        vmax = lambda t: t.new_zeros((V, t.shape[1])).scatter_reduce(0, inv.unsqueeze(1).expand_as(t), t, "amax",
                                                                      include_self=False)
        for i, blk in enumerate(self.pfn):
            x = blk(x)                                                             # Linear + BatchNorm1d (batch stats) + ReLU
            if i < len(self.pfn) - 1:
                x = torch.cat((x, vmax(x)[inv]), dim=-1)
        batch_dict['voxel_features'] = vmax(x).contiguous()
        batch_dict['voxel_coords'] = coords
        batch_dict['point_voxel'] = point_voxel
        return batch_dict

    def forward(self, batch_dict, **kwargs):
        if not batch_dict['points'].is_cuda:
            raise RuntimeError("mssvt_b200 operators run on CUDA tensors only; there is no CPU path")
        if self.training:
            return self._forward_train(batch_dict)
        with torch.no_grad():
            return self._forward_eval(batch_dict)

    def _forward_eval(self, batch_dict):
        points = batch_dict['points']                     # (P, 1 + F) [batch_idx, x, y, z, ...]
        if not points.is_cuda:
            raise RuntimeError("mssvt_b200 operators run on CUDA tensors only; there is no CPU path")
        points = points.float().contiguous()
        P, stride, B, dev = points.shape[0], points.shape[1], int(batch_dict['batch_size']), points.device
        gx, gy, gz = self.grid_size
        words = int(call("mssvt_vfe_bitmap_words", B, gx, gy, gz))
        i32 = dict(dtype=torch.int32, device=dev)
        bitmap, counts, base = torch.empty(words, **i32), torch.empty(words, **i32), torch.empty(words + 1, **i32)
        work = torch.empty((words + 1) // 1024 + 2, **i32)
        point_voxel, coords = torch.empty(max(P, 1), **i32), torch.empty((max(P, 1), 4), **i32)
        xyz_sum = torch.empty((max(P, 1), 4), dtype=torch.float32, device=dev)   # (also: #points per voxel)
        vs, lo = host_floats(self.voxel_size), host_floats(self.point_cloud_range[0:3])
        call("mssvt_vfe_voxelize", P, ptr(points), stride, B, gx, gy, gz, vs, lo, ptr(bitmap), ptr(counts), ptr(base),
             ptr(work), ptr(point_voxel), ptr(coords), ptr(xyz_sum), stream())
        layers = self._folded()
        (w0, b0), (w1, b1) = layers[0], (layers[1] if len(layers) > 1 else (None, None))
        c0, c1 = w0.shape[0], (w1.shape[0] if w1 is not None else 0)
        scratch = torch.empty((max(P, 1), c0), dtype=torch.float32, device=dev) if c1 else None
        out = torch.empty((max(P, 1), c1 or c0), dtype=torch.float32, device=dev)
        call("mssvt_vfe_features", P, ptr(points), stride, self.num_point_features_in,
             int(bool(self.with_cluster_center)), int(bool(self.with_voxel_center)), int(bool(self.with_distance)),
             vs, lo, host_floats(self.centre_offset), ptr(point_voxel), ptr(xyz_sum), max(P, 1), ptr(w0), ptr(b0), c0,
             ptr(w1), ptr(b1), c1, ptr(scratch), ptr(out), stream())
        V = int(base[words].item())                        # the reference's torch.unique synchronises here as well
        batch_dict['voxel_features'] = out[:V]
        batch_dict['voxel_coords'] = coords[:V]
        batch_dict['point_voxel'] = point_voxel[:P]        # (extra) voxel row of every point, -1 outside the range
        return batch_dict
