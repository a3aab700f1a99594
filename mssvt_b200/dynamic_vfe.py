"""DynamicVFE: points -> (voxel_features, voxel_coords), the producer of the backbone's inputs.

Mirror of pcdet/models/backbones_3d/vfe/dynamic_vfe.py:12-130 (same constructor, config keys, state-dict
names `pfn.{i}.0` Linear / `pfn.{i}.1` BatchNorm1d, batch_dict keys), running on mssvt_vfe_voxelize /
mssvt_vfe_features (csrc/vfe.cu): occupancy bitmap + popcount scan instead of torch.unique, atomics instead of
torch_scatter, BatchNorm (eval) folded into the linear layers.  Inference only (training raises: BatchNorm
batch statistics over the points are not implemented).  One or two PFN layers.
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

    @torch.no_grad()
    def forward(self, batch_dict, **kwargs):
        if self.training:
            raise RuntimeError("mssvt_b200.DynamicVFE runs in eval mode only (BatchNorm batch statistics over the "
                               "points are not implemented)")
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
