"""SparseTensor and MixedScaleAttention of pcdet/models/model_utils/mssvt_utils.py.

`SparseTensor` keeps the reference's attributes (features, indices, spatial_shape, batch_size,
voxel_size, point_cloud_range, hash_size, gather_dict, map_table) and `.dense()`; construction
builds the voxel hash on the device without the per-sample `.item()` loop
(mssvt_utils.py:33-48).  `MixedScaleAttention` owns the same parameters under the same names
(to_qs / to_kvs / projs) so reference checkpoints load; its arithmetic runs inside the fused
window kernels (mssvt_block_attention / mssvt_compress_attention), which is where the blocks
call it from.
"""
import torch
from torch import nn

from . import mssvt_ops
from ._lib import call, ptr, stream, host_floats


def sample_counts(indices, batch_size):
    """(counts (B), start (B+1)) int32 on the device, no host sync."""
    counts = torch.empty(batch_size, dtype=torch.int32, device=indices.device)
    start = torch.empty(batch_size + 1, dtype=torch.int32, device=indices.device)
    call("mssvt_count_samples", indices.shape[0], batch_size, ptr(indices), ptr(counts), ptr(start),
         stream())
    return counts, start


class SparseTensor(object):
    """mssvt_utils.py:21-62."""

    def __init__(self, features, indices, spatial_shape, voxel_size, point_cloud_range, batch_size,
                 hash_size, map_table=None, gather_dict=None):
        self.features = features            # (N, C), samples contiguous
        self.indices = indices              # (N, 4) int32 [b, z, y, x]
        self.spatial_shape = spatial_shape  # [x, y, z]
        self.batch_size = batch_size
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.hash_size = hash_size
        self.gather_dict = gather_dict
        self._derived = {}                  # per-coordinate-set caches (counts, world xyz, geometry)
        self._map_table = map_table         # reference-contract hash table, built on first use

    @property
    def map_table(self):
        """(B, hash_size, 2) table of the reference (mssvt_utils.py:31).  The fused blocks look
        voxels up through grid_index() instead, so the table is only built when somebody asks
        for it (e.g. to call mssvt_ops.gather_two_window_voxels)."""
        if self._map_table is None:
            self._map_table = self.build_map_table()
        return self._map_table

    @map_table.setter
    def map_table(self, value):
        self._map_table = value

    @map_table.deleter
    def map_table(self):
        self._map_table = None

    # ---- coordinate-derived state, recomputed only when `indices` is replaced
    def _cache(self):
        key = (self.indices.data_ptr(), tuple(self.indices.shape), tuple(self.spatial_shape))
        if self._derived.get("key") != key:
            self._derived = {"key": key}
        return self._derived

    def sample_counts(self):
        c = self._cache()
        if "counts" not in c:
            c["counts"], c["start"] = sample_counts(self.indices, self.batch_size)
        return c["counts"], c["start"]

    def world_coords(self):
        """(N, 3) voxel centres, with_coords of mssvt_backbone.py:132-137."""
        c = self._cache()
        if "xyz" not in c:
            xyz = torch.empty((self.indices.shape[0], 3), dtype=torch.float32, device=self.indices.device)
            call("mssvt_voxel_world_coords", self.indices.shape[0], ptr(self.indices),
                 host_floats(self.voxel_size), host_floats(self.point_cloud_range[0:3]), ptr(xyz), stream())
            c["xyz"] = xyz
        return c["xyz"]

    def grid_index(self):
        """(cells, vals): occupancy-bitmap + rank lookup structure of the fused path."""
        c = self._cache()
        if "grid" not in c:
            x, y, z = (int(v) for v in self.spatial_shape)
            dev = self.indices.device
            words = call("mssvt_grid_index_words", x, y, z, self.batch_size)
            cells = torch.empty((words, 2), dtype=torch.int32, device=dev)
            vals = torch.empty(max(self.indices.shape[0], 1), dtype=torch.int32, device=dev)
            work = torch.empty((words + 1023) // 1024 + 1, dtype=torch.int32, device=dev)
            _, start = self.sample_counts()
            call("mssvt_grid_index_build", x, y, z, self.indices.shape[0], self.batch_size,
                 ptr(self.indices), ptr(start), ptr(cells), ptr(vals), ptr(work), stream())
            c["grid"] = (cells, vals)
        return c["grid"]

    @torch.no_grad()
    def build_map_table(self):
        counts, _ = self.sample_counts()
        return mssvt_ops.build_hash_table(self.batch_size, self.hash_size, self.spatial_shape,
                                          self.indices, counts)

    def dense(self, channels_first=True):
        x, y, z = (int(v) for v in self.spatial_shape)
        feats = self.features.float().contiguous()
        C = feats.shape[1]
        out = torch.empty((self.batch_size, C, z, y, x), dtype=torch.float32, device=feats.device)
        call("mssvt_dense_scatter", feats.shape[0], None, self.batch_size, C, z, y, x, ptr(feats),
             ptr(self.indices), ptr(out), stream())
        return out if channels_first else out.permute(0, 2, 3, 4, 1).contiguous()


class MixedScaleAttention(nn.Module):
    """Parameter container with the reference's layout (mssvt_utils.py:65-86): head group g owns
    channel slice [c_g, c_{g+1}) of width per_head_dim * num_heads[g] and its own q / kv / proj."""

    def __init__(self, embed_dim, num_heads, dropout=0.):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = list(num_heads)
        self.num_head_groups = len(num_heads)
        self.tot_num_heads = sum(num_heads)
        assert self.embed_dim % self.tot_num_heads == 0
        self.per_head_dim = self.embed_dim // self.tot_num_heads
        self.group_c_idx = [self.per_head_dim * sum(num_heads[:i + 1]) for i in range(self.num_head_groups)]
        self.scale_dims = [self.per_head_dim * h for h in num_heads]
        self.to_qs = nn.ModuleList([nn.Linear(sd, sd) for sd in self.scale_dims])
        self.to_kvs = nn.ModuleList([nn.Linear(sd, 2 * sd) for sd in self.scale_dims])
        self.scale = self.per_head_dim ** -0.5
        self.attn_drop = nn.Dropout(dropout)
        self.projs = nn.ModuleList([nn.Linear(sd, sd) for sd in self.scale_dims])
        self.proj_drop = nn.Dropout(dropout)
        self.dropout = dropout

    def forward(self, *args, **kwargs):
        raise RuntimeError(
            "MixedScaleAttention is evaluated inside the fused window kernels "
            "(mssvt_block_attention / mssvt_compress_attention); call the enclosing "
            "MixedScaleSparseTransformerBlock instead. There is no dense PyTorch path.")
