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


class _PinnedPairs:
    """Pinned int32[2] buffers for asynchronous (row count, dropped windows) readbacks.  A buffer is handed
    out again only after the event recorded behind its last copy has completed, so any number of
    un-materialised outputs may be in flight (each keeps its own buffer until it is read)."""

    def __init__(self):
        self.free, self.busy = [], []

    def get(self):
        still = []
        for buf, ev in self.busy:
            if ev is None or ev.query():
                self.free.append(buf)
            else:
                still.append((buf, ev))
        self.busy = still
        return self.free.pop() if self.free else torch.empty(2, dtype=torch.int32).pin_memory()

    def release(self, buf, event):
        self.busy.append((buf, event))


_PINNED = _PinnedPairs()


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
        self._lazy = None                   # (features_cap, indices_cap, count_dev) until first access
        self._features = features           # (N, C), samples contiguous
        self._indices = indices             # (N, 4) int32 [b, z, y, x]
        self.spatial_shape = spatial_shape  # [x, y, z]
        self.batch_size = batch_size
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.hash_size = hash_size
        self.gather_dict = gather_dict
        self._derived = {}                  # per-coordinate-set caches (counts, world xyz, geometry)
        self._count_host = self._ready = None
        self._map_table = map_table         # reference-contract hash table, built on first use

    # A compress block produces one row per non-empty window; that count lives on the device.  The
    # rows are kept at capacity and sliced to the exact count the first time somebody looks at
    # .features / .indices (one 4-byte D2H copy), so the forward itself never waits for the host.
    def set_lazy_rows(self, features_cap, indices_cap, count_dev):
        self._lazy = (features_cap, indices_cap, count_dev)
        self._features = self._indices = None
        # the count is produced on the stream that is current now; whoever materialises later (maybe
        # under another current stream) must wait for it
        self._ready = None
        if not torch.cuda.is_current_stream_capturing():   # (a captured forward is re-armed by GraphedForward)
            self._ready = torch.cuda.Event()
            self._ready.record()
        self._count_host = None

    def note_window_overflow(self, flag):
        """flag: device int32[1], number of windows a window partition had to drop (max_num_wins per sample /
        list capacity).  The kernels stay memory-safe (dropped windows simply get no attention); the error is
        raised at the next point where the host looks at the tensor anyway (.features / .indices,
        prefetch_row_count + check_window_overflow), never by an extra synchronisation in the forward."""
        flags = self.__dict__.setdefault("_overflow_flags", [])
        if not any(f.data_ptr() == flag.data_ptr() for f in flags):
            flags.append(flag)
            self._overflow_host = None

    def _overflow_total_dev(self):
        flags = self.__dict__.get("_overflow_flags") or []
        if not flags:
            return None
        return flags[0] if len(flags) == 1 else torch.stack([f.reshape(()) for f in flags]).sum().reshape(1).int()

    def prefetch_row_count(self):
        """Start the 8-byte readback of (row count, dropped windows) without waiting for it, so that a
        pipelined caller can queue the next frame before it looks at .features of this one."""
        if self._count_host is None and (self._lazy is not None or self.__dict__.get("_overflow_flags")):
            host = _PINNED.get()
            if self._lazy is not None:
                host[0:1].copy_(self._lazy[2], non_blocking=True)
            dropped = self._overflow_total_dev()
            if dropped is not None:
                host[1:2].copy_(dropped, non_blocking=True)
            else:
                host[1] = 0
            self._count_host = host
            self._ready = torch.cuda.Event()
            self._ready.record()

    def check_window_overflow(self):
        """Raise if any window partition of the forward that produced this tensor dropped windows (one
        4-byte readback, or none after prefetch_row_count)."""
        if self.__dict__.get("_overflow_host") is None:
            if self._count_host is not None:
                if self._ready is not None:
                    self._ready.synchronize()
                self._overflow_host = int(self._count_host[1])
            else:
                dropped = self._overflow_total_dev()
                self._overflow_host = int(dropped.item()) if dropped is not None else 0
        if self._overflow_host:
            raise RuntimeError("window partition: %d windows exceed max_num_wins (or the window list capacity); "
                               "their voxels received no attention" % self._overflow_host)

    def _materialise(self):
        if self._lazy is not None:
            f, i, count = self._lazy
            if self._ready is not None:
                self._ready.synchronize()
            if self._count_host is not None:
                n = int(self._count_host[0])
                self._overflow_host = int(self._count_host[1])
                _PINNED.release(self._count_host, self._ready)
                self._count_host = None
            else:
                n = int(count.item())
            self._features, self._indices, self._lazy = f[:n], i[:n], None
            self.check_window_overflow()

    def _armed_check(self):
        # armed by MixedScaleSparseTransformer.forward when it hands the tensor out: the first look at the rows
        # from outside also reports dropped windows (inside the forward nothing synchronises)
        if self.__dict__.get("_overflow_armed") and not torch.cuda.is_current_stream_capturing():
            self._overflow_armed = False
            self.check_window_overflow()

    @property
    def features(self):
        self._materialise()
        self._armed_check()
        return self._features

    @features.setter
    def features(self, value):
        self._materialise()
        self._features = value

    @property
    def indices(self):
        self._materialise()
        self._armed_check()
        return self._indices

    @indices.setter
    def indices(self, value):
        self._materialise()
        self._indices = value

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
        if self._lazy is not None:  # no host sync needed: the kernel reads the row count on the device
            feats, idx, count = self._lazy
        else:
            feats, idx, count = self.features.float().contiguous(), self.indices, None
        C = feats.shape[1]
        if torch.is_grad_enabled() and feats.requires_grad:
            # training: the scatter has to stay in the autograd graph (index_put; backward = row gather)
            out = feats.new_zeros((self.batch_size, z, y, x, C))
            i = idx.long()
            out = out.index_put((i[:, 0], i[:, 1], i[:, 2], i[:, 3]), feats)
            return out.permute(0, 4, 1, 2, 3) if channels_first else out
        out = torch.empty((self.batch_size, C, z, y, x), dtype=torch.float32, device=feats.device)
        call("mssvt_dense_scatter", feats.shape[0], ptr(count), self.batch_size, C, z, y, x, ptr(feats),
             ptr(idx), ptr(out), stream())
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

    def forward(self, query, keys, batch_first=False, query_mask=None, key_masks=None):
        """Differentiable dense form (mssvt_utils.py:88-157), used by the PADDED training path of the blocks (the
        cross-check of the ragged training kernels, train_ops.py); inference runs the same mathematics inside the fused
        window kernels (mssvt_block_attention[_tc] / mssvt_compress_attention[_tc]) and never comes through here.  query (b, nq, C), keys (b, G * nk, C)
        when batch_first; head group g reads channel slice g of the queries and key chunk g only; key_masks
        (b, G * nk) bool (True = masked, additive -100 like the reference); padded queries give zero rows."""
        if not query.is_cuda:
            raise RuntimeError("mssvt_b200 runs on CUDA tensors only; there is no CPU path")
        if not batch_first:
            query, keys = query.transpose(1, 0), keys.transpose(1, 0)
        b, nq, _ = query.shape
        nk = keys.shape[1] // self.num_head_groups
        outs, c0 = [], 0
        for g, heads in enumerate(self.num_heads):
            c1 = self.group_c_idx[g]
            q = self.to_qs[g](query[:, :, c0:c1]).reshape(b, nq, heads, self.per_head_dim).permute(0, 2, 1, 3)
            kv = self.to_kvs[g](keys[:, g * nk:(g + 1) * nk, c0:c1])
            kv = kv.reshape(b, nk, 2, heads, self.per_head_dim).permute(2, 0, 3, 1, 4)
            k, v = kv[0], kv[1]
            c0 = c1
            attn = (q * self.scale) @ k.transpose(-2, -1)                      # (b, heads, nq, nk)
            if key_masks is not None:
                km = key_masks[:, g * nk:(g + 1) * nk]
                attn = attn + (km.to(attn.dtype) * -100.0).view(b, 1, 1, nk)
                attn = torch.softmax(attn, dim=-1)
            attn = self.attn_drop(attn)
            x = (attn @ v).transpose(1, 2).reshape(b, nq, -1)
            outs.append(self.proj_drop(self.projs[g](x)))
        out = torch.cat(outs, dim=-1)
        if query_mask is not None:
            out = out * (~query_mask).unsqueeze(-1).to(out.dtype)
        return out if batch_first else out.transpose(1, 0)
