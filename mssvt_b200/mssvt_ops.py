"""Operator API of pcdet/ops/mssvt/mssvt_ops.py on top of libmssvt_b200.so.

Same names, argument order, tensor contracts and autograd behaviour as the reference's five
`autograd.Function`s, so code written against `pcdet.ops.mssvt.mssvt_ops` runs unchanged:
    build_hash_table, get_non_empty_window_center, gather_two_window_voxels,
    gather_one_window_voxels, grouping_operation
Differences in HOW (not in results): outputs are allocated on the device (the reference fills
them on the CPU and copies them over PCIe, mssvt_ops.py:16-17, 36-41, 77-85), launches go to the
current stream, windows are numbered deterministically, errors raise instead of exit(-1).
There is no CPU path: CPU tensors raise.
"""
import torch
from torch.autograd import Function

from . import _lib
from ._lib import call, ptr, stream


def _i32(t):
    return t if t.dtype == torch.int32 and t.is_contiguous() else t.to(torch.int32).contiguous()


class BuildHashTable(Function):
    """mssvt_ops.py:7-26."""

    @staticmethod
    def forward(ctx, batch_size, hash_size, spatial_shape, voxel_indices, v_bs_cnt):
        x_max, y_max, z_max = (int(v) for v in spatial_shape)
        assert voxel_indices.is_contiguous()
        voxel_indices, v_bs_cnt = _i32(voxel_indices), _i32(v_bs_cnt)
        dense_map = torch.empty((batch_size, hash_size, 2), dtype=torch.int32, device=voxel_indices.device)
        call("mssvt_build_hash_table", x_max, y_max, z_max, voxel_indices.shape[0], hash_size,
             batch_size, ptr(voxel_indices), ptr(v_bs_cnt), ptr(dense_map), stream())
        return dense_map

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None, None


build_hash_table = BuildHashTable.apply


def hash_lookup(dense_map, batch_ids, keys):
    """values stored for (sample, key) pairs, -1 when absent (content view of a table)."""
    batch_ids, keys = _i32(batch_ids), _i32(keys)
    out = torch.empty(keys.shape[0], dtype=torch.int32, device=keys.device)
    call("mssvt_hash_lookup", dense_map.shape[1], keys.shape[0], ptr(batch_ids), ptr(keys),
         ptr(dense_map), ptr(out), stream())
    return out


def window_partition_device(win_size, max_num_wins, batch_size, hash_size, spatial_shape,
                            voxel_indices, capacity=None):
    """Sync-free form used by the backbone: returns (win_list (capacity, 4), table, win_count
    (B + 2) on the device: per-sample counts, total, dropped)."""
    x_ws, y_ws, z_ws = (int(v) for v in win_size)
    x_wgs, y_wgs, z_wgs = (int(v) for v in spatial_shape)
    voxel_indices = _i32(voxel_indices)
    n = voxel_indices.shape[0]
    dev = voxel_indices.device
    if capacity is None:
        capacity = min(n, batch_size * max_num_wins)
    table = torch.empty((batch_size, hash_size, 2), dtype=torch.int32, device=dev)
    win_list = torch.empty((max(capacity, 1), 4), dtype=torch.int32, device=dev)
    win_count = torch.empty(batch_size + 2, dtype=torch.int32, device=dev)
    ws_bytes = call("mssvt_window_partition_workspace_bytes", n)
    workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    call("mssvt_window_partition", x_wgs, y_wgs, z_wgs, x_ws, y_ws, z_ws, n, max_num_wins,
         hash_size, batch_size, capacity, ptr(voxel_indices), ptr(win_list), ptr(table),
         ptr(win_count), ptr(workspace), ws_bytes, stream())
    return win_list, table, win_count


def window_list_device(win_size, max_num_wins, batch_size, spatial_shape, voxel_indices, capacity=None):
    """The window list of window_partition_device WITHOUT the window hash table (the fused backbone finds voxels
    through its grid index): (win_list (capacity, 4), win_count (B + 2)) on the device, same rows, same order."""
    x_ws, y_ws, z_ws = (int(v) for v in win_size)
    x_wgs, y_wgs, z_wgs = (int(v) for v in spatial_shape)
    voxel_indices = _i32(voxel_indices)
    n, dev = voxel_indices.shape[0], voxel_indices.device
    if capacity is None:
        capacity = min(n, batch_size * max_num_wins)
    win_list = torch.empty((max(capacity, 1), 4), dtype=torch.int32, device=dev)
    win_count = torch.empty(batch_size + 2, dtype=torch.int32, device=dev)
    ws_bytes = call("mssvt_window_list_workspace_bytes", x_wgs, y_wgs, z_wgs, batch_size, n)
    workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    call("mssvt_window_list", x_wgs, y_wgs, z_wgs, x_ws, y_ws, z_ws, n, max_num_wins, batch_size, capacity,
         ptr(voxel_indices), ptr(win_list), ptr(win_count), ptr(workspace), ws_bytes, stream())
    return win_list, win_count


class WindowPartition(Function):
    """mssvt_ops.py:29-60.  Rows are ordered by first occurrence in voxel order (one legal
    outcome of the reference's atomicAdd numbering, and the same on every run)."""

    @staticmethod
    def forward(ctx, win_size, max_num_wins, batch_size, hash_size, spatial_shape, voxel_indices):
        assert voxel_indices.is_contiguous()
        win_list, table, win_count = window_partition_device(
            win_size, max_num_wins, batch_size, hash_size, spatial_shape, voxel_indices)
        counts = win_count.tolist()  # the reference API returns exact shapes: one sync
        if counts[batch_size + 1]:
            raise RuntimeError("get_non_empty_window_center: %d windows exceed max_num_wins=%d per "
                               "sample" % (counts[batch_size + 1], max_num_wins))
        return win_list[:counts[batch_size]].contiguous(), table

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None, None, None, None


get_non_empty_window_center = WindowPartition.apply


class GatherTwoWindowVoxels(Function):
    """mssvt_ops.py:63-102."""

    @staticmethod
    def forward(ctx, spatial_shape, win_size, max_num_odd, max_num_even, max_num_win1, max_num_win2,
                vox_query_odd, vox_query_even, vox_query_win1, vox_query_win2, win_indices, dense_map):
        x_max, y_max, z_max = (int(v) for v in spatial_shape)
        x_ws, y_ws, z_ws = (int(v) for v in win_size)
        hash_size = dense_map.shape[1]
        assert win_indices.is_contiguous()
        tabs = [_i32(t) for t in (vox_query_odd, vox_query_even, vox_query_win1, vox_query_win2)]
        win_indices = _i32(win_indices)
        W, dev = win_indices.shape[0], win_indices.device
        caps = (max_num_odd, max_num_even, max_num_win1, max_num_win2)
        inds = [torch.empty((W, c), dtype=torch.int32, device=dev) for c in caps]
        coords = [torch.empty((W, c, 3), dtype=torch.int32, device=dev) for c in caps]
        call("mssvt_gather_two_window", x_max, y_max, z_max, x_ws, y_ws, z_ws, *caps, W, hash_size,
             *[t.shape[0] for t in tabs], *[ptr(t) for t in inds], *[ptr(t) for t in coords],
             *[ptr(t) for t in tabs], ptr(win_indices), ptr(dense_map), stream())
        return (*inds, *coords)

    @staticmethod
    def backward(ctx, *a):
        return (None,) * 12


gather_two_window_voxels = GatherTwoWindowVoxels.apply


class GatherOneWindowVoxels(Function):
    """mssvt_ops.py:105-133."""

    @staticmethod
    def forward(ctx, spatial_shape, win_size, max_num_win1, vox_query_win1, win_indices, dense_map):
        x_max, y_max, z_max = (int(v) for v in spatial_shape)
        x_ws, y_ws, z_ws = (int(v) for v in win_size)
        hash_size = dense_map.shape[1]
        assert win_indices.is_contiguous()
        vox_query_win1, win_indices = _i32(vox_query_win1), _i32(win_indices)
        W, dev = win_indices.shape[0], win_indices.device
        ind = torch.empty((W, max_num_win1), dtype=torch.int32, device=dev)
        coord = torch.empty((W, max_num_win1, 3), dtype=torch.int32, device=dev)
        call("mssvt_gather_one_window", x_max, y_max, z_max, x_ws, y_ws, z_ws, max_num_win1, W,
             hash_size, vox_query_win1.shape[0], ptr(ind), ptr(coord), ptr(vox_query_win1),
             ptr(win_indices), ptr(dense_map), stream())
        return ind, coord

    @staticmethod
    def backward(ctx, *a):
        return (None,) * 6


gather_one_window_voxels = GatherOneWindowVoxels.apply


class GroupingOperation(Function):
    """mssvt_ops.py:136-192: out[m, :, s] = features[start(sample of m) + idx[m, s], :], zeros for
    idx < 0; backward scatter-adds into the features."""

    @staticmethod
    def forward(ctx, features, features_batch_cnt, idx, idx_batch_cnt):
        assert features.is_contiguous() and features_batch_cnt.is_contiguous()
        assert idx.is_contiguous() and idx_batch_cnt.is_contiguous()
        # (the reference also asserts the two count sums on the host: two syncs per call;
        #  the kernel here never reads past `M` / the sample starts, so the check is dropped)
        M, nsample = idx.size()
        N, C = features.size()
        B = idx_batch_cnt.shape[0]
        output = torch.empty((M, C, nsample), dtype=torch.float32, device=features.device)
        call("mssvt_group_features", B, M, C, nsample, ptr(features), ptr(_i32(features_batch_cnt)),
             ptr(_i32(idx)), ptr(_i32(idx_batch_cnt)), ptr(output), stream())
        ctx.for_backwards = (B, N, idx, features_batch_cnt, idx_batch_cnt)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        B, N, idx, features_batch_cnt, idx_batch_cnt = ctx.for_backwards
        M, C, nsample = grad_out.size()
        grad_features = torch.empty((N, C), dtype=torch.float32, device=grad_out.device)
        grad_out = grad_out.contiguous()
        call("mssvt_group_features_grad", B, M, C, N, nsample, ptr(grad_out), ptr(_i32(idx)),
             ptr(_i32(idx_batch_cnt)), ptr(_i32(features_batch_cnt)), ptr(grad_features), stream())
        return grad_features, None, None, None


grouping_operation = GroupingOperation.apply

__all__ = ["build_hash_table", "get_non_empty_window_center", "gather_two_window_voxels",
           "gather_one_window_voxels", "grouping_operation", "hash_lookup",
           "window_partition_device"]
_ = _lib
