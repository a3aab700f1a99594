"""CPU oracle: the reference's operator API (pcdet/ops/mssvt/mssvt_ops.py and the four
pcdet/ops/pointnet2/pointnet2_batch/pointnet2_utils.py ops on the hot path) on CPU tensors.

TEST INFRASTRUCTURE -- not product.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this package.

Each function allocates and pre-fills its outputs exactly as the reference's autograd.Function
does (citations per function) and then runs the sequential C restatement of the kernel in
oracle/mssvt_oracle.c through ctypes.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libmssvt_oracle.so")
        src = os.path.join(_HERE, "mssvt_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s", "all"])
        _LIB = ctypes.CDLL(path)
        _LIB.orc_window_partition.restype = ctypes.c_int
        _LIB.orc_num_threads.restype = ctypes.c_int
    return _LIB


def _p(t):
    assert t.device.type == "cpu" and t.is_contiguous(), "oracle works on contiguous CPU tensors"
    return ctypes.c_void_p(t.data_ptr())


def _i32(t):
    return t.to(torch.int32).contiguous()


def num_threads():
    return int(lib().orc_num_threads())


# ----------------------------------------------------------------------------- mssvt_ops

def build_hash_table(batch_size, hash_size, spatial_shape, voxel_indices, v_bs_cnt):
    """mssvt_ops.py:7-26 (BuildHashTable): (B, H, 2) int32 filled -1, then the insert kernel."""
    x_max, y_max, z_max = (int(v) for v in spatial_shape)
    voxel_indices, v_bs_cnt = _i32(voxel_indices), _i32(v_bs_cnt)
    table = torch.full((batch_size, hash_size, 2), -1, dtype=torch.int32)
    lib().orc_build_hash_table(x_max, y_max, z_max, voxel_indices.shape[0], hash_size,
                               _p(voxel_indices), _p(v_bs_cnt), _p(table))
    return table


def hash_lookup(table, batch_ids, keys):
    """Content view of a table: value per (sample, key), -1 when absent."""
    batch_ids, keys = _i32(batch_ids), _i32(keys)
    table = _i32(table)
    out = torch.empty(keys.shape[0], dtype=torch.int32)
    lib().orc_hash_lookup(table.shape[1], keys.shape[0], _p(batch_ids), _p(keys), _p(table), _p(out))
    return out


def get_non_empty_window_center(win_size, max_num_wins, batch_size, hash_size, spatial_shape,
                                voxel_indices):
    """mssvt_ops.py:29-60 (WindowPartition): returns (win_list (W,4) [b,wz,wy,wx], table)."""
    x_ws, y_ws, z_ws = (int(v) for v in win_size)
    x_wgs, y_wgs, z_wgs = (int(v) for v in spatial_shape)
    voxel_indices = _i32(voxel_indices)
    table = torch.full((batch_size, hash_size, 2), -1, dtype=torch.int32)
    rows = torch.full((batch_size, max_num_wins, 3), -1, dtype=torch.int32)
    vcount = torch.zeros(batch_size, dtype=torch.int32)
    over = lib().orc_window_partition(x_wgs, y_wgs, z_wgs, x_ws, y_ws, z_ws,
                                      voxel_indices.shape[0], max_num_wins, hash_size,
                                      _p(voxel_indices), _p(rows), _p(table), _p(vcount))
    if over:
        raise RuntimeError("window list overflow: %d windows beyond max_num_wins=%d "
                           "(the reference writes out of bounds here)" % (over, max_num_wins))
    parts = []
    for b in range(batch_size):
        live = rows[b][rows[b][:, 0] >= 0]
        parts.append(torch.cat([torch.full((live.shape[0], 1), b, dtype=torch.int32), live], 1))
    return torch.cat(parts, 0).contiguous(), table


def gather_two_window_voxels(spatial_shape, win_size, max_num_odd, max_num_even, max_num_win1,
                             max_num_win2, q_odd, q_even, q_win1, q_win2, win_indices, table):
    """mssvt_ops.py:63-102 (GatherTwoWindowVoxels): 4 index lists (-1 pad) + 4 offset lists (0 pad)."""
    x_max, y_max, z_max = (int(v) for v in spatial_shape)
    x_ws, y_ws, z_ws = (int(v) for v in win_size)
    q_odd, q_even, q_win1, q_win2 = _i32(q_odd), _i32(q_even), _i32(q_win1), _i32(q_win2)
    win_indices, table = _i32(win_indices), _i32(table)
    W = win_indices.shape[0]
    caps = (max_num_odd, max_num_even, max_num_win1, max_num_win2)
    inds = [torch.full((W, c), -1, dtype=torch.int32) for c in caps]
    coords = [torch.zeros((W, c, 3), dtype=torch.int32) for c in caps]
    lib().orc_gather_two_window(
        x_max, y_max, z_max, x_ws, y_ws, z_ws, *caps, W, table.shape[1],
        q_odd.shape[0], q_even.shape[0], q_win1.shape[0], q_win2.shape[0],
        *[_p(t) for t in inds], *[_p(t) for t in coords],
        _p(q_odd), _p(q_even), _p(q_win1), _p(q_win2), _p(win_indices), _p(table))
    return (*inds, *coords)


def gather_one_window_voxels(spatial_shape, win_size, max_num_win1, q_win1, win_indices, table):
    """mssvt_ops.py:105-133 (GatherOneWindowVoxels)."""
    x_max, y_max, z_max = (int(v) for v in spatial_shape)
    x_ws, y_ws, z_ws = (int(v) for v in win_size)
    q_win1, win_indices, table = _i32(q_win1), _i32(win_indices), _i32(table)
    W = win_indices.shape[0]
    ind = torch.full((W, max_num_win1), -1, dtype=torch.int32)
    coord = torch.zeros((W, max_num_win1, 3), dtype=torch.int32)
    lib().orc_gather_one_window(x_max, y_max, z_max, x_ws, y_ws, z_ws, max_num_win1, W,
                                table.shape[1], q_win1.shape[0], _p(ind), _p(coord), _p(q_win1),
                                _p(win_indices), _p(table))
    return ind, coord


def grouping_operation(features, features_batch_cnt, idx, idx_batch_cnt):
    """mssvt_ops.py:136-170 (GroupingOperation.forward): (M, C, ns), zeros where idx < 0."""
    features = features.float().contiguous()
    fbc, ibc, idx = _i32(features_batch_cnt), _i32(idx_batch_cnt), _i32(idx)
    assert features.shape[0] == int(fbc.sum()) and idx.shape[0] == int(ibc.sum())
    M, ns = idx.shape
    C = features.shape[1]
    out = torch.zeros((M, C, ns), dtype=torch.float32)
    lib().orc_group_features(ibc.shape[0], M, C, ns, _p(features), _p(fbc), _p(idx), _p(ibc), _p(out))
    return out


def grouping_operation_grad(grad_out, num_features, features_batch_cnt, idx, idx_batch_cnt):
    """mssvt_ops.py:172-190 (GroupingOperation.backward): scatter-add into (N, C)."""
    grad_out = grad_out.float().contiguous()
    fbc, ibc, idx = _i32(features_batch_cnt), _i32(idx_batch_cnt), _i32(idx)
    M, C, ns = grad_out.shape
    grad = torch.zeros((num_features, C), dtype=torch.float32)
    lib().orc_group_features_grad(ibc.shape[0], M, C, num_features, ns, _p(grad_out), _p(idx),
                                  _p(ibc), _p(fbc), _p(grad))
    return grad


# ----------------------------------------------------------------------------- pointnet2_utils

def farthest_point_sample(xyz, npoint):
    """pointnet2_utils.py:10-36: (B, N, 3) float -> (B, npoint) int32; temp starts at 1e10."""
    xyz = xyz.float().contiguous()
    B, N, _ = xyz.shape
    out = torch.zeros((B, npoint), dtype=torch.int32)
    temp = torch.full((B, N), 1e10, dtype=torch.float32)
    lib().orc_fps(B, N, npoint, _p(xyz), _p(temp), _p(out))
    return out


def gather_operation(features, idx):
    """pointnet2_utils.py:39-73: (B, C, N), (B, np) -> (B, C, np)."""
    features, idx = features.float().contiguous(), _i32(idx)
    B, C, N = features.shape
    out = torch.empty((B, C, idx.shape[1]), dtype=torch.float32)
    lib().orc_gather_points(B, C, N, idx.shape[1], _p(features), _p(idx), _p(out))
    return out


def three_nn(unknown, known):
    """pointnet2_utils.py:76-105: returns (sqrt(dist2) (B,n,3), idx (B,n,3))."""
    unknown, known = unknown.float().contiguous(), known.float().contiguous()
    B, n, _ = unknown.shape
    dist2 = torch.empty((B, n, 3), dtype=torch.float32)
    idx = torch.empty((B, n, 3), dtype=torch.int32)
    lib().orc_three_nn(B, n, known.shape[1], _p(unknown), _p(known), _p(dist2), _p(idx))
    return torch.sqrt(dist2), idx


def group_points(features, idx):
    """pointnet2_utils.py:156-177 (GroupingOperation.forward): (B,C,N), (B,np,ns) -> (B,C,np,ns)."""
    features, idx = features.float().contiguous(), _i32(idx)
    B, C, N = features.shape
    _, npnt, ns = idx.shape
    out = torch.empty((B, C, npnt, ns), dtype=torch.float32)
    lib().orc_group_points(B, C, N, npnt, ns, _p(features), _p(idx), _p(out))
    return out
