"""The reference's OWN CUDA kernels (oracle/_ref/libmssvt_ref.so, built unmodified from
/root/reference by `make -C oracle ref`) behind the same Python API as oracle/ops.py, on CUDA
tensors.  TEST INFRASTRUCTURE: used on the GPU box to pin the CPU oracle (and the product) against
the real reference kernels.  Outputs are allocated and pre-filled exactly as the reference's
autograd.Functions do (mssvt_ops.py / pointnet2_utils.py)."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libmssvt_ref.so")
_LIB = None


def available():
    return os.path.exists(PATH)


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(PATH)
    return _LIB


def _p(t):
    assert t.is_cuda and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def _ok(rc):
    if rc != 0:
        raise RuntimeError("reference kernel failed: cudaError %d" % rc)


def _i32(t):
    return t.to(torch.int32).contiguous()


def build_hash_table(batch_size, hash_size, spatial_shape, voxel_indices, v_bs_cnt):
    x, y, z = (int(v) for v in spatial_shape)
    voxel_indices, v_bs_cnt = _i32(voxel_indices), _i32(v_bs_cnt)
    table = torch.full((batch_size, hash_size, 2), -1, dtype=torch.int32, device=voxel_indices.device)
    torch.cuda.synchronize()
    _ok(lib().ref_build_hash_table(x, y, z, voxel_indices.shape[0], hash_size, _p(voxel_indices),
                                   _p(v_bs_cnt), _p(table)))
    return table


def get_non_empty_window_center(win_size, max_num_wins, batch_size, hash_size, spatial_shape, voxel_indices):
    xw, yw, zw = (int(v) for v in win_size)
    xg, yg, zg = (int(v) for v in spatial_shape)
    voxel_indices = _i32(voxel_indices)
    dev = voxel_indices.device
    table = torch.full((batch_size, hash_size, 2), -1, dtype=torch.int32, device=dev)
    rows = torch.full((batch_size, max_num_wins, 3), -1, dtype=torch.int32, device=dev)
    vcount = torch.zeros(batch_size, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    _ok(lib().ref_window_partition(xg, yg, zg, xw, yw, zw, voxel_indices.shape[0], max_num_wins, hash_size,
                                   _p(voxel_indices), _p(rows), _p(table), _p(vcount)))
    parts = []
    for b in range(batch_size):
        live = rows[b][rows[b][:, 0] >= 0]
        parts.append(torch.cat([torch.full((live.shape[0], 1), b, dtype=torch.int32, device=dev), live], 1))
    return torch.cat(parts, 0).contiguous(), table


def gather_two_window_voxels(spatial_shape, win_size, mo, me, m1, m2, q_odd, q_even, q_win1, q_win2,
                             win_indices, table):
    x, y, z = (int(v) for v in spatial_shape)
    xw, yw, zw = (int(v) for v in win_size)
    tabs = [_i32(t) for t in (q_odd, q_even, q_win1, q_win2)]
    win_indices, table = _i32(win_indices), _i32(table)
    W, dev = win_indices.shape[0], win_indices.device
    caps = (mo, me, m1, m2)
    inds = [torch.full((W, c), -1, dtype=torch.int32, device=dev) for c in caps]
    coords = [torch.zeros((W, c, 3), dtype=torch.int32, device=dev) for c in caps]
    torch.cuda.synchronize()
    _ok(lib().ref_gather_two_window(x, y, z, xw, yw, zw, *caps, W, table.shape[1], *[t.shape[0] for t in tabs],
                                    *[_p(t) for t in inds], *[_p(t) for t in coords], *[_p(t) for t in tabs],
                                    _p(win_indices), _p(table)))
    return (*inds, *coords)


def gather_one_window_voxels(spatial_shape, win_size, m1, q_win1, win_indices, table):
    x, y, z = (int(v) for v in spatial_shape)
    xw, yw, zw = (int(v) for v in win_size)
    q_win1, win_indices, table = _i32(q_win1), _i32(win_indices), _i32(table)
    W, dev = win_indices.shape[0], win_indices.device
    ind = torch.full((W, m1), -1, dtype=torch.int32, device=dev)
    coord = torch.zeros((W, m1, 3), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    _ok(lib().ref_gather_one_window(x, y, z, xw, yw, zw, m1, W, table.shape[1], q_win1.shape[0], _p(ind),
                                    _p(coord), _p(q_win1), _p(win_indices), _p(table)))
    return ind, coord


def grouping_operation(features, fbc, idx, ibc):
    features, fbc, idx, ibc = features.float().contiguous(), _i32(fbc), _i32(idx), _i32(ibc)
    M, ns = idx.shape
    C = features.shape[1]
    out = torch.zeros((M, C, ns), dtype=torch.float32, device=features.device)
    torch.cuda.synchronize()
    _ok(lib().ref_group_features(ibc.shape[0], M, C, ns, _p(features), _p(fbc), _p(idx), _p(ibc), _p(out)))
    return out


def farthest_point_sample(xyz, npoint):
    xyz = xyz.float().contiguous()
    B, N, _ = xyz.shape
    out = torch.zeros((B, npoint), dtype=torch.int32, device=xyz.device)
    temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
    torch.cuda.synchronize()
    _ok(lib().ref_fps(B, N, npoint, _p(xyz), _p(temp), _p(out)))
    return out


def gather_operation(features, idx):
    features, idx = features.float().contiguous(), _i32(idx)
    B, C, N = features.shape
    out = torch.empty((B, C, idx.shape[1]), dtype=torch.float32, device=features.device)
    torch.cuda.synchronize()
    _ok(lib().ref_gather_points(B, C, N, idx.shape[1], _p(features), _p(idx), _p(out)))
    return out


def three_nn(unknown, known):
    unknown, known = unknown.float().contiguous(), known.float().contiguous()
    B, n, _ = unknown.shape
    dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknown.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknown.device)
    torch.cuda.synchronize()
    _ok(lib().ref_three_nn(B, n, known.shape[1], _p(unknown), _p(known), _p(dist2), _p(idx)))
    return torch.sqrt(dist2), idx


def group_points(features, idx):
    features, idx = features.float().contiguous(), _i32(idx)
    B, C, N = features.shape
    _, npnt, ns = idx.shape
    out = torch.empty((B, C, npnt, ns), dtype=torch.float32, device=features.device)
    torch.cuda.synchronize()
    _ok(lib().ref_group_points(B, C, N, npnt, ns, _p(features), _p(idx), _p(out)))
    return out
