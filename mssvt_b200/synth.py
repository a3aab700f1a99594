"""Deterministic synthetic LiDAR-like frames on the S0 grid.

There is no dataset access, so benchmarks and tests run on frames made here: a radial
half-normal range profile, ~55 % ground returns in a thin z slab and ~45 % vertical structures
snapped to a 6 m lattice in x.  Voxels come out unique and ordered by (b, x, y, z) with columns
[b, z, y, x], the way the reference's DynamicVFE emits them
(pcdet/models/backbones_3d/vfe/dynamic_vfe.py:93-118), contiguous per sample as the kernels
require (pcdet/datasets/dataset.py:173-178).
"""
import numpy as np

S0_GRID = (468, 468, 32)                       # x, y, z cells
S0_VOXEL = (0.32, 0.32, 0.1875)                # metres
S0_RANGE = (-74.88, -74.88, -2.0, 74.88, 74.88, 4.0)


def _points(rng, n_pts, half_extent):
    r = np.abs(rng.normal(0.0, 28.0, n_pts)) + 2.0
    th = rng.uniform(0.0, 2 * np.pi, n_pts)
    x, y = r * np.cos(th), r * np.sin(th)
    ground = rng.random(n_pts) < 0.55
    z = np.where(ground, rng.normal(-1.6, 0.08, n_pts), rng.uniform(-1.6, 3.98, n_pts))
    x = np.where(ground, x, np.round(x / 6.0) * 6.0 + rng.normal(0.0, 0.15, n_pts))
    keep = (np.abs(x) < half_extent) & (np.abs(y) < half_extent)
    return x[keep], y[keep], z[keep]


def _voxelise(x, y, z, grid, voxel, rng_range):
    ix = np.floor((x - rng_range[0]) / voxel[0]).astype(np.int64)
    iy = np.floor((y - rng_range[1]) / voxel[1]).astype(np.int64)
    iz = np.floor((z - rng_range[2]) / voxel[2]).astype(np.int64)
    ok = (ix >= 0) & (ix < grid[0]) & (iy >= 0) & (iy < grid[1]) & (iz >= 0) & (iz < grid[2])
    key = (ix[ok] * grid[1] + iy[ok]) * grid[2] + iz[ok]
    return np.unique(key)  # sorted by (x, y, z)


def synth_coords(seed, n_target, grid=S0_GRID, voxel=S0_VOXEL, pc_range=S0_RANGE, crop=1.0):
    """Exactly n_target unique voxel keys of one sample, sorted by (x, y, z).
    crop < 1 restricts the frame to the central crop*extent square so that a small frame keeps
    Waymo-like density (BASELINE config 1: 20 k voxels on ~1/7 of the area => crop ~ 0.38)."""
    rng = np.random.default_rng(seed)
    half = 0.5 * (pc_range[3] - pc_range[0]) * crop
    n_pts = int(n_target * 3)
    for _ in range(12):
        keys = _voxelise(*_points(rng, n_pts, half), grid, voxel, pc_range)
        if n_target <= len(keys) <= 1.03 * n_target:
            break
        n_pts = max(16, int(n_pts * (1.015 * n_target / max(len(keys), 1)) ** 1.35))
    if len(keys) < n_target:  # top up with uniform cells (tiny frames only)
        extra = rng.choice(grid[0] * grid[1] * grid[2], size=4 * n_target, replace=False)
        keys = np.unique(np.concatenate([keys, extra[: n_target - len(keys) + 64]]))
    if len(keys) > n_target:
        keys = np.sort(rng.choice(keys, size=n_target, replace=False))
    assert len(keys) == n_target
    return keys


def synth_frame(seed, n_target, batch_size=1, channels=64, grid=S0_GRID, voxel=S0_VOXEL,
                pc_range=S0_RANGE, crop=1.0):
    """-> (voxel_features (N, C) float32 ~N(0,1), voxel_coords (N, 4) int32 [b, z, y, x]);
    N = batch_size * n_target; sample s uses seed + s."""
    coords = []
    for b in range(batch_size):
        keys = synth_coords(seed + b, n_target, grid, voxel, pc_range, crop)
        z = keys % grid[2]
        y = (keys // grid[2]) % grid[1]
        x = keys // (grid[2] * grid[1])
        coords.append(np.stack([np.full_like(x, b), z, y, x], 1))
    coords = np.concatenate(coords, 0).astype(np.int32)
    feats = np.random.default_rng(seed + 7919).standard_normal((coords.shape[0], channels),
                                                                dtype=np.float32)
    return feats, coords


def synth_points(seed, n_voxels, extra=0.2, batch_size=1, point_features=5, grid=S0_GRID, voxel=S0_VOXEL,
                 pc_range=S0_RANGE, crop=1.0):
    """A raw point cloud whose voxelisation is exactly the frame synth_frame(seed, n_voxels) is built on: one
    point inside every occupied voxel plus `extra` * n_voxels more in randomly chosen occupied voxels, shuffled.
    -> points (P, 1 + point_features) float32 [batch_idx, x, y, z, intensity, elongation, ...] as DynamicVFE
    takes them (pcdet/models/backbones_3d/vfe/dynamic_vfe.py:71-92), P = batch_size * n_voxels * (1 + extra)."""
    out = []
    for b in range(batch_size):
        rng = np.random.default_rng(seed + b + 104729)
        keys = synth_coords(seed + b, n_voxels, grid, voxel, pc_range, crop)
        keys = np.concatenate([keys, rng.choice(keys, size=int(extra * n_voxels))])
        rng.shuffle(keys)
        cell = np.stack([keys // (grid[2] * grid[1]), (keys // grid[2]) % grid[1], keys % grid[2]], 1)
        # strictly inside the voxel, away from its faces (the voxel of a point must not depend on rounding)
        xyz = (cell + rng.uniform(0.05, 0.95, cell.shape)) * np.asarray(voxel) + np.asarray(pc_range[:3])
        rest = rng.random((len(keys), point_features - 3))
        out.append(np.concatenate([np.full((len(keys), 1), b), xyz, rest], 1))
    return np.concatenate(out, 0).astype(np.float32)
