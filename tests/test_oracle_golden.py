"""CPU: the oracle against the committed golden vectors (made by oracle/pin_against_reference.py
from the reference's unmodified Python) and against closed-form properties of its C kernels."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from oracle import backbone as orc
from oracle import ops as oops
from mssvt_b200.synth import S0_VOXEL


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_backbone_matches_golden(name):
    blob, cfg, state = load_golden(name)
    taps = []
    with torch.no_grad():
        out = orc.backbone_forward(state, cfg, blob["grid"].tolist(), list(S0_VOXEL),
                                   blob["pc_range"].tolist(), torch.from_numpy(blob["voxel_features"]),
                                   torch.from_numpy(blob["voxel_coords"]), int(blob["batch_size"]), taps=taps)
    assert np.array_equal(out.indices.numpy(), blob["out_indices"])          # bit-exact
    assert np.abs(out.features.numpy() - blob["out_features"]).max() <= 2e-5  # fp32, same library
    for k in ("win_ind", "q_ind", "win1_ind", "win2_ind", "fps_win1", "fps_win2", "k_ind_win1",
              "k_ind_win2", "k_mask_win1", "k_mask_win2", "nn_idx"):
        if "tap0/" + k in blob:
            assert np.array_equal(taps[0][k].numpy(), blob["tap0/" + k]), k


def _fps_rule(pts, m):
    """closed form of the reference FPS: pick arg-max of the running min distance,
    ties broken by (bit_reverse(k mod B), k) with B = 2^floor(log2 n) <= 1024."""
    n = len(pts)
    logb = min(int(np.log(float(n)) / np.log(2.0)), 10)
    B = 1 << logb
    rev = lambda v: int(format(v, "0%db" % logb)[::-1], 2) if logb else 0
    tmin = np.full(n, 1e10, dtype=np.float32)
    picks, old = [0], 0
    for _ in range(1, m):
        d = ((pts - pts[old]) ** 2).sum(1).astype(np.float32)
        tmin = np.minimum(tmin, d)
        best = tmin.max()
        cands = [k for k in range(n) if tmin[k] == best]
        old = min(cands, key=lambda k: (rev(k % B), k))
        picks.append(old)
    return np.array(picks, dtype=np.int32)


@pytest.mark.parametrize("n,m", [(27, 32), (125, 32), (12, 8), (32, 32), (1, 4), (2, 3), (100, 40)])
def test_fps_literal_simulation_equals_tie_rule(n, m):
    rng = np.random.default_rng(n * 1000 + m)
    for _ in range(20):
        pts = rng.integers(-3, 4, size=(n, 3)).astype(np.float32)
        pts[rng.random(n) < 0.4] = 0.0  # padding slots at the origin
        got = oops.farthest_point_sample(torch.from_numpy(pts)[None], m)[0].numpy()
        assert np.array_equal(got, _fps_rule(pts, m))


def test_hash_table_roundtrip_and_collisions():
    rng = np.random.default_rng(3)
    grid = (20, 18, 6)
    B, H = 3, 257  # tiny prime table: long probe chains, ~78 % load
    coords = []
    for b in range(B):
        keys = rng.choice(grid[0] * grid[1] * grid[2], size=200, replace=False)
        keys.sort()
        coords.append(np.stack([np.full(200, b), keys % grid[2], (keys // grid[2]) % grid[1],
                                keys // (grid[2] * grid[1])], 1))
    coords = torch.from_numpy(np.concatenate(coords).astype(np.int32))
    cnt = torch.full((B,), 200, dtype=torch.int32)
    tab = oops.build_hash_table(B, H, grid, coords, cnt)
    key = coords[:, 3] * grid[1] * grid[2] + coords[:, 2] * grid[2] + coords[:, 1]
    val = oops.hash_lookup(tab, coords[:, 0], key)
    assert torch.equal(val, torch.arange(200, dtype=torch.int32).repeat(B))
    absent = oops.hash_lookup(tab, torch.zeros(50, dtype=torch.int32), torch.arange(50, dtype=torch.int32) + 10 ** 6)
    assert (absent == -1).all()
    # every sample's table holds exactly its 200 keys
    assert ((tab[:, :, 0] >= 0).sum(1) == 200).all()


def test_window_partition_first_occurrence_order_and_empty():
    coords = torch.tensor([[0, 0, 5, 5], [0, 1, 0, 0], [0, 2, 5, 4], [0, 31, 1, 1], [1, 0, 0, 0]], dtype=torch.int32)
    win, tab = oops.get_non_empty_window_center([3, 3, 3], 90000, 2, 101, [16, 16, 10], coords)
    # z = 31 falls in the remainder strip (32 // 3 = 10 windows) and opens no window
    assert win.tolist() == [[0, 0, 1, 1], [0, 0, 0, 0], [1, 0, 0, 0]]
    empty = torch.zeros((0, 4), dtype=torch.int32)
    win, _ = oops.get_non_empty_window_center([3, 3, 3], 10, 1, 11, [4, 4, 4], empty)
    assert win.shape == (0, 4)


def test_gather_lists_caps_and_order():
    # one full 3x3x3 window around (4,4,4) in a 9^3 grid; caps smaller than the population
    grid = (9, 9, 9)
    xs = np.array([(x, y, z) for x in range(9) for y in range(9) for z in range(9)])
    coords = torch.from_numpy(np.stack([np.zeros(len(xs)), xs[:, 2], xs[:, 1], xs[:, 0]], 1).astype(np.int32))
    tab = oops.build_hash_table(1, 2003, grid, coords, torch.tensor([len(xs)], dtype=torch.int32))
    t = orc.vox_query_table([3, 3, 3], [5, 5, 5])
    win = torch.tensor([[0, 1, 1, 1]], dtype=torch.int32)
    outs = oops.gather_two_window_voxels(grid, [3, 3, 3], 12, 3, 20, 60, t["odd"], t["even"], t["win1"],
                                         t["win2"], win, tab)
    i_odd, i_even, i_w1, i_w2, c_odd, c_even, c_w1, c_w2 = outs
    assert (i_odd >= 0).all() and (i_even >= 0).all() and (i_w1 >= 0).all() and (i_w2 >= 0).all()
    # win1 list = odd ++ even ++ rest, capped at 20: 12 odd, 3 even, first 5 of the rest
    assert torch.equal(i_w1[0, :12], i_odd[0]) and torch.equal(i_w1[0, 12:15], i_even[0])
    assert torch.equal(c_w1[0, 15:], t["win1"][:5])
    assert torch.equal(i_w2[0, :20], i_w1[0])
    # value stored = row index of voxel centre + offset
    key = lambda o: (4 + o[0]) * 81 + (4 + o[1]) * 9 + (4 + o[2])
    assert i_odd[0].tolist() == [key(o) for o in t["odd"].tolist()]


def test_three_nn_ties_pick_lowest_index_and_short_known():
    unknown = torch.zeros((1, 2, 3))
    known = torch.tensor([[[1., 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0]]])
    dist, idx = oops.three_nn(unknown, known)
    assert idx[0, 0].tolist() == [0, 1, 2] and torch.allclose(dist[0, 0], torch.ones(3))
    dist, idx = oops.three_nn(unknown, known[:, :2])
    assert idx[0, 0].tolist() == [0, 1, 0] and torch.isinf(dist[0, 0, 2])


def test_grouping_grad_is_transpose_of_forward():
    rng = np.random.default_rng(5)
    feats = torch.from_numpy(rng.standard_normal((30, 8)).astype(np.float32))
    fbc, ibc = torch.tensor([10, 20], dtype=torch.int32), torch.tensor([3, 4], dtype=torch.int32)
    idx = torch.from_numpy(rng.integers(-1, 10, size=(7, 5)).astype(np.int32))
    out = oops.grouping_operation(feats, fbc, idx, ibc)
    g = torch.from_numpy(rng.standard_normal(out.shape).astype(np.float32))
    gf = oops.grouping_operation_grad(g, 30, fbc, idx, ibc)
    assert abs(float((out * g).sum()) - float((feats * gf).sum())) < 1e-3
