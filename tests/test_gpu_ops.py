"""GPU: every operator of the drop-in API, called through the C-ABI, against the CPU oracle on
the same seeded inputs -- bit-exact for all integer outputs and pure copies -- and, when
oracle/_ref is present, against the reference's own CUDA kernels as well (which pins the oracle's
C kernels to the real thing)."""
import numpy as np
import pytest
import torch

from oracle import backbone as orc
from oracle import ops as oops
from oracle import ref_kernels as ref
from mssvt_b200 import mssvt_ops, pointnet2_utils
from mssvt_b200.synth import S0_GRID, synth_frame

pytestmark = pytest.mark.gpu
HAVE_REF = ref.available()


def cuda(t):
    return t.cuda()


def frame(seed, n, batch, crop):
    _, coords = synth_frame(seed, n, batch_size=batch, crop=crop)
    coords = torch.from_numpy(coords)
    cnt = torch.bincount(coords[:, 0].long(), minlength=batch).int()
    return coords, cnt


def keys_of(coords, grid):
    return (coords[:, 3] * grid[1] * grid[2] + coords[:, 2] * grid[2] + coords[:, 1]).int()


@pytest.mark.parametrize("n,batch,hash_size", [(20000, 2, 50021), (3000, 3, 3001), (5, 1, 7)])
def test_hash_build_lookups_bit_exact(n, batch, hash_size):
    coords, cnt = frame(1, n, batch, 0.4)
    grid = S0_GRID
    want_tab = oops.build_hash_table(batch, hash_size, grid, coords, cnt)
    got_tab = mssvt_ops.build_hash_table(batch, hash_size, grid, cuda(coords), cuda(cnt))
    assert got_tab.shape == (batch, hash_size, 2) and got_tab.dtype == torch.int32
    rng = np.random.default_rng(0)
    absent = torch.from_numpy(rng.integers(0, grid[0] * grid[1] * grid[2], 4096).astype(np.int32))
    q_keys = torch.cat([keys_of(coords, grid), absent])
    q_b = torch.cat([coords[:, 0], torch.from_numpy(rng.integers(0, batch, 4096).astype(np.int32))]).int()
    want = oops.hash_lookup(want_tab, q_b, q_keys)
    got = mssvt_ops.hash_lookup(got_tab, cuda(q_b), cuda(q_keys)).cpu()
    assert torch.equal(got, want)
    # same set of occupied keys per sample (slot placement may differ: insertion order is free)
    assert torch.equal(torch.sort(got_tab.cpu()[:, :, 0], 1)[0], torch.sort(want_tab[:, :, 0], 1)[0])
    # tables are interchangeable: the oracle's table read by our lookup kernel
    assert torch.equal(mssvt_ops.hash_lookup(cuda(want_tab), cuda(q_b), cuda(q_keys)).cpu(), want)
    if HAVE_REF:
        ref_tab = ref.build_hash_table(batch, hash_size, grid, cuda(coords), cuda(cnt))
        assert torch.equal(mssvt_ops.hash_lookup(ref_tab, cuda(q_b), cuda(q_keys)).cpu(), want)


def test_hash_build_out_of_range_and_empty():
    coords = torch.tensor([[0, 1, 2, 3], [0, 40, 2, 3], [0, 1, -1, 3], [0, 2, 2, 2]], dtype=torch.int32)
    cnt = torch.tensor([4], dtype=torch.int32)
    want = oops.build_hash_table(1, 13, [8, 8, 8], coords, cnt)
    got = mssvt_ops.build_hash_table(1, 13, [8, 8, 8], cuda(coords), cuda(cnt)).cpu()
    assert torch.equal(torch.sort(got[0, :, 0])[0], torch.sort(want[0, :, 0])[0])
    q = torch.tensor([3 * 64 + 2 * 8 + 1, 2 * 64 + 2 * 8 + 2], dtype=torch.int32)
    assert mssvt_ops.hash_lookup(cuda(got), cuda(torch.zeros(2, dtype=torch.int32)), cuda(q)).tolist() == [0, 3]
    empty = mssvt_ops.build_hash_table(2, 5, [8, 8, 8], torch.zeros((0, 4), dtype=torch.int32).cuda(),
                                       torch.zeros(2, dtype=torch.int32).cuda())
    assert (empty == -1).all()


@pytest.mark.parametrize("win", [[3, 3, 3], [1, 1, 32], [2, 2, 4]])
def test_window_partition_matches_oracle_order(win):
    coords, cnt = frame(2, 20000, 2, 0.4)
    grid = [S0_GRID[i] // win[i] for i in range(3)]
    want_list, want_tab = oops.get_non_empty_window_center(win, 90000, 2, 50021, grid, coords)
    got_list, got_tab = mssvt_ops.get_non_empty_window_center(win, 90000, 2, 50021, grid, cuda(coords))
    assert torch.equal(got_list.cpu(), want_list)  # same rows in the same (first-occurrence) order
    wkey = (want_list[:, 3] * grid[1] * grid[2] + want_list[:, 2] * grid[2] + want_list[:, 1]).int()
    want_val = oops.hash_lookup(want_tab, want_list[:, 0], wkey)
    got_val = mssvt_ops.hash_lookup(got_tab, cuda(want_list[:, 0].contiguous()), cuda(wkey)).cpu()
    assert torch.equal(got_val, want_val)
    if HAVE_REF:  # the reference numbers windows in atomic order: same set, any order
        ref_list, _ = ref.get_non_empty_window_center(win, 90000, 2, 50021, grid, cuda(coords))
        canon = lambda t: sorted(map(tuple, t.cpu().tolist()))
        assert canon(ref_list) == canon(want_list)


def test_window_partition_overflow_is_an_error_not_a_write_past_the_end():
    coords, _ = frame(3, 5000, 1, 0.2)
    with pytest.raises(RuntimeError, match="max_num_wins"):
        mssvt_ops.get_non_empty_window_center([3, 3, 3], 100, 1, 20011, [156, 156, 10], cuda(coords))


@pytest.mark.parametrize("w1,w2,caps", [([3, 3, 3], [5, 5, 5], (27, 125)), ([3, 3, 3], [5, 5, 5], (10, 40)),
                                        ([3, 3, 5], [7, 7, 9], (45, 441))])
def test_gather_two_window_bit_exact(w1, w2, caps):
    coords, cnt = frame(4, 20000, 2, 0.4)
    grid = [S0_GRID[i] // w1[i] for i in range(3)]
    tab = oops.build_hash_table(2, 50021, S0_GRID, coords, cnt)
    win, _ = oops.get_non_empty_window_center(w1, 90000, 2, 50021, grid, coords)
    t = orc.vox_query_table(w1, w2)
    args = (S0_GRID, w1, t["odd"].shape[0], t["even"].shape[0], caps[0], caps[1])
    want = oops.gather_two_window_voxels(*args, t["odd"], t["even"], t["win1"], t["win2"], win, tab)
    got = mssvt_ops.gather_two_window_voxels(*args, *[cuda(t[k]) for k in ("odd", "even", "win1", "win2")],
                                             cuda(win), cuda(tab))
    for g, w in zip(got, want):
        assert torch.equal(g.cpu(), w)
    if HAVE_REF:
        r = ref.gather_two_window_voxels(*args, *[cuda(t[k]) for k in ("odd", "even", "win1", "win2")],
                                         cuda(win), cuda(tab))
        for g, w in zip(r, want):
            assert torch.equal(g.cpu(), w)


def test_gather_one_window_bit_exact_and_empty():
    coords, cnt = frame(5, 20000, 2, 0.4)
    w1 = [1, 1, 32]
    grid = [S0_GRID[i] // w1[i] for i in range(3)]
    tab = oops.build_hash_table(2, 50021, S0_GRID, coords, cnt)
    win, _ = oops.get_non_empty_window_center(w1, 90000, 2, 50021, grid, coords)
    t = orc.vox_query_table(w1)
    want = oops.gather_one_window_voxels(S0_GRID, w1, 32, t["win1"], win, tab)
    got = mssvt_ops.gather_one_window_voxels(S0_GRID, w1, 32, cuda(t["win1"]), cuda(win), cuda(tab))
    assert torch.equal(got[0].cpu(), want[0]) and torch.equal(got[1].cpu(), want[1])
    if HAVE_REF:
        r = ref.gather_one_window_voxels(S0_GRID, w1, 32, cuda(t["win1"]), cuda(win), cuda(tab))
        assert torch.equal(r[0].cpu(), want[0]) and torch.equal(r[1].cpu(), want[1])
    none = mssvt_ops.gather_one_window_voxels(S0_GRID, w1, 32, cuda(t["win1"]), cuda(win[:0].contiguous()), cuda(tab))
    assert none[0].shape == (0, 32) and none[1].shape == (0, 32, 3)


@pytest.mark.parametrize("n,m,kind", [(27, 32, "grid"), (125, 32, "grid"), (12, 8, "grid"), (32, 32, "grid"),
                                      (1, 4, "grid"), (64, 16, "float"), (200, 64, "float"),
                                      (300, 40, "float"), (5000, 128, "float")])
def test_fps_bit_exact_including_tie_order(n, m, kind):
    rng = np.random.default_rng(n + m)
    rows = 64 if n <= 300 else 3
    if kind == "grid":  # integer offsets with padding at the origin: ties everywhere
        pts = rng.integers(-3, 4, size=(rows, n, 3)).astype(np.float32)
        pts[rng.random((rows, n)) < 0.5] = 0.0
    else:
        pts = rng.standard_normal((rows, n, 3)).astype(np.float32)
    pts = torch.from_numpy(pts)
    want = oops.farthest_point_sample(pts, m)
    got = pointnet2_utils.farthest_point_sample(cuda(pts), m).cpu()
    assert torch.equal(got, want)
    if HAVE_REF:
        assert torch.equal(ref.farthest_point_sample(cuda(pts), m).cpu(), want)


def test_three_nn_bit_exact_on_voxel_grids():
    rng = np.random.default_rng(9)
    # voxel-centre coordinates: (idx + 0.5) * vs + lo in fp32 -> equidistant neighbours differ only
    # by rounding noise, so the FMA contraction order decides the indices
    vs, lo = np.float32([0.32, 0.32, 0.1875]), np.float32([-74.88, -74.88, -2.0])
    idx_u = rng.integers(0, 468, size=(512, 27, 3)) % np.array([468, 468, 32])
    idx_k = idx_u[:, :12] + rng.integers(-1, 2, size=(512, 12, 3))
    unknown = torch.from_numpy(((idx_u.astype(np.float32) + np.float32(0.5)) * vs + lo).astype(np.float32))
    known = torch.from_numpy(((idx_k.astype(np.float32) + np.float32(0.5)) * vs + lo).astype(np.float32))
    known[:, 8:] = 0.0  # padded query slots sit at the world origin
    wd, wi = oops.three_nn(unknown, known)
    gd, gi = pointnet2_utils.three_nn(cuda(unknown), cuda(known))
    assert torch.equal(gi.cpu(), wi)
    assert torch.equal(gd.cpu() ** 2, wd ** 2) or torch.allclose(gd.cpu(), wd, rtol=1e-6, atol=0)
    if HAVE_REF:
        rd, ri = ref.three_nn(cuda(unknown), cuda(known))
        assert torch.equal(ri.cpu(), wi) and torch.equal(rd.cpu(), gd.cpu())
    # fewer than three known points: the unused bests stay at +inf / index 0
    gd, gi = pointnet2_utils.three_nn(cuda(unknown[:4]), cuda(known[:4, :2].contiguous()))
    wd, wi = oops.three_nn(unknown[:4], known[:4, :2].contiguous())
    assert torch.equal(gi.cpu(), wi) and torch.isinf(gd[..., 2]).all()


@pytest.mark.parametrize("C,ns", [(64, 32), (3, 27), (32, 125), (48, 5)])
def test_grouping_operation_forward_backward(C, ns):
    rng = np.random.default_rng(C * ns)
    fbc = torch.tensor([700, 300, 500], dtype=torch.int32)
    ibc = torch.tensor([40, 0, 25], dtype=torch.int32)
    feats = torch.from_numpy(rng.standard_normal((1500, C)).astype(np.float32))
    idx = torch.from_numpy(rng.integers(-1, 300, size=(65, ns)).astype(np.int32))
    want = oops.grouping_operation(feats, fbc, idx, ibc)
    f = cuda(feats).requires_grad_(True)
    got = mssvt_ops.grouping_operation(f, cuda(fbc), cuda(idx), cuda(ibc))
    assert torch.equal(got.detach().cpu(), want)  # pure copy: bit-exact
    if HAVE_REF:
        assert torch.equal(ref.grouping_operation(cuda(feats), cuda(fbc), cuda(idx), cuda(ibc)).cpu(), want)
    g = torch.from_numpy(rng.standard_normal(want.shape).astype(np.float32))
    got.backward(cuda(g))
    want_grad = oops.grouping_operation_grad(g, 1500, fbc, idx, ibc)
    assert torch.allclose(f.grad.cpu(), want_grad, rtol=1e-5, atol=1e-5)  # float atomics: order-free sum


def test_gather_and_group_points_bit_exact():
    rng = np.random.default_rng(2)
    feats = torch.from_numpy(rng.standard_normal((7, 5, 40)).astype(np.float32))
    idx = torch.from_numpy(rng.integers(0, 40, size=(7, 16)).astype(np.int32))
    assert torch.equal(pointnet2_utils.gather_operation(cuda(feats), cuda(idx)).cpu(), oops.gather_operation(feats, idx))
    idx3 = torch.from_numpy(rng.integers(0, 40, size=(7, 9, 3)).astype(np.int32))
    want = oops.group_points(feats, idx3)
    f = cuda(feats).requires_grad_(True)
    got = pointnet2_utils.grouping_operation(f, cuda(idx3))
    assert torch.equal(got.detach().cpu(), want)
    got.sum().backward()
    counts = torch.zeros(7, 40)
    for b in range(7):
        counts[b] = torch.bincount(idx3[b].reshape(-1).long(), minlength=40).float()
    assert torch.allclose(f.grad.cpu(), counts[:, None, :].expand(7, 5, 40))
    if HAVE_REF:
        assert torch.equal(ref.group_points(cuda(feats), cuda(idx3)).cpu(), want)
        assert torch.equal(ref.gather_operation(cuda(feats), cuda(idx)).cpu(), oops.gather_operation(feats, idx))


def test_reference_kernels_were_present_on_this_box():
    """oracle/_ref (the reference's own kernels, built unmodified in the build container) must
    travel to the GPU box; without it the cross-checks above silently reduce to oracle-only."""
    assert HAVE_REF, "oracle/_ref/libmssvt_ref.so missing: run `make -C oracle ref` where /root/reference exists"


@pytest.mark.parametrize("rows,k", [(64, 32), (128, 64), (8, 8)])
def test_pack_operand_tf32_layout_and_rounding(rows, k):
    """mssvt_pack_operand_tf32: K-major 8-row x 16-byte core matrices (chunk c of row n at byte
    c * rows * 16 + (n / 8) * 128 + (n % 8) * 16), values rounded to TF32 (round to nearest, ties away)"""
    from mssvt_b200._lib import call, ptr, stream
    torch.manual_seed(rows)
    w = torch.randn(rows, k)
    def rna(a):   # cvt.rna.tf32.f32: round to nearest, ties away, 10 explicit mantissa bits
        bits = a.view(np.uint32).astype(np.uint64)
        return (((bits + 0x1000) & 0xFFFFE000) & 0xFFFFFFFF).astype(np.uint32).view(np.float32)

    def layout(m):
        want = np.zeros(rows * k, np.float32)
        for n in range(rows):
            for c in range(k // 4):
                at = (c * rows * 16 + (n // 8) * 128 + (n % 8) * 16) // 4
                want[at:at + 4] = m[n, 4 * c:4 * c + 4]
        return want

    hi = rna(w.numpy())
    out = torch.empty(rows * k, device="cuda")
    call("mssvt_pack_operand_tf32", ptr(w.cuda().contiguous()), rows, k, 1, ptr(out), stream())
    assert np.array_equal(out.cpu().numpy(), layout(hi))
    # terms = 3: [hi | lo] with lo = tf32(w - hi); hi + lo reproduces w to ~2^-21
    out3 = torch.empty(2 * rows * k, device="cuda")
    call("mssvt_pack_operand_tf32", ptr(w.cuda().contiguous()), rows, k, 3, ptr(out3), stream())
    got = out3.cpu().numpy()
    lo = rna((w.numpy() - hi).astype(np.float32))
    assert np.array_equal(got[:rows * k], layout(hi)) and np.array_equal(got[rows * k:], layout(lo))
    assert np.abs(hi + lo - w.numpy()).max() <= 2.0 ** -20 * np.abs(w.numpy()).max()


def test_attention_tile_plan_invariants():
    """mssvt_attention_tiles: per scale, every window with a real query lies in exactly one tile; a tile
    holds <= 128 rows (distinct keys + real queries), <= 48 queries, <= 32 windows, <= 2048 score slots; offsets
    are prefix sums; the row table names the window and the kind of every row"""
    from mssvt_b200.config import block_cfg
    from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformerBlock
    from mssvt_b200.mssvt_utils import SparseTensor
    from mssvt_b200.synth import S0_RANGE, S0_VOXEL
    cfg = block_cfg()
    blk = MixedScaleSparseTransformerBlock(cfg, 64, 128, 64, [2, 2], drop_path=0.0, window_size=cfg.window_size,
                                           cbs_pattern=1).cuda().eval()
    feats, coords = synth_frame(5, 30000, crop=0.45)
    sp = SparseTensor(features=torch.from_numpy(feats).cuda(), indices=torch.from_numpy(coords).cuda(),
                      spatial_shape=list(S0_GRID), voxel_size=list(S0_VOXEL), point_cloud_range=list(S0_RANGE),
                      batch_size=1, hash_size=400000, map_table=None, gather_dict=None)
    g = blk.geometry(sp)
    tiles, tile_count, win_rec, win_ctr, tile_rows = (t.cpu() for t in blk._tile_plan(sp, g, 2))
    W = int(g["total"].item())
    meta, q_base = g["meta"][:W].cpu(), g["q_base"][:W + 1].cpu()
    heads = 2
    for s in range(2):
        nqr = meta[:, 0]
        nrep = torch.where(nqr > 0, meta[:, 2 + s] & 0xff, torch.zeros_like(nqr))
        rec = win_rec[s, :W]
        assert torch.equal(rec[:, 0], q_base[:W])
        assert torch.equal(rec[:, 1] & 0xff, nqr) and torch.equal((rec[:, 1] >> 8) & 0xff, nrep)
        seen = torch.zeros(W, dtype=torch.int32)
        for t in range(int(tile_count[s])):
            ws, nw = (int(v) for v in tiles[s, t])
            assert 0 < nw <= 32 and ws + nw <= W
            seen[ws:ws + nw] += 1
            r, q = nrep[ws:ws + nw], nqr[ws:ws + nw]
            assert int(r.sum() + q.sum()) <= 128 and int(q.sum()) <= 48 and int((r * q).sum()) * heads <= 2048
            # row table: keys of window l, then (behind all keys) the queries of window l; the rest idle
            want_rows = torch.zeros(128, dtype=torch.uint8)
            nT, at_k, at_q = int(r.sum()), 0, 0
            for l in range(nw):
                want_rows[at_k:at_k + int(r[l])] = l | 0x40
                want_rows[nT + at_q:nT + at_q + int(q[l])] = l | 0x80
                at_k, at_q = at_k + int(r[l]), at_q + int(q[l])
            assert torch.equal(tile_rows[s, t], want_rows)
            off = rec[ws:ws + nw, 2]
            zero = torch.zeros(1, dtype=r.dtype)
            assert torch.equal(off & 0xff, torch.cat([zero, r.cumsum(0)[:-1]]).int())
            assert torch.equal((off >> 8) & 0xff, torch.cat([zero, q.cumsum(0)[:-1]]).int())
            assert torch.equal(off >> 16, torch.cat([zero, (r * q * heads).cumsum(0)[:-1]]).int())
        assert bool((seen[nrep > 0] == 1).all()) and int(seen.max()) <= 1
    cell = torch.tensor([S0_VOXEL[i] * 3 for i in range(3)])
    want = (g["win_list"][:W].cpu()[:, [3, 2, 1]].float() + 0.5) * cell + torch.tensor(S0_RANGE[:3])
    assert torch.allclose(win_ctr[:W, :3], want, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("rows", [1, 31, 128, 1000, 150001])
def test_tma_copy_rows(rows):
    """tensor-map row movement of the tensor-core FFN (tma.cuh): box loads / swizzled boxes / box stores copy a
    (rows, 64) matrix bit for bit and never write past the last row (partial tiles are clipped by the tensor map)"""
    import torch
    from mssvt_b200._lib import call, ptr, stream
    torch.manual_seed(rows)
    src = torch.randn(rows, 64, device="cuda")
    dst = torch.full((rows + 40, 64), -7.0, device="cuda")
    call("mssvt_tma_copy_rows", ptr(src), ptr(dst), rows, stream())
    torch.cuda.synchronize()
    assert torch.equal(dst[:rows], src)
    assert bool((dst[rows:] == -7.0).all())


@pytest.mark.gpu
@pytest.mark.parametrize("cap,n,stride", [(1, 1, 1), (5000, 3777, 4), (150000, 39031, 4), (262144, 262144, 1), (300000, 299999, 2)])
def test_exclusive_scan(cap, n, stride):
    """mssvt_exclusive_scan, single-pass form (lists up to 262144) and two-pass form: dst[i] = sum of src[j * stride], j < i,
    for i in [0, n] with the count read on the device"""
    import torch
    from mssvt_b200._lib import call, ptr, stream
    g = torch.Generator().manual_seed(cap)
    src = torch.randint(0, 13, (cap, stride), generator=g, dtype=torch.int32).cuda()
    n_dev = torch.tensor([n], dtype=torch.int32, device="cuda")
    dst = torch.full((cap + 1,), -1, dtype=torch.int32, device="cuda")
    ws = torch.empty((cap + 1 + 1023) // 1024 + 1, dtype=torch.int32, device="cuda")
    call("mssvt_exclusive_scan", cap, ptr(n_dev), ptr(src), stride, ptr(dst), ptr(ws), stream())
    want = torch.zeros(n + 1, dtype=torch.int64)
    want[1:] = torch.cumsum(src[:n, 0].cpu().long(), 0)
    assert torch.equal(dst[:n + 1].cpu().long(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("n,crop", [(30000, 0.45), (6000, 0.12)])
def test_geometry_distinct_key_sets_without_fps(n, crop):
    """mssvt_block_geometry without k_row / k_mask (what the tensor-core attention asks for) derives the distinct keys of a
    window without running FPS whenever the list has at most K distinct positions; against the full form (FPS picks,
    bit-exact vs the reference elsewhere): same key SET per window and scale, same masked key, same multiplicity.
    The dense crop has 5^3 lists with more than K voxels, where the real FPS has to run in both forms."""
    import torch
    from mssvt_b200.config import block_cfg
    from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformerBlock
    from mssvt_b200.mssvt_utils import SparseTensor
    from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL, synth_frame
    cfg = block_cfg()
    blk = MixedScaleSparseTransformerBlock(cfg, 64, 128, 64, [2, 2], drop_path=0.0, window_size=cfg.window_size,
                                           cbs_pattern=1).cuda().eval()
    blk.precision = "tf32x3"
    feats, coords = synth_frame(7, n, crop=crop)
    sp = SparseTensor(features=torch.from_numpy(feats).cuda(), indices=torch.from_numpy(coords).cuda(),
                      spatial_shape=list(S0_GRID), voxel_size=list(S0_VOXEL), point_cloud_range=list(S0_RANGE),
                      batch_size=1, hash_size=400000, map_table=None, gather_dict=None)
    fast, full = blk.geometry(sp), blk.geometry(sp, keys=True)
    assert fast["k_row"] is None and full["k_row"] is not None
    W, K = int(full["total"].item()), blk.key_num_sample
    assert int(fast["total"].item()) == W
    mf, mu = fast["meta"][:W].cpu(), full["meta"][:W].cpu()
    assert torch.equal(mf, mu)                      # (#queries, #win1 voxels, nrep | nmask << 8 per scale)
    rf, ru = fast["rep_row"][:W].cpu(), full["rep_row"][:W].cpu()
    many = 0
    for s in range(2):
        nrep, nmask = (mu[:, 2 + s] & 0xff).tolist(), (mu[:, 2 + s] >> 8).tolist()
        for w in range(W):
            a, b = rf[w, s * K:s * K + nrep[w]].tolist(), ru[w, s * K:s * K + nrep[w]].tolist()
            live = nrep[w] - (1 if nmask[w] else 0)
            assert sorted(a[:live]) == sorted(b[:live]), (w, s)
            if nmask[w]:
                assert a[-1] == b[-1], (w, s)
            else:
                many += 1
    assert many > 0 or crop > 0.2        # (the dense crop exercises lists where FPS has to choose)
    for k in ("q_row", "win1_row", "nn_idx", "nn_w", "covered", "vox_slot"):
        assert torch.equal(fast[k][:W] if fast[k].shape[0] >= W and k not in ("covered", "vox_slot") else fast[k],
                           full[k][:W] if full[k].shape[0] >= W and k not in ("covered", "vox_slot") else full[k]), k
