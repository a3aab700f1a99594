"""CPU: host logic of the training path (mssvt_b200/mssvt_backbone.py::_window_lists, train_ops.WindowLists) and the
identity the ragged kernels rest on -- a softmax over padded slots that all hold ONE masked key equals a softmax over
the distinct keys with that key weighted by its multiplicity, for the outputs and for every gradient
(mssvt_utils.py:123-139 with the key lists of mssvt_backbone.py:260-300).  No compute call into the CUDA library."""
import numpy as np
import torch

from mssvt_b200.config import s0_model_cfg
from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer
from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL


class _FakeTensor:
    def __init__(self, n, batch_size):
        self.indices = torch.zeros((n, 4), dtype=torch.int32)
        self.batch_size = batch_size


def _fake_geometry(rng, cap, W, nq, K, cap1, N):
    """geometry dict with the layout of mssvt_block_geometry: rows past W hold garbage"""
    meta = rng.integers(-5, 1 << 20, (cap, 4)).astype(np.int32)
    q_row = rng.integers(-5, N, (cap, nq)).astype(np.int32)
    rep_row = rng.integers(0, N, (cap, 2 * K)).astype(np.int32)
    want = {"q": [], "k": [[], []]}
    for w in range(W):
        n_real = int(rng.integers(0, nq + 1))
        q_row[w] = -1
        q_row[w, :n_real] = rng.integers(0, N, n_real)
        meta[w, 0] = n_real
        want["q"] += [(int(r), w) for r in q_row[w, :n_real]]
        for s in range(2):
            nrep = int(rng.integers(1, K + 1))
            mult = int(rng.integers(0, 2)) * int(rng.integers(1, K))
            meta[w, 2 + s] = nrep | (mult << 8)
            if n_real:
                want["k"][s] += [(int(rep_row[w, s * K + j]), w, bool(mult > 0 and j == nrep - 1)) for j in range(nrep)]
    vox_slot = np.full(N, -1, np.int32)
    owners = rng.permutation(W * cap1)[:N // 2]
    vox_slot[rng.permutation(N)[:N // 2]] = owners
    g = {"cap": cap, "meta": torch.from_numpy(meta), "q_row": torch.from_numpy(q_row), "rep_row": torch.from_numpy(rep_row),
         "total": torch.tensor([W], dtype=torch.int32), "win_count": torch.tensor([W, W, 0], dtype=torch.int32),
         "vox_slot": torch.from_numpy(vox_slot),
         "nn_idx": torch.from_numpy(rng.integers(0, nq, (cap, cap1, 3)).astype(np.uint8)),
         "nn_w": torch.from_numpy(rng.random((cap, cap1, 3)).astype(np.float32))}
    return g, want


def test_window_lists_csr_matches_brute_force():
    rng = np.random.default_rng(0)
    model = MixedScaleSparseTransformer(s0_model_cfg(), 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE))
    blk = model.backbone[0]
    nq, K, cap1, N, cap, W = 14, blk.key_num_sample, blk.max_num_win1, 400, 120, 77
    g, want = _fake_geometry(rng, cap, W, nq, K, cap1, N)
    L = blk._window_lists(_FakeTensor(N, 1), g)
    assert L["W"] == W
    assert list(zip(L["q_rows"].tolist(), L["q_win"].tolist())) == want["q"]
    for s in range(2):
        rows, k_win, masked, lists = L["groups"][s]
        assert list(zip(rows.tolist(), k_win.tolist(), masked.tolist())) == want["k"][s]
        # CSR offsets are consistent with the row lists
        ko, qo = lists.key_off.tolist(), lists.q_off.tolist()
        for w in range(W):
            assert k_win.tolist()[ko[w]:ko[w + 1]] == [w] * (ko[w + 1] - ko[w])
            assert L["q_win"].tolist()[qo[w]:qo[w + 1]] == [w] * (qo[w + 1] - qo[w])
        assert lists.num_keys == len(want["k"][s]) and lists.num_queries == len(want["q"])
    # three-NN map in compact query ids: -2 = uncovered voxel, -1 = padded query slot
    src, slot = L["merge_src"], g["vox_slot"].long()
    qo = L["groups"][0][3].q_off.tolist()
    for v in range(N):
        if slot[v] < 0:
            assert src[v].tolist() == [-2, -2, -2]
            continue
        w = int(slot[v]) // cap1
        for j in range(3):
            i = int(g["nn_idx"].reshape(-1, 3)[slot[v], j])
            assert int(src[v, j]) == (qo[w] + i if i < int(g["meta"][w, 0]) else -1)
    assert L is blk._window_lists(_FakeTensor(N, 1), g)          # cached with the geometry


def test_masked_key_multiplicity_identity_outputs_and_gradients():
    torch.manual_seed(0)
    hd, nk, mult = 16, 5, 9
    q = torch.randn(3, hd, dtype=torch.float64, requires_grad=True)
    k = torch.randn(nk, hd, dtype=torch.float64, requires_grad=True)
    v = torch.randn(nk, hd, dtype=torch.float64, requires_grad=True)
    go = torch.randn(3, hd, dtype=torch.float64)
    # padded form: the last key repeated `mult` times, each copy with the additive -100 of the reference
    idx = torch.cat((torch.arange(nk - 1), torch.full((mult,), nk - 1)))
    bias = torch.cat((torch.zeros(nk - 1), torch.full((mult,), -100.0))).double()
    out_p = torch.softmax(q @ k[idx].T * hd ** -0.5 + bias, -1) @ v[idx]
    gp = torch.autograd.grad(out_p, (q, k, v), go)
    # compact form: distinct keys, the masked one weighted by its multiplicity
    s = q @ k.T * hd ** -0.5 + torch.cat((torch.zeros(nk - 1), torch.tensor([-100.0]))).double()
    w = torch.exp(s - s.max(-1, keepdim=True)[0]) * torch.cat((torch.ones(nk - 1), torch.tensor([float(mult)]))).double()
    out_c = (w / w.sum(-1, keepdim=True)) @ v
    gc = torch.autograd.grad(out_c, (q, k, v), go)
    assert torch.allclose(out_p, out_c, atol=1e-12)
    for a, b in zip(gp, gc):
        assert torch.allclose(a, b, atol=1e-12)


def test_compress_lists_match_brute_force():
    """keys of a pillar window = its voxels + one pad key (row -1) standing for the padded slots (quirk Q6)"""
    rng = np.random.default_rng(3)
    model = MixedScaleSparseTransformer(s0_model_cfg(), 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE))
    blk = model.backbone[3]
    n1, cap, W, B = blk.max_num_win1, 90, 61, 2
    k_row = rng.integers(-9, 1000, (cap, n1)).astype(np.int32)            # rows past W: garbage
    want_rows, want_win, want_off, want_mult = [], [], [0], []
    for w in range(W):
        c = int(rng.integers(1, n1 + 1))
        k_row[w] = -1
        k_row[w, :c] = rng.integers(0, 1000, c)
        rows = [int(v) for v in k_row[w, :c]] + ([-1] if c < n1 else [])
        want_rows += rows
        want_win += [w] * len(rows)
        want_off.append(want_off[-1] + len(rows))
        want_mult.append(n1 - c)
    win_count = torch.tensor([30, 31, W, 0], dtype=torch.int32)           # per-sample counts, total, dropped
    got_w, rows, k_win, lists = blk._compress_lists_torch(torch.from_numpy(k_row), win_count, B, cap, n1)
    assert got_w == W and rows.tolist() == want_rows and k_win.tolist() == want_win
    assert lists.key_off.tolist()[:W + 1] == want_off and lists.key_mult.tolist()[:W] == want_mult
    assert lists.q_off.tolist()[:W + 1] == list(range(W + 1)) and lists.q_win.tolist() == list(range(W))
    win_count[B + 1] = 4
    try:
        blk._compress_lists_torch(torch.from_numpy(k_row), win_count, B, cap, n1)
        assert False, "dropped windows must raise"
    except RuntimeError as e:
        assert "max_num_wins" in str(e)
