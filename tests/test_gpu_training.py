"""GPU: the hand-written forward / backward kernels of the training path (csrc/train.cu through train_ops.py).

1. op level: mssvt_ragged_attention_fwd / _bwd against torch autograd over the PADDED form of the same windows
   (what mssvt_utils.py:100-157 computes: softmax(q k^T * scale - 100 * key_mask) v with the masked slots all holding
   one key), float64 on the torch side; mssvt_interp_merge_fwd / _bwd against torch indexing.
2. module level: the ragged training path of the blocks against the padded autograd path of the same module
   (features, input gradient, every parameter gradient) on the reference's golden configurations and on a
   20 000-voxel frame.
Bar: 1e-4 * max|ref| (fp32 arithmetic on both sides, summation order differs)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from mssvt_b200 import mssvt_backbone
from mssvt_b200.config import s0_model_cfg
from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer
from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL, synth_frame
from mssvt_b200.train_ops import (WindowLists, embed_rows, interp_merge, layer_norm_rows, linear_rows,
                                  ragged_window_attention, segment_max)

pytestmark = pytest.mark.gpu
TOL = 1e-4


def random_lists(rng, W, max_q, max_k):
    """windows with 0..max_q queries and 1..max_k keys (none for a window without a query), ~half with a masked key"""
    nq = rng.integers(0, max_q + 1, W)
    nk = np.where(nq > 0, rng.integers(1, max_k + 1, W), 0)
    mult = np.where((nq > 0) & (rng.random(W) < 0.5), rng.integers(1, 40, W), 0)
    q_off, k_off = np.concatenate(([0], np.cumsum(nq))), np.concatenate(([0], np.cumsum(nk)))
    q_win, k_win = np.repeat(np.arange(W), nq), np.repeat(np.arange(W), nk)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return WindowLists(t(q_off), t(q_win), t(k_off), t(k_win), t(mult)), nq, nk, mult


def padded_attention(q, kv, lists, nq, nk, mult, heads, scale):
    """the reference's dense form on padded tensors built from the same lists (float64)"""
    W, D = len(nq), q.shape[1]
    Q, Kn = int(nq.max()), int((nk + mult).max())
    qo, ko = lists.q_off.cpu().numpy(), lists.key_off.cpu().numpy()
    q_idx = torch.full((W, Q), -1, dtype=torch.long)
    k_idx = torch.zeros((W, Kn), dtype=torch.long)
    k_mask = torch.ones((W, Kn), dtype=torch.bool)       # True = masked (-100)
    k_pad = torch.ones((W, Kn), dtype=torch.bool)        # True = slot does not exist at all (-inf)
    for w in range(W):
        q_idx[w, :nq[w]] = torch.arange(qo[w], qo[w + 1])
        n = nk[w]
        if n == 0:
            continue
        live = n - (1 if mult[w] > 0 else 0)
        k_idx[w, :live] = torch.arange(ko[w], ko[w] + live)
        k_mask[w, :live] = False
        k_pad[w, :live] = False
        if mult[w] > 0:                                   # the masked key, once per slot it stands for
            k_idx[w, live:live + mult[w]] = ko[w + 1] - 1
            k_pad[w, live:live + mult[w]] = False
    dev = q.device
    q_idx, k_idx, k_mask, k_pad = q_idx.to(dev), k_idx.to(dev), k_mask.to(dev), k_pad.to(dev)
    hd = D // heads
    qp = torch.cat((q, q.new_zeros(1, D)))[q_idx].view(W, Q, heads, hd).permute(0, 2, 1, 3)
    kvp = kv[k_idx]                                        # (W, Kn, 2D)
    kp = kvp[..., :D].reshape(W, Kn, heads, hd).permute(0, 2, 1, 3)
    vp = kvp[..., D:].reshape(W, Kn, heads, hd).permute(0, 2, 1, 3)
    s = (qp * scale) @ kp.transpose(-2, -1) + (k_mask.to(q.dtype) * -100.0).view(W, 1, 1, Kn)
    s = s.masked_fill(k_pad.view(W, 1, 1, Kn), float("-inf"))
    has_keys = torch.from_numpy(nk > 0).to(dev).view(W, 1, 1, 1)
    p = torch.softmax(torch.where(has_keys, s, torch.zeros_like(s)), -1)
    o = (p @ vp).permute(0, 2, 1, 3).reshape(W, Q, D)
    return o[q_idx >= 0]                                   # (#queries, D), window-major like the compact form


@pytest.mark.parametrize("heads,hd", [(2, 16), (4, 8), (1, 32), (4, 16), (2, 32)])
def test_ragged_attention_forward_backward_vs_padded_torch(heads, hd):
    rng = np.random.default_rng(heads * 100 + hd)
    lists, nq, nk, mult = random_lists(rng, 300, 14, 33)
    D = heads * hd
    torch.manual_seed(hd)
    q = torch.randn(lists.num_queries, D, device="cuda") * 1.5
    kv = torch.randn(lists.num_keys, 2 * D, device="cuda") * 1.5
    go = torch.randn(lists.num_queries, D, device="cuda")
    scale = hd ** -0.5
    q1, kv1 = q.clone().requires_grad_(True), kv.clone().requires_grad_(True)
    out = ragged_window_attention(q1, kv1, lists, heads, scale)
    out.backward(go)
    q2, kv2 = q.double().requires_grad_(True), kv.double().requires_grad_(True)
    ref = padded_attention(q2, kv2, lists, nq, nk, mult, heads, scale)
    ref.backward(go.double())
    for got, want, name in ((out, ref, "out"), (q1.grad, q2.grad, "dq"), (kv1.grad, kv2.grad, "dkv")):
        err = (got.double() - want).abs().max().item()
        assert err <= TOL * want.abs().max().item(), (name, err, want.abs().max().item())
    # deterministic: no atomics in the attention backward
    q3, kv3 = q.clone().requires_grad_(True), kv.clone().requires_grad_(True)
    ragged_window_attention(q3, kv3, lists, heads, scale).backward(go)
    assert torch.equal(q3.grad, q1.grad) and torch.equal(kv3.grad, kv1.grad)


def test_ragged_attention_empty_and_single():
    e = torch.zeros(0, dtype=torch.int32, device="cuda")
    lists = WindowLists(torch.zeros(1, dtype=torch.int32, device="cuda"), e, torch.zeros(1, dtype=torch.int32, device="cuda"), e, e)
    q = torch.zeros((0, 32), device="cuda", requires_grad=True)
    kv = torch.zeros((0, 64), device="cuda", requires_grad=True)
    out = ragged_window_attention(q, kv, lists, 2, 0.25)
    assert out.shape == (0, 32)
    out.sum().backward()
    # one window, one query, one key that is the masked one: softmax over identical slots = the key's value
    t = lambda *a: torch.tensor(a, dtype=torch.int32, device="cuda")
    lists = WindowLists(t(0, 1), t(0), t(0, 1), t(0), t(7))
    q, kv = torch.randn(1, 32, device="cuda"), torch.randn(1, 64, device="cuda")
    out = ragged_window_attention(q, kv, lists, 2, 0.25)
    assert torch.allclose(out, kv[:, 32:], atol=1e-6)


def test_interp_merge_forward_backward_vs_torch():
    rng = np.random.default_rng(5)
    N, R, C = 5000, 700, 64
    src = rng.integers(-1, R, (N, 3)).astype(np.int32)
    src[rng.random(N) < 0.2] = -2
    src_t = torch.from_numpy(src).cuda()
    torch.manual_seed(1)
    w = torch.rand(N, 3, device="cuda")
    rows, x, go = torch.randn(R, C, device="cuda"), torch.randn(N, C, device="cuda"), torch.randn(N, C, device="cuda")
    r1, x1 = rows.clone().requires_grad_(True), x.clone().requires_grad_(True)
    out = interp_merge(r1, x1, src_t, w)
    out.backward(go)
    r2, x2 = rows.double().requires_grad_(True), x.double().requires_grad_(True)
    pad = torch.cat((r2, r2.new_zeros(1, C)))
    idx = torch.where(src_t < 0, torch.full_like(src_t, R), src_t).long()
    blend = (pad[idx.reshape(-1)].view(N, 3, C) * w.double().unsqueeze(-1)).sum(1)
    ref = torch.where((src_t[:, :1] == -2), x2, blend)
    ref.backward(go.double())
    for got, want, name in ((out, ref, "out"), (r1.grad, r2.grad, "drows"), (x1.grad, x2.grad, "dx")):
        err = (got.double() - want).abs().max().item()
        assert err <= TOL * want.abs().max().item(), (name, err)


@pytest.mark.parametrize("C", [64, 128])
def test_layernorm_rows_forward_backward_vs_torch(C):
    torch.manual_seed(C)
    n = 20011
    x = torch.randn(n, C, device="cuda") * 2 + 0.5
    go = torch.randn(n, C, device="cuda")
    norm = torch.nn.LayerNorm(C).cuda()
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.uniform_(-0.5, 0.5)
    x1 = x.clone().requires_grad_(True)
    y = layer_norm_rows(norm, x1)
    y.backward(go)
    got = (y.detach(), x1.grad.clone(), norm.weight.grad.clone(), norm.bias.grad.clone())
    norm.zero_grad()
    ref_norm = torch.nn.LayerNorm(C).cuda().double()
    ref_norm.load_state_dict({k: v.double() for k, v in norm.state_dict().items()})
    x2 = x.double().requires_grad_(True)
    y2 = ref_norm(x2)
    y2.backward(go.double())
    want = (y2.detach(), x2.grad, ref_norm.weight.grad, ref_norm.bias.grad)
    for a, b, name in zip(got, want, ("y", "dx", "dgamma", "dbeta")):
        err = (a.double() - b).abs().max().item()
        assert err <= TOL * b.abs().max().item(), (name, err, b.abs().max().item())


def test_embed_rows_forward_backward_vs_torch():
    """gather + one-layer positional embedding of three row sets (queries over all channels, two key groups over a
    32-channel slice each, masked keys, pad rows) against torch indexing + linear, float64"""
    rng = np.random.default_rng(9)
    N, W, C = 3000, 400, 64
    torch.manual_seed(2)
    xn = torch.randn(N, C, device="cuda")
    xyz = torch.randn(N, 3, device="cuda") * 10
    centre = torch.randn(W, 3, device="cuda") * 10
    pw, pb = torch.randn(C, 6, device="cuda") * 0.3, torch.randn(C, device="cuda") * 0.3
    sets = []
    for n_rows, c0, c1, with_mask, part in ((900, 0, 64, False, "both"), (7001, 0, 32, True, "both"), (5000, 32, 64, True, "both"),
                                            (3001, 0, 64, False, "x"), (3001, 0, 64, False, "pos")):
        rows = rng.integers(0, N, n_rows)
        rows[rng.random(n_rows) < 0.05] = -1
        win = rng.integers(0, W, n_rows)
        masked = torch.from_numpy(rng.random(n_rows) < 0.3).cuda() if with_mask else None
        sets.append((torch.from_numpy(rows).cuda(), torch.from_numpy(win).cuda(), masked, c0, c1, part))
    gos = [torch.randn(t[0].shape[0], t[4] - t[3], device="cuda") for t in sets]
    a1, w1, b1 = xn.clone().requires_grad_(True), pw.clone().requires_grad_(True), pb.clone().requires_grad_(True)
    outs = embed_rows(a1, w1, b1, xyz, centre, sets)
    torch.autograd.backward(outs, gos)
    a2, w2, b2 = xn.double().requires_grad_(True), pw.double().requires_grad_(True), pb.double().requires_grad_(True)
    refs = []
    for rows, win, masked, c0, c1, part in sets:
        idx = torch.where(rows < 0, torch.full_like(rows, N), rows)
        ctr = centre.double()[win]
        rel = torch.cat((xyz.double(), xyz.new_zeros(1, 3).double()))[idx] - ctr
        if masked is not None:
            rel = rel * (~masked).unsqueeze(1)
        pos = torch.cat((rel, ctr), 1)
        ref_x = torch.cat((a2, a2.new_zeros(1, C)))[idx][:, c0:c1]
        ref_p = torch.relu(torch.nn.functional.linear(pos, w2[c0:c1], b2[c0:c1]))
        refs.append(ref_x if part == "x" else ref_p if part == "pos" else ref_x + ref_p)
    torch.autograd.backward(refs, [g.double() for g in gos])
    for o, r in zip(outs, refs):
        assert (o.double() - r).abs().max().item() <= TOL * r.abs().max().item()
    for got, want, name in ((a1.grad, a2.grad, "dxn"), (w1.grad, w2.grad, "dw"), (b1.grad, b2.grad, "db")):
        err = (got.double() - want).abs().max().item()
        assert err <= TOL * want.abs().max().item(), (name, err, want.abs().max().item())


@pytest.mark.parametrize("K", [32, 64, 128])
@pytest.mark.parametrize("N", [32, 64, 128])
def test_linear_rows_forward_backward_vs_torch_float64(K, N):
    """mssvt_linear_rows_fwd / _wgrad (split-TF32 mma.sync) against nn.Linear in float64: a row count that is no multiple
    of any tile, x as a column slice of a wider matrix, with and without the fused ReLU; fp32-grade bar"""
    torch.manual_seed(K * 7 + N)
    R = 70001
    wide = torch.randn(R, K + 32, device="cuda")
    layer = torch.nn.Linear(K, N).cuda()
    ref = torch.nn.Linear(K, N).cuda().double()
    ref.load_state_dict({k: v.double() for k, v in layer.state_dict().items()})
    go = torch.randn(R, N, device="cuda")
    for relu in (False, True):
        layer.zero_grad()
        ref.zero_grad()
        w1 = wide.clone().requires_grad_(True)
        y = linear_rows(layer, w1[:, 32:], relu=relu)
        y.backward(go)
        w2 = wide.double().requires_grad_(True)
        y2 = ref(w2[:, 32:])
        if relu:                       # (the kernel's own sign pattern: outputs within rounding of 0 may differ in sign)
            assert ((y > 0) == (y2 > 0)).float().mean().item() > 0.9999
            y2 = y2 * (y > 0)
        y2.backward(go.double())
        for got, want, name in ((y, y2, "y"), (w1.grad, w2.grad, "dx"), (layer.weight.grad, ref.weight.grad, "dw"),
                                (layer.bias.grad, ref.bias.grad, "db")):
            err = (got.double() - want).abs().max().item()
            assert err <= 2e-5 * want.abs().max().item(), (name, relu, err, want.abs().max().item())
    # deterministic weight gradient (fixed-order reduction of the per-CTA partial sums)
    g1 = layer.weight.grad.clone()
    layer.zero_grad()
    linear_rows(layer, wide[:, 32:], relu=True).backward(go)
    assert torch.equal(g1, layer.weight.grad)


def test_segment_max_forward_backward_vs_torch():
    rng = np.random.default_rng(4)
    lists, nq, nk, mult = random_lists(rng, 500, 1, 20)
    W, C = 500, 64
    torch.manual_seed(8)
    rows = torch.randn(lists.num_keys, C, device="cuda")
    go = torch.randn(W, C, device="cuda")
    r1 = rows.clone().requires_grad_(True)
    out = segment_max(r1, lists, W)
    out.backward(go)
    r2 = rows.clone().requires_grad_(True)
    ko = lists.key_off.cpu().tolist()
    ref = torch.stack([r2[ko[w]:ko[w + 1]].max(0)[0] if ko[w + 1] > ko[w] else r2.new_zeros(C) for w in range(W)])
    ref.backward(go)
    assert torch.equal(out, ref.detach())
    assert torch.equal(r1.grad, r2.grad)


def _train_run(cfg, state, grid, pc_range, feats, coords, batch, path, monkeypatch, train=True):
    monkeypatch.setattr(mssvt_backbone, "TRAIN_PATH", path)
    model = MixedScaleSparseTransformer(cfg, feats.shape[1], list(grid), list(S0_VOXEL), list(pc_range))
    if state is not None:
        model.load_state_dict(state, strict=True)
    model = model.cuda()
    model.train(train)
    for m in model.modules():            # deterministic: no stochastic depth / dropout
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if hasattr(m, "drop_prob"):
            m.drop_prob = 0.0
    x = feats.cuda().requires_grad_(True)
    sp = model({"voxel_features": x, "voxel_coords": coords.cuda().float(), "batch_size": batch})["encoded_spconv_tensor"]
    (sp.features ** 2).sum().backward()
    return sp.features.detach(), sp.indices, x.grad.clone(), {n: p.grad.clone() for n, p in model.named_parameters()}


def _compare_runs(a, b):
    fa, ia, gxa, gpa = a
    fb, ib, gxb, gpb = b
    assert torch.equal(ia, ib)
    assert (fa - fb).abs().max().item() <= TOL * fb.abs().max().item()
    assert (gxa - gxb).abs().max().item() <= TOL * gxb.abs().max().item()
    for n in gpb:
        scale = gpb[n].abs().max().item()
        assert (gpa[n] - gpb[n]).abs().max().item() <= TOL * max(scale, 1e-6), n


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_ragged_training_path_vs_padded_autograd_golden_configs(name, monkeypatch):
    """hand-written kernels on the compact form == torch autograd over the reference's padded tensors: features, input
    gradient and the gradient of every parameter (S0 with patterns 1/0/2; the mixed configuration with capped lists,
    a block without interpolation, out_linear and a two-group compress block)"""
    blob, cfg, state = load_golden(name)
    feats, coords = torch.from_numpy(blob["voxel_features"]), torch.from_numpy(blob["voxel_coords"])
    args = (cfg, state, blob["grid"], blob["pc_range"], feats, coords, int(blob["batch_size"]))
    _compare_runs(_train_run(*args, "ragged", monkeypatch), _train_run(*args, "padded", monkeypatch))


def test_ragged_training_path_vs_padded_autograd_20k_frame(monkeypatch):
    feats, coords = synth_frame(11, 20000, crop=0.38)
    feats, coords = torch.from_numpy(feats), torch.from_numpy(coords)
    torch.manual_seed(3)
    ref_model = MixedScaleSparseTransformer(s0_model_cfg(cbs_patterns=(1, 0, 2)), 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE))
    state = {k: v.clone() for k, v in ref_model.state_dict().items()}
    args = (s0_model_cfg(cbs_patterns=(1, 0, 2)), state, S0_GRID, S0_RANGE, feats, coords, 1)
    _compare_runs(_train_run(*args, "ragged", monkeypatch), _train_run(*args, "padded", monkeypatch))


def test_ragged_training_path_batch_with_empty_sample_and_overflow(monkeypatch):
    """three samples, the middle one empty: ragged == padded; max_num_wins below the window count raises in training too"""
    feats, coords = synth_frame(5, 3000, batch_size=3, crop=0.3)
    keep = coords[:, 0] != 1
    feats, coords = torch.from_numpy(feats[keep]), torch.from_numpy(coords[keep])
    torch.manual_seed(5)
    cfg = s0_model_cfg(cbs_patterns=(2, 1, 0))
    ref_model = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE))
    state = {k: v.clone() for k, v in ref_model.state_dict().items()}
    args = (cfg, state, S0_GRID, S0_RANGE, feats, coords, 3)
    _compare_runs(_train_run(*args, "ragged", monkeypatch), _train_run(*args, "padded", monkeypatch))
    monkeypatch.setattr(mssvt_backbone, "TRAIN_PATH", "ragged")
    model = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE)).cuda().train()
    for b in model.backbone:
        b.max_num_wins = 50
    with pytest.raises(RuntimeError, match="max_num_wins"):
        model({"voxel_features": feats.cuda().requires_grad_(True), "voxel_coords": coords.cuda().float(), "batch_size": 3})


def test_device_window_lists_match_the_torch_lists():
    """csrc/train_lists.cu (mssvt_ragged_lists_*, mssvt_ragged_merge_map, mssvt_compress_lists_*) against the same lists
    from torch index operations, which tests/test_training_host.py checks against brute force on the CPU"""
    feats, coords = synth_frame(3, 6000, batch_size=2, crop=0.25)
    model = MixedScaleSparseTransformer(s0_model_cfg(cbs_patterns=(1, 0, 2)), 64, list(S0_GRID), list(S0_VOXEL),
                                        list(S0_RANGE)).cuda()
    sp = model._sparse_tensor(torch.from_numpy(feats).cuda(), torch.from_numpy(coords).cuda(), 2)
    for blk in model.backbone[:3]:
        g = blk.geometry(sp)
        L1 = blk._window_lists(sp, g)
        del g[("ragged",)]
        L2 = blk._window_lists_torch(sp, g)
        del g[("ragged",)]
        W = L1["W"]
        assert W == L2["W"] and W > 100
        assert torch.equal(L1["q_rows"].long(), L2["q_rows"]) and torch.equal(L1["q_win"].long(), L2["q_win"])
        assert torch.equal(L1["merge_src"], L2["merge_src"])
        cov = L1["merge_src"][:, 0] != -2
        assert torch.equal(L1["merge_w"][cov], L2["merge_w"][cov])
        for (r1, w1, m1, l1), (r2, w2, m2, l2) in zip(L1["groups"], L2["groups"]):
            assert r1.numel() > 0 and torch.equal(r1.long(), r2) and torch.equal(w1.long(), w2) and torch.equal(m1, m2)
            assert torch.equal(l1.key_off[:W + 1], l2.key_off[:W + 1]) and torch.equal(l1.q_off[:W + 1], l2.q_off[:W + 1])
            assert torch.equal(l1.key_mult[:W], l2.key_mult[:W])
    cb = model.backbone[3]
    grid, win_list, win_table, win_count, k_row, plan = cb.prepare(sp)
    a = cb._compress_lists(k_row, win_count, 2, win_list.shape[0], cb.max_num_win1)
    b = cb._compress_lists_torch(k_row, win_count, 2, win_list.shape[0], cb.max_num_win1)
    W = a[0]
    assert W == b[0] and torch.equal(a[1].long(), b[1]) and torch.equal(a[2].long(), b[2])
    assert torch.equal(a[3].key_off[:W + 1], b[3].key_off[:W + 1]) and torch.equal(a[3].key_mult[:W], b[3].key_mult[:W])
