"""GPU: the fused backbone (module API -> C-ABI -> sm_100a kernels) against the golden vectors of
the reference's Python layer and against the CPU oracle run live on the same seeded inputs.
Bar: every index tensor bit-exact; features within 1e-4 * max|ref| (exact-fp32 FFMA kernels;
only the summation order differs from the oracle's MKL GEMMs)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from oracle import backbone as orc
from mssvt_b200.config import s0_model_cfg
from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer
from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL, synth_frame

pytestmark = pytest.mark.gpu
FEATURE_TOL = 1e-4  # relative to max|reference|, fp32 mode


def run_product(cfg, state, grid, pc_range, feats, coords, batch):
    model = MixedScaleSparseTransformer(cfg, feats.shape[1], list(grid), list(S0_VOXEL), list(pc_range))
    model.load_state_dict(state, strict=True)
    model = model.cuda().eval()
    with torch.no_grad():
        out = model({"voxel_features": feats.cuda(), "voxel_coords": coords.cuda().float(),
                     "batch_size": batch})
    assert out["encoded_spconv_tensor_stride"] == 1
    return model, out["encoded_spconv_tensor"]


def to_local(rows, win_b, v_start):
    """global feature rows -> per-sample indices of the reference lists (-1 stays -1)"""
    base = v_start[win_b.long()].unsqueeze(1)
    return torch.where(rows >= 0, rows - base, rows)


@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "bf16x3"])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_backbone_matches_reference_golden(name, precision):
    """the modes held to the fp32 bar: FFMA kernels and the two split-operand tensor-core forms"""
    blob, cfg, state = load_golden(name)
    cfg["PRECISION"] = precision
    feats, coords = torch.from_numpy(blob["voxel_features"]), torch.from_numpy(blob["voxel_coords"])
    model, sp = run_product(cfg, state, blob["grid"], blob["pc_range"], feats, coords, int(blob["batch_size"]))
    assert torch.equal(sp.indices.cpu(), torch.from_numpy(blob["out_indices"]))
    ref = torch.from_numpy(blob["out_features"])
    err = (sp.features.cpu() - ref).abs().max().item()
    assert err <= FEATURE_TOL * ref.abs().max().item(), err
    dense = sp.dense()
    assert dense.shape[0] == int(blob["batch_size"]) and dense.shape[1] == ref.shape[1]
    idx = sp.indices.long()
    assert torch.equal(dense[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]], sp.features)
    assert int((dense != 0).any(1).sum()) <= idx.shape[0]


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_block_geometry_maps_bit_exact(name):
    """sampled indices and gather maps of the first block vs the reference taps"""
    blob, cfg, state = load_golden(name)
    from mssvt_b200.mssvt_utils import SparseTensor
    model = MixedScaleSparseTransformer(cfg, blob["voxel_features"].shape[1], blob["grid"].tolist(),
                                        list(S0_VOXEL), blob["pc_range"].tolist()).cuda().eval()
    coords = torch.from_numpy(blob["voxel_coords"]).cuda()
    B = int(blob["batch_size"])
    sp = SparseTensor(torch.from_numpy(blob["voxel_features"]).cuda(), coords, blob["grid"].tolist(),
                      list(S0_VOXEL), blob["pc_range"].tolist(), B, cfg.HASH_SIZE)
    blk = model.backbone[0]
    g = blk.geometry(sp, taps=True)
    W = int(g["win_count"][B])
    assert W == blob["tap0/win_ind"].shape[0]
    win = g["win_list"][:W].cpu()
    assert torch.equal(win, torch.from_numpy(blob["tap0/win_ind"]))
    v_start = sp.sample_counts()[1].cpu()
    K = blk.key_num_sample
    loc = lambda t: to_local(t[:W].cpu(), win[:, 0], v_start)
    assert torch.equal(loc(g["q_row"]), torch.from_numpy(blob["tap0/q_ind"]))
    assert torch.equal(loc(g["win1_row"]), torch.from_numpy(blob["tap0/win1_ind"]))
    k_local = loc(g["k_row"])
    assert torch.equal(k_local[:, :K], torch.from_numpy(blob["tap0/k_ind_win1"]))
    assert torch.equal(k_local[:, K:], torch.from_numpy(blob["tap0/k_ind_win2"]))
    fps = g["fps_idx"][:W].cpu()
    assert torch.equal(fps[:, :K], torch.from_numpy(blob["tap0/fps_win1"]))
    assert torch.equal(fps[:, K:], torch.from_numpy(blob["tap0/fps_win2"]))
    mask = g["k_mask"][:W].cpu().bool()
    assert torch.equal(mask[:, :K], torch.from_numpy(blob["tap0/k_mask_win1"]))
    assert torch.equal(mask[:, K:], torch.from_numpy(blob["tap0/k_mask_win2"]))
    if "tap0/nn_idx" in blob:
        real = torch.from_numpy(blob["tap0/win1_ind"]) >= 0  # padded win1 slots go to the throw-away row
        got = g["nn_idx"][:W].cpu().int()
        assert torch.equal(got[real], torch.from_numpy(blob["tap0/nn_idx"])[real])


def test_s0_block_and_backbone_vs_live_oracle_20k():
    """BASELINE config 1 size: 20 k voxels, S0, batch 1 -- oracle computed live on the CPU"""
    feats, coords = synth_frame(0, 20000, crop=0.38)
    feats, coords = torch.from_numpy(feats), torch.from_numpy(coords)
    cfg = s0_model_cfg()
    torch.manual_seed(0)
    model = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE))
    state = {k: v.clone() for k, v in model.state_dict().items()}
    taps = []
    with torch.no_grad():
        want = orc.backbone_forward(state, cfg, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE), feats, coords, 1, taps=taps)
    _, sp = run_product(cfg, state, S0_GRID, S0_RANGE, feats, coords, 1)
    assert torch.equal(sp.indices.cpu(), want.indices)
    err = (sp.features.cpu() - want.features).abs().max().item()
    assert err <= FEATURE_TOL * want.features.abs().max().item(), err


def test_full_size_frame_properties_150k():
    """BASELINE config 2 size (too slow for the oracle in a unit test): size-independent checks"""
    feats, coords = synth_frame(1, 150000)
    feats, coords = torch.from_numpy(feats), torch.from_numpy(coords)
    cfg = s0_model_cfg()
    torch.manual_seed(0)
    model = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE)).cuda().eval()
    run = lambda f: model({"voxel_features": f.cuda(), "voxel_coords": coords.cuda().float(), "batch_size": 1})["encoded_spconv_tensor"]
    with torch.no_grad():
        a, b = run(feats), run(feats)
    assert torch.equal(a.features, b.features) and torch.equal(a.indices, b.indices)  # run-to-run identical
    assert torch.isfinite(a.features).all()
    # output rows = occupied (x, y) pillars, in first-occurrence order of the sorted voxels
    pillars = torch.unique(coords[:, 2].long() * 468 + coords[:, 3].long())
    assert a.indices.shape[0] == pillars.numel() and a.spatial_shape == [468, 468, 1]
    got = a.indices[:, 2].long() * 468 + a.indices[:, 3].long()
    assert torch.equal(torch.sort(got.cpu())[0], pillars)
    dense = a.dense()
    assert dense.shape == (1, 64, 1, 468, 468)
    assert torch.equal(dense[0, :, 0, a.indices[:, 2].long(), a.indices[:, 3].long()].t(), a.features)


TF32_TOL = 2e-3  # relative to max|reference|: FFN GEMMs with TF32 operands on tcgen05, fp32 accumulate
BF16_TOL, BF16_RMS = 2e-2, 5e-3  # bf16 operands (tcgen05.mma.kind::f16), fp32 accumulate


@pytest.mark.parametrize("n_rows,mode", [(1000, 1), (128, 0), (37, 1), (5000, 0)])
def test_tensor_core_ffn_matches_fp32_ffn(n_rows, mode):
    """mssvt_ffn_tc (tcgen05, TF32) against mssvt_ffn (FFMA, fp32) and a float64 PyTorch reference"""
    import ctypes
    from mssvt_b200._lib import call, ptr, stream
    from mssvt_b200.config import block_cfg
    from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformerBlock
    torch.manual_seed(n_rows)
    cfg = block_cfg()
    blk = MixedScaleSparseTransformerBlock(cfg, 64, 128, 64, [2, 2], drop_path=0.0, window_size=cfg.window_size,
                                           cbs_pattern=1).cuda().eval()
    with torch.no_grad():
        blk.norm2.weight.add_(0.2 * torch.randn_like(blk.norm2.weight))
        blk.norm2.bias.add_(0.2 * torch.randn_like(blk.norm2.bias))
    x = torch.randn(n_rows, 64, device="cuda")
    merged = torch.randn(n_rows, 64, device="cuda")
    covered = (torch.rand(n_rows, device="cuda") < 0.7).to(torch.uint8)
    u = merged if mode == 0 else torch.where(covered.bool()[:, None], merged + x, 2 * x)
    with torch.no_grad():
        ud = u.double()
        h = torch.nn.functional.layer_norm(ud, (64,), blk.norm2.weight.double(), blk.norm2.bias.double(), blk.norm2.eps)
        ref = ud + torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(
            h, blk.linear1.weight.double(), blk.linear1.bias.double())), blk.linear2.weight.double(), blk.linear2.bias.double())
        S, buf = blk._ffn_descriptor(mode)
        blk.precision = "fp32"
        y32 = blk._ffn(S, buf, n_rows, x, merged, covered)
        blk.precision = "tf32"
        ytc = blk._ffn(S, buf, n_rows, x, merged, covered)
        # fused epilogue: LayerNorm of the next block applied to y in the same kernel
        nxt = torch.nn.LayerNorm(64).cuda()
        nxt.weight.add_(0.3 * torch.randn_like(nxt.weight)); nxt.bias.add_(0.3 * torch.randn_like(nxt.bias))
        blk.__dict__["_next_norm1"] = nxt
        ytc2 = blk._ffn(S, buf, n_rows, x, merged, covered)
        y_keep, xn_next = blk.__dict__["_xn_for_next"]
        assert y_keep is ytc2 and torch.equal(ytc2, ytc)
        want = torch.nn.functional.layer_norm(ytc2, (64,), nxt.weight, nxt.bias, nxt.eps)
        assert (xn_next - want).abs().max().item() <= 1e-5 * want.abs().max().item()
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    assert (y32.double() - ref).abs().max().item() <= 1e-5 * scale
    err = (ytc.double() - ref).abs().max().item()
    assert err <= TF32_TOL * scale, (err, scale)
    assert err > 0  # it really is a different arithmetic path


@pytest.mark.parametrize("name", ["s0_b2_n1200"])
def test_backbone_tf32_mode_within_stated_tolerance(name):
    blob, cfg, state = load_golden(name)
    cfg["PRECISION"] = "tf32"
    feats, coords = torch.from_numpy(blob["voxel_features"]), torch.from_numpy(blob["voxel_coords"])
    model, sp = run_product(cfg, state, blob["grid"], blob["pc_range"], feats, coords, int(blob["batch_size"]))
    assert model.backbone[0].precision == "tf32"
    assert torch.equal(sp.indices.cpu(), torch.from_numpy(blob["out_indices"]))  # indices stay bit-exact
    ref = torch.from_numpy(blob["out_features"])
    err = (sp.features.cpu() - ref).abs().max().item()
    assert err <= TF32_TOL * ref.abs().max().item(), err


def test_height_compression_consumes_the_backbone_output():
    """map_to_bev mirror (height_compression.py:31-51): dense() -> (B, C*D, H, W), stride passed on"""
    from mssvt_b200.config import AttrDict
    from mssvt_b200.height_compression import HeightCompression
    blob, cfg, state = load_golden("s0_b2_n1200")
    feats, coords = torch.from_numpy(blob["voxel_features"]), torch.from_numpy(blob["voxel_coords"])
    model = MixedScaleSparseTransformer(cfg, feats.shape[1], list(blob["grid"]), list(S0_VOXEL), list(blob["pc_range"]))
    model.load_state_dict(state, strict=True)
    model = model.cuda().eval()
    bev = HeightCompression(AttrDict(NUM_BEV_FEATURES=64, COMPRESS_LAYER_NUMS=0)).cuda().eval()
    with torch.no_grad():
        out = bev(model({"voxel_features": feats.cuda(), "voxel_coords": coords.cuda().float(),
                         "batch_size": int(blob["batch_size"])}))
    sf = out["spatial_features"]
    sp = out["encoded_spconv_tensor"]
    B, (X, Y, Z) = int(blob["batch_size"]), [int(v) for v in sp.spatial_shape]
    assert sf.shape == (B, 64 * Z, Y, X) and out["spatial_features_stride"] == 1
    ref = torch.zeros((B, 64, Z, Y, X))
    idx = torch.from_numpy(blob["out_indices"]).long()
    ref[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]] = torch.from_numpy(blob["out_features"])
    err = (sf.cpu().view(B, 64, Z, Y, X) - ref).abs().max().item()
    assert err <= FEATURE_TOL * ref.abs().max().item(), err
    bev3 = HeightCompression(AttrDict(NUM_BEV_FEATURES=64 * Z)).cuda().eval()   # default 3-layer conv stack
    with torch.no_grad():
        assert bev3(out)["spatial_features"].shape == sf.shape


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_training_path_forward_matches_reference_golden(name):
    """the autograd (training) path computes the same features as the fused inference kernels / the reference"""
    blob, cfg, state = load_golden(name)
    feats, coords = torch.from_numpy(blob["voxel_features"]), torch.from_numpy(blob["voxel_coords"])
    model = MixedScaleSparseTransformer(cfg, feats.shape[1], list(blob["grid"]), list(S0_VOXEL), list(blob["pc_range"]))
    model.load_state_dict(state, strict=True)
    model = model.cuda().eval()          # eval: DropPath / Dropout are identities, gradients still flow
    x = feats.cuda().requires_grad_(True)
    sp = model({"voxel_features": x, "voxel_coords": coords.cuda().float(),
                "batch_size": int(blob["batch_size"])})["encoded_spconv_tensor"]
    assert sp.features.requires_grad
    assert torch.equal(sp.indices.cpu(), torch.from_numpy(blob["out_indices"]))
    ref = torch.from_numpy(blob["out_features"])
    err = (sp.features.detach().cpu() - ref).abs().max().item()
    assert err <= FEATURE_TOL * ref.abs().max().item(), (err, ref.abs().max().item())
    dense = sp.dense()
    assert dense.requires_grad and dense.shape[1] == ref.shape[1]
    (dense ** 2).mean().backward()
    assert x.grad is not None and torch.isfinite(x.grad).all() and x.grad.abs().max() > 0
    for n, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n


def test_training_path_gradients_cuda_gather_backward_vs_autograd_indexing(monkeypatch):
    """Same graph twice: row gathers through GroupingOperation (mssvt_group_features /
    mssvt_group_features_grad, the CUDA scatter-add) and through plain torch indexing (autograd's own
    backward).  Input and parameter gradients must agree (fp32 atomics: order-dependent rounding only)."""
    from mssvt_b200 import mssvt_backbone, mssvt_ops
    monkeypatch.setattr(mssvt_backbone, "TRAIN_PATH", "padded")   # (the path that gathers through GroupingOperation)
    blob, cfg, state = load_golden("s0_b2_n1200")
    feats, coords = torch.from_numpy(blob["voxel_features"]), torch.from_numpy(blob["voxel_coords"])

    def run():
        model = MixedScaleSparseTransformer(cfg, feats.shape[1], list(blob["grid"]), list(S0_VOXEL),
                                            list(blob["pc_range"]))
        model.load_state_dict(state, strict=True)
        model = model.cuda().train()
        for m in model.modules():            # deterministic: no stochastic depth / dropout
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
            if hasattr(m, "drop_prob"):
                m.drop_prob = 0.0
        x = feats.cuda().requires_grad_(True)
        sp = model({"voxel_features": x, "voxel_coords": coords.cuda().float(),
                    "batch_size": int(blob["batch_size"])})["encoded_spconv_tensor"]
        (sp.features ** 2).sum().backward()
        return x.grad.clone(), {n: p.grad.clone() for n, p in model.named_parameters()}

    gx, gp = run()

    def torch_gather(features, features_batch_cnt, idx, idx_batch_cnt):
        pad = torch.cat([features, features.new_zeros(1, features.shape[1])], 0)
        rows = torch.where(idx < 0, torch.full_like(idx, features.shape[0]), idx).long()
        return pad[rows].permute(0, 2, 1)                      # (M, C, ns)

    monkeypatch.setattr(mssvt_ops, "grouping_operation", torch_gather)
    rx, rp = run()
    assert (gx - rx).abs().max().item() <= 1e-4 * rx.abs().max().item()
    for n in rp:
        scale = rp[n].abs().max().item()
        assert (gp[n] - rp[n]).abs().max().item() <= 1e-4 * max(scale, 1e-6), n


def test_tf32_mode_vs_live_oracle_20k_and_fp32_mode_150k():
    """tensor-core mode at BASELINE config 1 size against the CPU oracle, and at config 2 size against
    the exact-fp32 kernels of the same library (the oracle takes seconds per 150 k frame)"""
    feats, coords = synth_frame(3, 20000, crop=0.38)
    feats, coords = torch.from_numpy(feats), torch.from_numpy(coords)
    cfg = s0_model_cfg()
    torch.manual_seed(0)
    model = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE))
    state = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        want = orc.backbone_forward(state, cfg, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE), feats, coords, 1)
    model = model.cuda().eval()
    model.set_precision("tf32")
    run = lambda f, c: model({"voxel_features": f.cuda(), "voxel_coords": c.cuda().float(),
                              "batch_size": 1})["encoded_spconv_tensor"]
    with torch.no_grad():
        sp = run(feats, coords)
    assert torch.equal(sp.indices.cpu(), want.indices)
    err = (sp.features.cpu() - want.features).abs().max().item()
    assert err <= TF32_TOL * want.features.abs().max().item(), err
    # full-size frame: both modes of the library on the same input
    feats, coords = synth_frame(4, 150000)
    feats, coords = torch.from_numpy(feats), torch.from_numpy(coords)
    with torch.no_grad():
        tc = run(feats, coords)
        tc_feat, tc_idx = tc.features.clone(), tc.indices.clone()
        model.set_precision("fp32")
        ex = run(feats, coords)
    assert torch.equal(tc_idx, ex.indices)
    err = (tc_feat - ex.features).abs().max().item()
    assert err <= TF32_TOL * ex.features.abs().max().item(), err
    assert err > 0


@pytest.mark.parametrize("n", [1, 37, 300])
def test_tiny_frames_in_every_mode(n):
    """a handful of voxels: single windows, tiles with one task, grids smaller than the SM count"""
    feats, coords = synth_frame(11, n, crop=0.05)
    feats, coords = torch.from_numpy(feats), torch.from_numpy(coords)
    cfg = s0_model_cfg()
    torch.manual_seed(0)
    model = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE))
    state = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        want = orc.backbone_forward(state, cfg, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE), feats, coords, 1)
    model = model.cuda().eval()
    for precision, tol in (("fp32", FEATURE_TOL), ("tf32x3", FEATURE_TOL), ("bf16x3", FEATURE_TOL), ("tf32", TF32_TOL), ("bf16", BF16_TOL)):
        model.set_precision(precision)
        with torch.no_grad():
            sp = model({"voxel_features": feats.cuda(), "voxel_coords": coords.cuda().float(),
                        "batch_size": 1})["encoded_spconv_tensor"]
        assert torch.equal(sp.indices.cpu(), want.indices)
        err = (sp.features.cpu() - want.features).abs().max().item()
        assert err <= tol * want.features.abs().max().item(), (precision, err)


@pytest.mark.parametrize("split", [False, True])
def test_cuda_graph_replay_matches_eager_forward(split):
    """MixedScaleSparseTransformer.capture: the graph (or the coordinate graph + feature graph pair), replayed
    on new contents of its static input buffers, gives exactly the eager results"""
    cfg = s0_model_cfg()
    cfg["PRECISION"] = "tf32"
    torch.manual_seed(0)
    model = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE)).cuda().eval()
    frames = [synth_frame(20 + i, 6000, crop=0.2) for i in range(3)]
    f0 = torch.from_numpy(frames[0][0]).cuda()
    c0 = torch.from_numpy(frames[0][1]).cuda()
    graphed = model.capture({"voxel_features": f0, "voxel_coords": c0, "batch_size": 1}, split=split)
    assert graphed.launches > 20 and (graphed.prepare_graph is not None) == split
    for feats, coords in frames[1:] + frames[:1]:
        f, c = torch.from_numpy(feats).cuda(), torch.from_numpy(coords).cuda()
        with torch.no_grad():
            want = model({"voxel_features": f, "voxel_coords": c, "batch_size": 1})["encoded_spconv_tensor"]
            want_f, want_i = want.features.clone(), want.indices.clone()
        f0.copy_(f)
        c0.copy_(c)
        if split:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                graphed.replay_prepare()           # coordinate pass on another stream, then the feature pass
            torch.cuda.current_stream().wait_stream(side)
            got = graphed.replay_features()
        else:
            got = graphed.replay()
        assert torch.equal(got.indices, want_i) and torch.equal(got.features, want_f)
        dense = got.dense()
        assert dense.shape == (1, 64, 1, 468, 468) and torch.isfinite(dense).all()


@pytest.mark.parametrize("name", ["s0_b2_n1200"])
def test_backbone_tf32x3_mode_meets_the_fp32_tolerance(name):
    """split-operand tensor-core mode (3xTF32): same kernels as tf32, fp32-grade results"""
    blob, cfg, state = load_golden(name)
    cfg["PRECISION"] = "tf32x3"
    feats, coords = torch.from_numpy(blob["voxel_features"]), torch.from_numpy(blob["voxel_coords"])
    model, sp = run_product(cfg, state, blob["grid"], blob["pc_range"], feats, coords, int(blob["batch_size"]))
    assert model.backbone[0].precision == "tf32x3"
    assert torch.equal(sp.indices.cpu(), torch.from_numpy(blob["out_indices"]))
    ref = torch.from_numpy(blob["out_features"])
    err = (sp.features.cpu() - ref).abs().max().item()
    assert err <= FEATURE_TOL * ref.abs().max().item(), err
    # and on a 20 k-voxel frame against the tf32 and fp32 modes of the library
    torch.manual_seed(0)
    model = MixedScaleSparseTransformer(s0_model_cfg(), 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE)).cuda().eval()
    f, c = synth_frame(9, 20000, crop=0.38)
    f, c = torch.from_numpy(f).cuda(), torch.from_numpy(c).cuda()
    outs = {}
    for mode in ("fp32", "tf32x3", "tf32"):
        model.set_precision(mode)
        with torch.no_grad():
            outs[mode] = model({"voxel_features": f, "voxel_coords": c, "batch_size": 1})["encoded_spconv_tensor"].features.clone()
    scale = outs["fp32"].abs().max().item()
    e3 = (outs["tf32x3"] - outs["fp32"]).abs().max().item() / scale
    e1 = (outs["tf32"] - outs["fp32"]).abs().max().item() / scale
    assert e3 <= 2e-5 and e3 < 0.1 * e1, (e3, e1)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_backbone_bf16_mode_within_stated_tolerance(name):
    """bf16 mode against the reference golden vectors: indices bit-exact, features within 2e-2 of max|ref| with
    rms <= 5e-3; the mode really runs other arithmetic than tf32 (results differ)"""
    blob, cfg, state = load_golden(name)
    cfg["PRECISION"] = "bf16"
    feats, coords = torch.from_numpy(blob["voxel_features"]), torch.from_numpy(blob["voxel_coords"])
    model, sp = run_product(cfg, state, blob["grid"], blob["pc_range"], feats, coords, int(blob["batch_size"]))
    assert model.backbone[0].precision == "bf16" and model.backbone[0]._terms() == 0
    assert torch.equal(sp.indices.cpu(), torch.from_numpy(blob["out_indices"]))
    ref = torch.from_numpy(blob["out_features"])
    got = sp.features.cpu()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    rms = (got - ref).pow(2).mean().sqrt().item()
    assert err <= BF16_TOL * scale and rms <= BF16_RMS * scale, (err / scale, rms / scale)
    model.set_precision("tf32")
    with torch.no_grad():
        other = model({"voxel_features": feats.cuda(), "voxel_coords": coords.cuda().float(),
                       "batch_size": int(blob["batch_size"])})["encoded_spconv_tensor"].features.cpu()
    assert (other - ref).abs().max().item() < err      # TF32 is closer to the fp32 reference than bf16
    print("bf16 %s: max %.2e rms %.2e of max|ref|" % (name, err / scale, rms / scale))


def test_shape_outside_the_tensor_core_family_warns():
    """a tensor-core precision mode on a shape the tcgen05 kernels do not cover runs the FFMA kernels (same results)
    and says so -- once per kind of kernel -- instead of silently being five times slower"""
    import mssvt_b200.mssvt_backbone as mb
    blob, cfg, state = load_golden("mixed_b3_n500")     # (two-group compress block, capped lists, out_linear ...)
    cfg["PRECISION"] = "tf32x3"
    mb._FFMA_WARNED.clear()
    feats, coords = torch.from_numpy(blob["voxel_features"]), torch.from_numpy(blob["voxel_coords"])
    with pytest.warns(RuntimeWarning, match="outside the shape family of the tcgen05 kernels"):
        run_product(cfg, state, blob["grid"], blob["pc_range"], feats, coords, int(blob["batch_size"]))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error", RuntimeWarning)     # the second forward stays quiet
        run_product(cfg, state, blob["grid"], blob["pc_range"], feats, coords, int(blob["batch_size"]))


@pytest.mark.parametrize("block_heads,compress_heads", [((1, 1), 2), ((4, 4), 8), ((2, 2), 4)])
def test_other_head_counts_of_the_tensor_core_family(block_heads, compress_heads):
    """the tcgen05 kernels are templated on the heads per group (1 / 2 / 4 in the mixed-scale blocks, 2 / 4 / 8 in the
    compress block); every instantiation, in every tensor-core mode, against the live oracle (no FFMA fallback)"""
    import warnings
    from mssvt_b200.config import AttrDict, block_cfg, compress_cfg
    feats, coords = synth_frame(21, 4000, crop=0.1)
    feats, coords = torch.from_numpy(feats), torch.from_numpy(coords)
    blocks = [block_cfg(num_heads=block_heads, cbs_pattern=p) for p in (1, 0)] + [compress_cfg(num_heads=(compress_heads,))]
    cfg = AttrDict(NAME="MixedScaleSparseTransformer", HASH_SIZE=400000, NUM_OUTPUT_FEATURES=64, PARAMS=blocks)
    torch.manual_seed(3)
    model = MixedScaleSparseTransformer(cfg, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE))
    state = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        want = orc.backbone_forward(state, cfg, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE), feats, coords, 1)
    model = model.cuda().eval()
    scale = want.features.abs().max().item()
    for precision, tol in (("bf16x3", FEATURE_TOL), ("tf32x3", FEATURE_TOL), ("tf32", TF32_TOL), ("bf16", BF16_TOL)):
        model.set_precision(precision)
        with warnings.catch_warnings():
            warnings.simplefilter("error", RuntimeWarning)      # (a fallback to the FFMA kernels would warn)
            with torch.no_grad():
                sp = model({"voxel_features": feats.cuda(), "voxel_coords": coords.cuda().float(),
                            "batch_size": 1})["encoded_spconv_tensor"]
        assert torch.equal(sp.indices.cpu(), want.indices)
        err = (sp.features.cpu() - want.features).abs().max().item()
        assert err <= tol * scale, (precision, block_heads, compress_heads, err / scale)
