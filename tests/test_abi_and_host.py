"""CPU: the C-ABI library exports every symbol the header declares (no compute calls), the
Python descriptor mirrors match the CUDA structs, and the host-side logic (offset tables, config,
synthetic frames, module construction, state-dict names) behaves like the reference's."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from mssvt_b200 import _lib
from mssvt_b200.config import s0_model_cfg
from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer, vox_query_table
from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL, synth_frame


def _declared():
    text = open(os.path.join(ROOT, "include", "mssvt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mssvt_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run `make -C mssvt_b200/csrc` (or __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "header declares %s but the library does not export it" % n
    assert set(_lib.EXPORTS) == set(names), set(_lib.EXPORTS) ^ set(names)


def test_descriptor_mirrors_match_cuda_structs():
    lib = _lib.load()  # loading needs no GPU
    assert lib.mssvt_sizeof_attn_shape() == ctypes.sizeof(_lib.AttnShape)
    assert lib.mssvt_sizeof_ffn_shape() == ctypes.sizeof(_lib.FfnShape)
    assert lib.mssvt_version().startswith(b"mssvt_b200")
    # the reference's host-side FPS block size (cuda_utils.h:10-14)
    for n, want in ((1, 0), (2, 1), (12, 3), (27, 4), (32, 5), (125, 6), (1024, 10), (5000, 10)):
        assert lib.mssvt_fps_log2_block(n) == want


def test_cpu_tensors_are_refused():
    from mssvt_b200 import mssvt_ops
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        mssvt_ops.build_hash_table(1, 11, [4, 4, 4], torch.zeros((1, 4), dtype=torch.int32),
                                   torch.ones(1, dtype=torch.int32))


def test_query_tables_match_oracle_and_probe_counts():
    from oracle.backbone import vox_query_table as orc_table
    for w1, w2 in (([3, 3, 3], [5, 5, 5]), ([3, 3, 5], [7, 7, 9]), ([2, 2, 4], None), ([1, 1, 32], None)):
        mine, ref = vox_query_table(w1, w2), orc_table(w1, w2)
        assert mine.keys() == ref.keys()
        for k in mine:
            assert np.array_equal(mine[k], ref[k].numpy()), (w1, w2, k)
    t = vox_query_table([3, 3, 3], [5, 5, 5])
    assert [len(t[k]) for k in ("odd", "even", "win1", "win2")] == [12, 3, 12, 98]
    t = vox_query_table([3, 3, 5], [7, 7, 9])
    assert [len(t[k]) for k in ("odd", "even", "win1", "win2")] == [20, 5, 20, 396]


def test_module_state_dict_names_load_reference_checkpoint():
    blob, cfg, state = load_golden("s0_b2_n1200")
    model = MixedScaleSparseTransformer(cfg, 64, blob["grid"].tolist(), list(S0_VOXEL), blob["pc_range"].tolist())
    assert set(model.state_dict().keys()) == set(state.keys())
    model.load_state_dict(state, strict=True)
    assert model.num_point_features == 64
    blk = model.backbone[0]
    assert blk.max_num_odd == 12 and blk.max_num_even == 3 and blk.max_num_win1 == 27 and blk.max_num_win2 == 125
    # construction rules of the reference: the last PARAMS entry cannot be a Block
    bad = s0_model_cfg()
    bad.PARAMS = bad.PARAMS[:3]
    with pytest.raises(IndexError):
        MixedScaleSparseTransformer(bad, 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE))


def test_fused_path_refuses_training_mode_and_cpu():
    model = MixedScaleSparseTransformer(s0_model_cfg(hash_size=101), 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        model({"voxel_features": torch.zeros(4, 64), "voxel_coords": torch.zeros(4, 4), "batch_size": 1})


def test_synthetic_frames_are_deterministic_unique_and_sorted():
    f1, c1 = synth_frame(3, 5000, batch_size=2, crop=0.3)
    f2, c2 = synth_frame(3, 5000, batch_size=2, crop=0.3)
    assert np.array_equal(c1, c2) and np.array_equal(f1, f2)
    assert c1.shape == (10000, 4) and c1.dtype == np.int32 and f1.shape == (10000, 64)
    key = ((c1[:, 0].astype(np.int64) * 468 + c1[:, 3]) * 468 + c1[:, 2]) * 32 + c1[:, 1]
    assert (np.diff(key) > 0).all()  # unique, ordered by (b, x, y, z), samples contiguous
    assert c1[:, 1].max() < 32 and c1[:, 2].max() < 468 and c1[:, 3].max() < 468 and c1.min() >= 0


def test_precision_modes_and_module_mirrors_on_cpu():
    """host logic that needs no GPU: precision switch, state-dict names of the producer / consumer mirrors"""
    import torch
    from mssvt_b200.config import AttrDict, s0_model_cfg
    from mssvt_b200.dynamic_vfe import DynamicVFE
    from mssvt_b200.height_compression import HeightCompression
    from mssvt_b200.mssvt_backbone import MixedScaleSparseTransformer
    from mssvt_b200.synth import S0_GRID, S0_RANGE, S0_VOXEL
    model = MixedScaleSparseTransformer(s0_model_cfg(), 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE))
    assert model.precision == "bf16x3" and all(b.precision == "bf16x3" for b in model.backbone)
    for mode in ("fp32", "tf32", "bf16", "bf16x3", "tf32x3"):
        assert model.set_precision(mode).backbone[-1].precision == mode
    with pytest.raises(ValueError):
        model.set_precision("fp16")
    assert model.set_precision("tf32x3").backbone[0]._terms() == 3 and model.set_precision("tf32").backbone[0]._terms() == 1
    # bf16: kind::f16 operands for the blocks and every FFN; the compress attention kernels stay TF32
    assert model.set_precision("bf16").backbone[0]._terms() == 0 and model.backbone[-1]._attn_terms() == 1
    assert model.set_precision("bf16x3").backbone[0]._terms() == 2 and model.backbone[-1]._attn_terms() == 2
    with pytest.raises(RuntimeError, match="CUDA tensors only"):   # no CPU path, also for the graph capture
        model.eval()({"voxel_features": torch.zeros(1, 64), "voxel_coords": torch.zeros(1, 4), "batch_size": 1})
    vfe = DynamicVFE(AttrDict(NUM_FILTERS=[32, 64]), 5, list(S0_VOXEL), list(S0_GRID), list(S0_RANGE))
    keys = set(vfe.state_dict())
    assert {"pfn.0.0.weight", "pfn.0.1.running_var", "pfn.1.0.bias"} <= keys   # dynamic_vfe.py:57-66 names
    assert vfe.pfn[0][0].in_features == 11 and vfe.pfn[1][0].in_features == 64 and vfe.get_output_feature_dim() == 64
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        vfe.eval()({"points": torch.zeros(3, 6), "batch_size": 1})
    bev = HeightCompression(AttrDict(NUM_BEV_FEATURES=64))
    assert [k for k in bev.state_dict() if k.endswith("weight")][0] == "compress_layers.0.weight"


def test_drop_path_stochastic_branch():
    """timm DropPath (mssvt_backbone.py:4, 42): per-row Bernoulli(keep) mask scaled by 1/keep in
    training mode, identity in eval mode and at drop_prob 0"""
    from mssvt_b200.mssvt_backbone import DropPath
    torch.manual_seed(0)
    x = torch.randn(20000, 8) + 3.0
    dp = DropPath(0.3)
    dp.eval()
    assert dp(x) is x
    dp.train()
    y = dp(x)
    dropped = (y == 0).all(1)
    kept = ~dropped
    assert abs(dropped.float().mean().item() - 0.3) < 0.02          # whole rows are dropped
    assert torch.allclose(y[kept], x[kept] / 0.7)                   # survivors are rescaled
    assert abs(y.mean().item() - x.mean().item()) < 0.05            # expectation is preserved
    assert DropPath(0.0).train()(x) is x
    # the backbone assigns linspace(0, 0.3, n_blocks - 1) by block position, nothing to the compress block
    model = MixedScaleSparseTransformer(s0_model_cfg(), 64, list(S0_GRID), list(S0_VOXEL), list(S0_RANGE))
    rates = [getattr(b.drop_path, "drop_prob", 0.0) for b in model.backbone]
    assert rates[0] == 0.0 and abs(rates[1] - 0.15) < 1e-6 and abs(rates[2] - 0.3) < 1e-6 and rates[3] == 0.0


def test_slab_plan_aligns_to_every_window_grid_and_rejects_narrow_slabs():
    """one-frame sharding: borders on the lcm of the blocks' window extents; a slab narrower than the halo would
    make neighbours disagree on the exchanged row sets, so it is refused up front (same error on every rank)"""
    from mssvt_b200.sharding import SlabPlan
    rng = np.random.default_rng(0)
    x = np.sort(rng.integers(0, 120, 4000)).astype(np.int32)
    coords = torch.from_numpy(np.stack([np.zeros_like(x), x % 7, x % 11, x], 1))
    for sorted_flag in (False, True):
        plans = [SlabPlan(coords, 6, 1, r, 3, grid_x=120, sorted_single_sample=sorted_flag) for r in range(3)]
        assert all(p.bounds == plans[0].bounds for p in plans)
        assert all(b % 6 == 0 for b in plans[0].bounds[:-1])
        for a, b in zip(plans, plans[1:]):                      # neighbours agree on the exchanged row counts
            assert a.send_right.numel() == b.recv_left.numel() and b.send_left.numel() == a.recv_right.numel()
    narrow = torch.from_numpy(np.stack([np.zeros(64, np.int32)] * 3 + [np.arange(64, dtype=np.int32) % 2], 1))
    with pytest.raises(ValueError, match="narrower than the halo"):
        SlabPlan(narrow, 2, 3, 0, 4, grid_x=2)


def test_training_ops_refuse_cpu_tensors():
    """the autograd.Functions of the training path (train_ops.py) have no CPU path either"""
    from mssvt_b200.train_ops import (WindowLists, embed_rows, interp_merge, layer_norm_rows, linear_rows,
                                      ragged_window_attention, segment_max)
    t = lambda *a: torch.tensor(a, dtype=torch.int32)
    lists = WindowLists(t(0, 1), t(0), t(0, 1), t(0), t(0))
    z = torch.zeros(1, dtype=torch.long)
    calls = [lambda: ragged_window_attention(torch.randn(1, 32), torch.randn(1, 64), lists, 2, 0.25),
             lambda: interp_merge(torch.randn(2, 64), torch.randn(3, 64), torch.zeros(3, 3, dtype=torch.int32), torch.rand(3, 3)),
             lambda: layer_norm_rows(torch.nn.LayerNorm(64), torch.randn(5, 64)),
             lambda: linear_rows(torch.nn.Linear(64, 128), torch.randn(5, 64)),
             lambda: segment_max(torch.randn(1, 64), lists, 1),
             lambda: embed_rows(torch.randn(4, 64), torch.randn(64, 6), torch.randn(64), torch.randn(4, 3), torch.randn(1, 3),
                                [(z, z, None, 0, 64)])]
    for fn in calls:
        with pytest.raises(RuntimeError, match="CUDA tensors only"):
            fn()
