"""Pin the CPU oracle against the reference's own, unmodified Python layer and write the golden
vectors under tests/golden/.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference); the GPU box
never executes it -- it uses the committed .npz files.

How: the reference's hot path is CUDA-only Python (mssvt_backbone.py, mssvt_utils.py,
mssvt_ops.py, pointnet2_utils.py).  We load those four files *as they are* from
/root/reference through the normal import machinery, with
  * stub packages in sys.modules so that none of the pcdet __init__.py files (spconv, numba,
    SharedArray ...) run,
  * a one-class stand-in for timm.models.layers.DropPath (identity in eval),
  * `mssvt_ops_cuda` and `pointnet2_batch_cuda` replaced by modules with the pybind signatures
    (ops/mssvt/src/ms_api.cpp:7-14, pointnet2_batch/src/pointnet2_api.cpp:10-24) that run the C
    restatement in oracle/mssvt_oracle.c on the tensors the reference's Python allocated,
  * `.cuda()` / torch.cuda.FloatTensor / device='cuda' mapped to the CPU.
So every line of the reference's Python (allocation conventions, Q1 index aliasing, masks,
pos-emb, attention, three-NN weights, merge loop, compress block) executes for real, and its
results are compared with oracle/backbone.py.  The C kernels themselves are pinned separately
on the GPU box against the reference's compiled CUDA (tests/test_gpu_ref_kernels.py).

Usage:  python -m oracle.pin_against_reference [--write]
"""
import argparse
import importlib
import os
import sys
import types

import numpy as np
import torch
from torch import nn

from . import backbone as orc
from . import ops as orc_ops

REF = os.environ.get("MSSVT_REFERENCE", "/root/reference")
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


# ----------------------------------------------------------------------------- CPU shims

def _install_cpu_shims():
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.FloatTensor = lambda *shape: torch.empty(*shape, dtype=torch.float32)
    torch.cuda.IntTensor = lambda *shape: torch.empty(*shape, dtype=torch.int32)
    real_tensor = torch.tensor

    def tensor(*a, **k):
        k.pop("device", None)
        return real_tensor(*a, **k)

    torch.tensor = tensor


def _stub_package(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    parent, _, leaf = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], leaf, m)
    return m


class _Recorder(list):
    def add(self, op, **tensors):
        self.append((op, {k: v.clone() for k, v in tensors.items()}))


RECORD = _Recorder()


def _fake_mssvt_ops_cuda():
    m = types.ModuleType("pcdet.ops.mssvt.mssvt_ops_cuda")
    L, p = orc_ops.lib(), orc_ops._p

    def build_mapping_with_hash_wrapper(x, y, z, n, h, v_indices, v_bs_cnt, table):
        L.orc_build_hash_table(x, y, z, n, h, p(v_indices), p(v_bs_cnt), p(table))
        RECORD.add("hash", table=table)
        return 1

    def window_with_hash_wrapper(xg, yg, zg, xw, yw, zw, n, max_wins, h, v_indices, w_indices,
                                 table, vcount):
        over = L.orc_window_partition(xg, yg, zg, xw, yw, zw, n, max_wins, h, p(v_indices),
                                      p(w_indices), p(table), p(vcount))
        assert over == 0
        RECORD.add("window", w_indices=w_indices, table=table, vcount=vcount)
        return 1

    def gather_two_window_voxels_with_hash_wrapper(x, y, z, xw, yw, zw, mo, me, m1, m2, W, h, no,
                                                   ne, n1, n2, io, ie, i1, i2, co, ce, c1, c2,
                                                   qo, qe, q1, q2, win, table):
        L.orc_gather_two_window(x, y, z, xw, yw, zw, mo, me, m1, m2, W, h, no, ne, n1, n2, p(io),
                                p(ie), p(i1), p(i2), p(co), p(ce), p(c1), p(c2), p(qo), p(qe),
                                p(q1), p(q2), p(win), p(table))
        RECORD.add("gather2", ind_odd=io, ind_even=ie, ind_win1=i1, ind_win2=i2, coord_odd=co,
                   coord_even=ce, coord_win1=c1, coord_win2=c2)
        return 1

    def gather_one_window_voxels_with_hash_wrapper(x, y, z, xw, yw, zw, m1, W, h, n1, i1, c1, q1,
                                                   win, table):
        L.orc_gather_one_window(x, y, z, xw, yw, zw, m1, W, h, n1, p(i1), p(c1), p(q1), p(win),
                                p(table))
        RECORD.add("gather1", ind_win1=i1, coord_win1=c1)
        return 1

    def group_features_wrapper(B, M, C, ns, features, fbc, idx, ibc, out):
        L.orc_group_features(B, M, C, ns, p(features.contiguous()), p(fbc), p(idx), p(ibc), p(out))
        return 1

    def group_features_grad_wrapper(B, M, C, N, ns, grad_out, idx, ibc, fbc, grad):
        L.orc_group_features_grad(B, M, C, N, ns, p(grad_out), p(idx), p(ibc), p(fbc), p(grad))
        return 1

    for f in (build_mapping_with_hash_wrapper, window_with_hash_wrapper,
              gather_two_window_voxels_with_hash_wrapper, gather_one_window_voxels_with_hash_wrapper,
              group_features_wrapper, group_features_grad_wrapper):
        setattr(m, f.__name__, f)
    return m


def _fake_pointnet2_batch_cuda():
    m = types.ModuleType("pcdet.ops.pointnet2.pointnet2_batch.pointnet2_batch_cuda")
    L, p = orc_ops.lib(), orc_ops._p

    def farthest_point_sampling_wrapper(B, N, npoint, xyz, temp, out):
        L.orc_fps(B, N, npoint, p(xyz), p(temp), p(out))
        RECORD.add("fps", idx=out)
        return 1

    def gather_points_wrapper(B, C, N, npoint, features, idx, out):
        L.orc_gather_points(B, C, N, npoint, p(features), p(idx), p(out))
        return 1

    def three_nn_wrapper(B, N, m_, unknown, known, dist2, idx):
        L.orc_three_nn(B, N, m_, p(unknown), p(known), p(dist2), p(idx))
        RECORD.add("three_nn", idx=idx, dist2=dist2)
        return 1

    def group_points_wrapper(B, C, N, npnt, ns, features, idx, out):
        L.orc_group_points(B, C, N, npnt, ns, p(features), p(idx), p(out))
        return 1

    for f in (farthest_point_sampling_wrapper, gather_points_wrapper, three_nn_wrapper,
              group_points_wrapper):
        setattr(m, f.__name__, f)
    return m


def load_reference():
    """-> (mssvt_backbone module, mssvt_utils module) of the reference, importable on CPU."""
    if "pcdet.models.backbones_3d.mssvt_backbone" in sys.modules:
        return (sys.modules["pcdet.models.backbones_3d.mssvt_backbone"],
                sys.modules["pcdet.models.model_utils.mssvt_utils"])
    _install_cpu_shims()
    pc = os.path.join(REF, "pcdet")
    for name, rel in (("pcdet", ""), ("pcdet.models", "models"),
                      ("pcdet.models.backbones_3d", "models/backbones_3d"),
                      ("pcdet.models.model_utils", "models/model_utils"), ("pcdet.ops", "ops"),
                      ("pcdet.ops.mssvt", "ops/mssvt"), ("pcdet.ops.pointnet2", "ops/pointnet2"),
                      ("pcdet.ops.pointnet2.pointnet2_batch", "ops/pointnet2/pointnet2_batch")):
        _stub_package(name, os.path.join(pc, rel))
    for name in ("timm", "timm.models", "timm.models.layers"):
        _stub_package(name, "/nonexistent")

    class DropPath(nn.Module):  # timm's DropPath is the identity in eval mode
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            assert not self.training
            return x

    sys.modules["timm.models.layers"].DropPath = DropPath
    for fake in (_fake_mssvt_ops_cuda(), _fake_pointnet2_batch_cuda()):
        sys.modules[fake.__name__] = fake
        parent, _, leaf = fake.__name__.rpartition(".")
        setattr(sys.modules[parent], leaf, fake)
    bb = importlib.import_module("pcdet.models.backbones_3d.mssvt_backbone")
    ut = importlib.import_module("pcdet.models.model_utils.mssvt_utils")
    return bb, ut


# ----------------------------------------------------------------------------- cases

def _cases():
    from mssvt_b200.config import block_cfg, compress_cfg, s0_model_cfg, AttrDict
    small = dict(grid=(48, 48, 32), pc_range=(-7.68, -7.68, -2.0, 7.68, 7.68, 4.0))
    yield "s0_b2_n1200", s0_model_cfg(hash_size=4001, cbs_patterns=(1, 0, 2)), 1200, 2, small, 0
    # odd z extent (remainder strip dropped by the window grid), capped lists, no interpolation,
    # channel change through out_linear, tiny hash table with heavy probing
    cfg = AttrDict(NAME="MixedScaleSparseTransformer", HASH_SIZE=1531, NUM_OUTPUT_FEATURES=32,
                   PARAMS=[block_cfg(channels=(32, 64, 32), num_heads=(1, 1), cbs_pattern=1,
                                     key_num_sample=16, max_num_win1=20, max_num_win2=60),
                           block_cfg(channels=(32, 64, 48), num_heads=(2, 2), cbs_pattern=0,
                                     window_size=((3, 3, 5), (7, 7, 9)), key_num_sample=32,
                                     use_feature_interpolation=False),
                           compress_cfg(channels=(48, 96, 32), num_heads=(2, 1),
                                        window_size=((2, 2, 4),), max_num_win1=None)])
    yield "mixed_b3_n500", cfg, 500, 3, small, 11


def run_case(name, model_cfg, n_per_sample, batch, geom, seed, write=False):
    from mssvt_b200.synth import synth_frame, S0_VOXEL
    bb, _ = load_reference()
    feats, coords = synth_frame(seed, n_per_sample, batch_size=batch,
                                channels=model_cfg.PARAMS[0].channels[0], grid=geom["grid"],
                                pc_range=geom["pc_range"], crop=1.0)
    feats, coords = torch.from_numpy(feats), torch.from_numpy(coords)
    torch.manual_seed(0)
    ref = bb.MixedScaleSparseTransformer(model_cfg, feats.shape[1], list(geom["grid"]),
                                         list(S0_VOXEL), list(geom["pc_range"])).eval()
    # make LayerNorm affine / biases non-trivial so that parity exercises them
    with torch.no_grad():
        for prm in ref.parameters():
            if prm.dim() == 1:
                prm.add_(0.1 * torch.randn_like(prm))
    state = {k: v.detach().clone() for k, v in ref.state_dict().items()}

    # Offset tables: the reference orders ties of its Chebyshev sort with an unstable torch.sort
    # (mssvt_backbone.py:85), so the order within one distance shell is implementation-defined
    # (CPU and CUDA differ).  The tables are *inputs* of the ops; we check the reference's
    # tables hold the same offsets with non-decreasing distance, then give both sides the
    # stable-order tables the product uses.
    for blk, cfg in zip(ref.backbone, model_cfg.PARAMS):
        ws = cfg.window_size
        mine = orc.vox_query_table(ws[0], ws[1] if len(ws) == 2 else None)
        for k, t in blk.vox_query_table.items():
            t = t.cpu().int()
            assert sorted(map(tuple, t.tolist())) == sorted(map(tuple, mine[k].tolist())), (name, k)
            cheb = t.abs().max(1)[0]
            assert bool((cheb[1:] >= cheb[:-1]).all()), (name, k)
            blk.vox_query_table[k] = mine[k].clone()

    per_block = []
    hooks = [blk.register_forward_hook(lambda m, i, o: per_block.append(
        (o.features.detach().clone(), o.indices.clone()))) for blk in ref.backbone]
    del RECORD[:]
    with torch.no_grad():
        out = ref({"voxel_features": feats.clone(), "voxel_coords": coords.float(),
                   "batch_size": batch})["encoded_spconv_tensor"]
    for h in hooks:
        h.remove()
    ref_record = list(RECORD)

    taps = []
    with torch.no_grad():
        got = orc.backbone_forward(state, model_cfg, list(geom["grid"]), list(S0_VOXEL),
                                   list(geom["pc_range"]), feats, coords, batch, taps=taps)
    assert torch.equal(got.indices, out.indices.int()), "final indices differ"
    err = (got.features - out.features).abs().max().item()
    scale = out.features.abs().max().item()
    assert err <= 2e-5 * max(scale, 1.0), ("final features differ", err, scale)
    assert torch.equal(got.dense(), out.dense()) or (got.dense() - out.dense()).abs().max() < 1e-4

    # integer stages, in call order, against the oracle's own taps
    it = iter(ref_record)
    first = next(it)
    assert first[0] == "hash"
    bi = 0
    for cfg, tap in zip(model_cfg.PARAMS, taps):
        op, rec = next(it)
        assert op == "window"
        if cfg.name.endswith("CompressBlock"):
            op, rec = next(it)
            assert op == "gather1" and torch.equal(rec["ind_win1"], tap["k_ind"])
            assert torch.equal(rec["coord_win1"], tap["k_off"])
        else:
            op, rec = next(it)
            assert op == "gather2"
            assert torch.equal(rec["ind_win1"], tap["win1_ind"]) and torch.equal(rec["ind_win2"], tap["win2_ind"])
            assert torch.equal(rec["ind_odd"], tap["ind_odd"]) and torch.equal(rec["ind_even"], tap["ind_even"])
            op, rec = next(it)
            assert op == "fps" and torch.equal(rec["idx"], tap["fps_win1"])
            op, rec = next(it)
            assert op == "fps" and torch.equal(rec["idx"], tap["fps_win2"])
            if cfg.use_feature_interpolation:
                op, rec = next(it)
                assert op == "three_nn" and torch.equal(rec["idx"], tap["nn_idx"])
        bi += 1
    for (f_ref, i_ref), blk_i in zip(per_block, range(len(per_block))):
        pass  # per-block features are stored in the golden file below

    print("[pin] %-16s N=%d W(block0)=%d  max|d|=%.2e (scale %.2f)  -> oracle == reference"
          % (name, feats.shape[0], taps[0]["win_ind"].shape[0], err, scale))

    if write:
        os.makedirs(GOLDEN, exist_ok=True)
        blob = {"voxel_features": feats.numpy(), "voxel_coords": coords.numpy(),
                "out_features": out.features.numpy(), "out_indices": out.indices.int().numpy(),
                "grid": np.array(geom["grid"]), "pc_range": np.array(geom["pc_range"]),
                "batch_size": np.array(batch)}
        for i, (f, idx) in enumerate(per_block):
            blob["block%d_features" % i] = f.numpy()
            blob["block%d_indices" % i] = idx.int().numpy()
        for k, v in state.items():
            blob["state/" + k] = v.numpy()
        t0 = taps[0]
        for k in ("win_ind", "q_ind", "win1_ind", "win2_ind", "win1_off", "win2_off", "fps_win1",
                  "fps_win2", "k_ind_win1", "k_ind_win2", "k_mask_win1", "k_mask_win2", "nn_idx"):
            if t0.get(k) is not None:
                blob["tap0/" + k] = t0[k].numpy()
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **blob)
        import json
        with open(os.path.join(GOLDEN, name + ".cfg.json"), "w") as f:
            json.dump(model_cfg, f, indent=1)
    return err


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--write", action="store_true", help="(re)write tests/golden/*.npz")
    args = ap.parse_args()
    if not os.path.isdir(REF):
        print("reference tree not present at %s: nothing to pin against" % REF)
        return 1
    for case in _cases():
        run_case(*case, write=args.write)
    return 0


if __name__ == "__main__":
    sys.exit(main())
