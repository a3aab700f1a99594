"""CPU oracle for the Python layer of the MsSVT backbone (eval-mode forward).

TEST INFRASTRUCTURE -- not product.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module.

Restates, in fp32 PyTorch-CPU on top of oracle/ops.py:
  * get_vox_query_table            pcdet/models/backbones_3d/mssvt_backbone.py:73-122
  * MixedScaleSparseTransformerBlock.forward          ...:201-346
  * MixedScaleSparseTransformerCompressBlock.forward  ...:349-398
  * MixedScaleSparseTransformer.forward               ...:450-472
  * SparseTensor.build_map_table / dense    pcdet/models/model_utils/mssvt_utils.py:33-62
  * MixedScaleAttention.forward             pcdet/models/model_utils/mssvt_utils.py:88-157
Weights come in as a state dict with the reference's parameter names
(backbone.{i}.ms_attn.to_qs.{g}.weight, ... SURVEY.md 5.4), so the same dict drives the
reference, this oracle and the product.

Pinned by oracle/pin_against_reference.py, which runs the reference's *unmodified* Python
(loaded from /root/reference) on the same inputs and weights and requires identical integer
tensors and features equal to ~1e-6; the vectors it writes live in tests/golden/.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import ops

MAX_NUM_WINS = 90000  # mssvt_backbone.py:56 (hard-coded there)


# ----------------------------------------------------------------------------- query tables

def vox_query_table(win1, win2=None):
    """Offset tables (mssvt_backbone.py:73-122).  All offsets of the larger window centred at
    size//2, ordered by Chebyshev distance (stable: ties keep x-major/z-fastest order), split
    into win1 / outside, and win1 by x,y parity (floor-mod: -1 % 2 == 1)."""
    big = win1 if win2 is None else win2
    if win2 is not None:
        assert all((win2[i] - win1[i]) % 2 == 0 for i in range(3))
    offs = np.array([(x, y, z) for x in range(big[0]) for y in range(big[1]) for z in range(big[2])],
                    dtype=np.int64) - np.array([s // 2 for s in big])
    order = np.argsort(np.abs(offs).max(1), kind="stable")
    offs = offs[order]
    as_t = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(np.int32))).reshape(-1, 3)
    if win2 is None:
        return {"win1": as_t(offs)}
    inside = np.ones(len(offs), dtype=bool)
    for ax in range(3):
        half, slack = win1[ax] // 2, 1 - win1[ax] % 2
        inside &= (offs[:, ax] <= half + slack) & (offs[:, ax] >= -half)
    near, far = offs[inside], offs[~inside]
    odd = (near[:, 0] % 2 == 1) & (near[:, 1] % 2 == 1)
    even = (near[:, 0] % 2 == 0) & (near[:, 1] % 2 == 0)
    return {"odd": as_t(near[odd]), "even": as_t(near[even]), "win1": as_t(near[~(odd | even)]),
            "win2": as_t(far)}


# ----------------------------------------------------------------------------- container

class Frame:
    """What the reference calls SparseTensor (mssvt_utils.py:21-62)."""

    def __init__(self, features, indices, spatial_shape, voxel_size, point_cloud_range,
                 batch_size, hash_size, map_table=None):
        self.features = features
        self.indices = indices.to(torch.int32).contiguous()
        self.spatial_shape = [int(v) for v in spatial_shape]
        self.voxel_size = list(voxel_size)
        self.point_cloud_range = list(point_cloud_range)
        self.batch_size = batch_size
        self.hash_size = hash_size
        if map_table is None:
            map_table = ops.build_hash_table(batch_size, hash_size, self.spatial_shape,
                                             self.indices, per_sample_count(self.indices, batch_size))
        self.map_table = map_table

    def dense(self, channels_first=True):
        shape = [self.batch_size] + self.spatial_shape[::-1] + [self.features.shape[1]]
        out = torch.zeros(shape, dtype=self.features.dtype)
        i = self.indices.long()
        out[i[:, 0], i[:, 1], i[:, 2], i[:, 3]] = self.features
        return out.permute(0, 4, 1, 2, 3).contiguous() if channels_first else out


def per_sample_count(indices, batch_size):
    """with_bs_cnt, mssvt_backbone.py:124-130."""
    return torch.bincount(indices[:, 0].long(), minlength=batch_size)[:batch_size].to(torch.int32)


def world_coords(indices, point_cloud_range, voxel_size):
    """with_coords, mssvt_backbone.py:132-137: three separate fp32 ops (add, mul, add)."""
    vs = torch.tensor(voxel_size).unsqueeze(0)
    lo = torch.tensor(point_cloud_range[0:3]).unsqueeze(0)
    return (indices[:, [3, 2, 1]].float() + 0.5) * vs + lo


# ----------------------------------------------------------------------------- attention

def mixed_scale_attention(P, num_heads, query, keys, query_mask=None, key_masks=None,
                          batch_first=False):
    """mssvt_utils.py:88-157.  Head group g reads channel slice g of the queries and key chunk g
    (tot_nk / groups keys) only; additive -100 mask, softmax only when a mask is given."""
    if not batch_first:
        query, keys = query.transpose(1, 0), keys.transpose(1, 0)
    b, nq, C = query.shape
    G = len(num_heads)
    hd = C // sum(num_heads)
    nk = keys.shape[1] // G
    outs, c0 = [], 0
    for g, h in enumerate(num_heads):
        c1 = c0 + hd * h
        q = F.linear(query[:, :, c0:c1], P["ms_attn.to_qs.%d.weight" % g], P["ms_attn.to_qs.%d.bias" % g])
        q = q.reshape(b, nq, h, hd).permute(0, 2, 1, 3)
        kv = F.linear(keys[:, g * nk:(g + 1) * nk, c0:c1], P["ms_attn.to_kvs.%d.weight" % g],
                      P["ms_attn.to_kvs.%d.bias" % g])
        kv = kv.reshape(b, nk, 2, h, hd).permute(2, 0, 3, 1, 4)
        k, v = kv[0], kv[1]
        c0 = c1
        attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
        if key_masks is not None:
            km = key_masks[:, g * nk:(g + 1) * nk]
            attn = attn + km.float().masked_fill(km != 0, -100.0).view(b, 1, 1, nk)
            attn = torch.softmax(attn, dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(b, nq, -1)
        outs.append(F.linear(x, P["ms_attn.projs.%d.weight" % g], P["ms_attn.projs.%d.bias" % g]))
    out = torch.cat(outs, dim=-1)
    if query_mask is not None:
        out = out * (~query_mask).unsqueeze(-1).float()
    return out if batch_first else out.transpose(1, 0)


def _pos_proj(P, x):
    """pos_proj (mssvt_backbone.py:43-54): Conv1d(6,C,1)+ReLU [+ Conv1d(C,C,1)+ReLU] on (W,6,n)."""
    y = F.relu(F.conv1d(x, P["pos_proj.0.weight"], P["pos_proj.0.bias"]))
    if "pos_proj.2.weight" in P:
        y = F.relu(F.conv1d(y, P["pos_proj.2.weight"], P["pos_proj.2.bias"]))
    return y


def _ffn(P, x):
    C = x.shape[1]
    y = F.layer_norm(x, (C,), P["norm2.weight"], P["norm2.bias"])
    y = F.linear(F.relu(F.linear(y, P["linear1.weight"], P["linear1.bias"])), P["linear2.weight"],
                 P["linear2.bias"])
    x = x + y
    if "out_linear.weight" in P:
        x = F.linear(x, P["out_linear.weight"], P["out_linear.bias"])
    return x


# ----------------------------------------------------------------------------- blocks

def block_geometry(cfg, sp):
    """Everything of Block.forward that depends on voxel coordinates only
    (mssvt_backbone.py:213-258, 264-269): windows, chessboard lists, FPS keys, masks."""
    win1, win2 = cfg["window_size"]
    K = cfg["key_num_sample"]
    grid = [sp.spatial_shape[i] // win1[i] for i in range(3)]
    win_ind, _ = ops.get_non_empty_window_center(win1, MAX_NUM_WINS, sp.batch_size, sp.hash_size,
                                                 grid, sp.indices)
    tab = vox_query_table(win1, win2)
    n1 = cfg.get("max_num_win1") or win1[0] * win1[1] * win1[2]
    n2 = cfg.get("max_num_win2") or win2[0] * win2[1] * win2[2]
    (i_odd, i_even, i_w1, i_w2, c_odd, c_even, c_w1, c_w2) = ops.gather_two_window_voxels(
        sp.spatial_shape, win1, tab["odd"].shape[0], tab["even"].shape[0], n1, n2,
        tab["odd"], tab["even"], tab["win1"], tab["win2"], win_ind, sp.map_table)
    pattern = cfg.get("cbs_pattern", 1)
    q_ind, q_off = {0: (i_even, c_even), 1: (i_odd, c_odd), 2: (i_w1, c_w1)}[pattern]
    g = {"win_ind": win_ind, "q_ind": q_ind, "q_mask": q_ind < 0, "q_off": q_off,
         "win1_ind": i_w1, "win1_off": c_w1, "win2_ind": i_w2, "win2_off": c_w2,
         "ind_odd": i_odd, "ind_even": i_even}
    for name, ind, off in (("win1", i_w1, c_w1), ("win2", i_w2, c_w2)):
        fps = ops.farthest_point_sample(off.float(), K)
        mask = fps == 0
        mask[:, 0] = False
        # Q1: (-1 + 0.1).int() == 0, padded picks alias voxel 0 of the sample
        k_ind = (ops.gather_operation(ind.unsqueeze(1).float(), fps).squeeze(1) + 0.1).int()
        g["fps_" + name], g["k_ind_" + name], g["k_mask_" + name] = fps, k_ind, mask | (k_ind < 0)
    g["v_cnt"] = per_sample_count(sp.indices, sp.batch_size)
    g["w_cnt"] = per_sample_count(win_ind, sp.batch_size)
    return g


def block_forward(P, cfg, sp, taps=None):
    """MixedScaleSparseTransformerBlock.forward in eval mode (mssvt_backbone.py:201-346)."""
    C = sp.features.shape[1]
    win1 = cfg["window_size"][0]
    x = sp.features
    xn = F.layer_norm(x, (C,), P["norm1.weight"], P["norm1.bias"])
    g = block_geometry(cfg, sp)
    v_cnt, w_cnt = g["v_cnt"], g["w_cnt"]
    group = lambda feats, idx: ops.grouping_operation(feats, v_cnt, idx, w_cnt)

    q_fea = group(xn, g["q_ind"])
    k_fea1, k_fea2 = group(xn, g["k_ind_win1"]), group(xn, g["k_ind_win2"])
    xyz = world_coords(sp.indices, sp.point_cloud_range, sp.voxel_size)
    q_xyz, w1_xyz = group(xyz, g["q_ind"]), group(xyz, g["win1_ind"])
    k_xyz1, k_xyz2 = group(xyz, g["k_ind_win1"]), group(xyz, g["k_ind_win2"])
    win_size = [sp.voxel_size[i] * win1[i] for i in range(3)]
    centre = world_coords(g["win_ind"], sp.point_cloud_range, win_size).unsqueeze(-1)  # (W,3,1)

    k_rel1 = (k_xyz1 - centre) * (~g["k_mask_win1"]).unsqueeze(1)
    k_rel2 = (k_xyz2 - centre) * (~g["k_mask_win2"]).unsqueeze(1)
    q_rel = (q_xyz - centre) * (~g["q_mask"]).unsqueeze(1)
    q_fea = q_fea + _pos_proj(P, torch.cat((q_rel, centre.expand_as(q_rel)), 1))
    k_rel = torch.cat([k_rel1, k_rel2], -1)
    k_fea = torch.cat([k_fea1, k_fea2], -1) + _pos_proj(P, torch.cat((k_rel, centre.expand_as(k_rel)), 1))
    k_mask = torch.cat([g["k_mask_win1"], g["k_mask_win2"]], -1)

    attn = mixed_scale_attention(P, cfg["num_heads"], q_fea.permute(0, 2, 1).contiguous(),
                                 k_fea.permute(0, 2, 1).contiguous(), query_mask=g["q_mask"],
                                 key_masks=k_mask, batch_first=True)  # (W, nq, C)

    interp = cfg.get("use_feature_interpolation", True)
    if interp:
        # Q4: 3-NN of every win1 slot among the window's query slots (padded ones sit at 0,0,0)
        dist, nn_idx = ops.three_nn(w1_xyz.permute(0, 2, 1).contiguous(),
                                    q_xyz.permute(0, 2, 1).contiguous())
        wgt = 1.0 / torch.clamp(dist, min=1e-10)
        wgt = wgt / wgt.sum(-1, keepdim=True)
        picked = ops.group_points(attn.permute(0, 2, 1).contiguous(), nn_idx)  # (W,C,n1,3)
        rows = (picked * wgt.unsqueeze(1)).sum(-1).permute(0, 2, 1).reshape(-1, C)
        tgt = g["win1_ind"]
    else:
        rows, tgt = attn.reshape(-1, C), g["q_ind"]
        nn_idx = wgt = None

    # Q5: scatter into a clone of the pre-norm features, -1 goes to a throw-away row
    merged = x.clone()
    v0 = w0 = 0
    per_win = tgt.shape[1]
    for nv, nw in zip(v_cnt.tolist(), w_cnt.tolist()):
        buf = torch.cat([x[v0:v0 + nv], torch.zeros(1, C)], 0)
        buf[tgt[w0:w0 + nw].reshape(-1).long()] = rows[w0 * per_win:(w0 + nw) * per_win]
        merged[v0:v0 + nv] = buf[:-1]
        v0, w0 = v0 + nv, w0 + nw
    y = _ffn_block(P, merged + x)
    if taps is not None:
        taps.append({**g, "attn": attn, "nn_idx": nn_idx, "nn_weight": wgt, "merged": merged})
    sp.features = y
    return sp


def _ffn_block(P, x):
    return _ffn(P, x)


def compress_forward(P, cfg, sp, taps=None):
    """MixedScaleSparseTransformerCompressBlock.forward, eval (mssvt_backbone.py:349-398)."""
    C = sp.features.shape[1]
    win1 = cfg["window_size"][0]
    xn = F.layer_norm(sp.features, (C,), P["norm1.weight"], P["norm1.bias"])
    grid = [sp.spatial_shape[i] // win1[i] for i in range(3)]
    win_ind, win_table = ops.get_non_empty_window_center(win1, MAX_NUM_WINS, sp.batch_size,
                                                         sp.hash_size, grid, sp.indices)
    tab = vox_query_table(win1)
    n1 = cfg.get("max_num_win1") or win1[0] * win1[1] * win1[2]
    k_ind, k_off = ops.gather_one_window_voxels(sp.spatial_shape, win1, n1, tab["win1"], win_ind,
                                                sp.map_table)
    k_mask = k_ind < 0
    v_cnt = per_sample_count(sp.indices, sp.batch_size)
    w_cnt = per_sample_count(win_ind, sp.batch_size)
    k_fea = ops.grouping_operation(xn, v_cnt, k_ind, w_cnt)  # (W,C,n1), zeros at padding
    xyz = world_coords(sp.indices, sp.point_cloud_range, sp.voxel_size)
    k_xyz = ops.grouping_operation(xyz, v_cnt, k_ind, w_cnt)
    win_size = [sp.voxel_size[i] * win1[i] for i in range(3)]
    centre = world_coords(win_ind, sp.point_cloud_range, win_size).unsqueeze(-1)
    q_fea = k_fea.max(dim=-1)[0].unsqueeze(0)  # Q6: max includes the zero padding
    k_rel = k_xyz - centre
    k_fea = k_fea + _pos_proj(P, torch.cat((k_rel, centre.expand_as(k_rel)), 1))
    attn = mixed_scale_attention(P, cfg["num_heads"], q_fea, k_fea.permute(2, 0, 1).contiguous(),
                                 key_masks=k_mask)  # (1, W, C)
    y = _ffn(P, attn.squeeze(0))
    if taps is not None:
        taps.append({"win_ind": win_ind, "k_ind": k_ind, "k_off": k_off, "attn": attn.squeeze(0)})
    return Frame(y, win_ind, grid, win_size, sp.point_cloud_range, sp.batch_size, sp.hash_size,
                 map_table=win_table)


def split_state_dict(state, num_blocks):
    """{'backbone.3.norm1.weight': t} -> per-block dicts keyed 'norm1.weight'."""
    per = [dict() for _ in range(num_blocks)]
    for k, v in state.items():
        parts = k.split(".")
        if parts[0] == "backbone":
            per[int(parts[1])][".".join(parts[2:])] = v.detach().float().cpu()
    return per


def backbone_forward(state, model_cfg, grid_size, voxel_size, point_cloud_range, voxel_features,
                     voxel_coords, batch_size, taps=None):
    """MixedScaleSparseTransformer.forward (mssvt_backbone.py:450-472) -> Frame."""
    params = model_cfg["PARAMS"]
    per = split_state_dict(state, len(params))
    sp = Frame(voxel_features.float().cpu(), voxel_coords.int().cpu(), grid_size, voxel_size,
               point_cloud_range, batch_size, model_cfg["HASH_SIZE"])
    for P, cfg in zip(per, params):
        if cfg["name"] == "MixedScaleSparseTransformerBlock":
            sp = block_forward(P, cfg, sp, taps)
        elif cfg["name"] == "MixedScaleSparseTransformerCompressBlock":
            sp = compress_forward(P, cfg, sp, taps)
        else:
            raise NotImplementedError(cfg["name"])
    return sp
