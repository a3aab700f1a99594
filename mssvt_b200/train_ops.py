"""autograd.Functions of the training path over the hand-written forward / backward kernels
of csrc/train.cu: the window attention on the compact (ragged) form and the three-NN blend back to voxels.

What they replace is torch autograd over the reference's padded tensors (mssvt_utils.py:100-157 inside
mssvt_backbone.py:260-336).  No CPU path: CPU tensors raise in `ptr`.
"""
import ctypes

import torch
from torch.amp import custom_bwd, custom_fwd
from torch.autograd import Function

from ._lib import call, ptr, stream


class WindowLists:
    """CSR description of the windows of one head group: which rows of q / kv belong to which window.
    q_off / key_off (W + 1) int32 offsets, q_win / k_win window of every row, key_mult (W) multiplicity of the
    masked key (the LAST key row of a window; 0 = the window has none)."""

    def __init__(self, q_off, q_win, key_off, k_win, key_mult):
        self.q_off, self.q_win, self.key_off, self.k_win, self.key_mult = (
            t.to(torch.int32).contiguous() for t in (q_off, q_win, key_off, k_win, key_mult))
        self.num_queries, self.num_keys = int(self.q_win.shape[0]), int(self.k_win.shape[0])


class RaggedWindowAttention(Function):
    """out[q] = softmax_k(scale * q . k + mask) v over the keys of q's window, per head (mssvt_utils.py:123-139)
    q (Q, heads * hd), kv (R, 2 * heads * hd) = [K | V] (the layout nn.Linear(sd, 2 sd) of `to_kvs` produces)."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)     # (fp32 kernels: bf16 autocast stops at their door)
    def forward(ctx, q, kv, lists, heads, scale):
        q, kv = q.float().contiguous(), kv.float().contiguous()
        D = q.shape[1]
        assert kv.shape[1] == 2 * D and D % heads == 0
        assert q.shape[0] == lists.num_queries and kv.shape[0] == lists.num_keys
        hd = D // heads
        out = torch.empty_like(q)
        lse = torch.empty((q.shape[0], heads), dtype=torch.float32, device=q.device)
        v = kv[:, D:]
        call("mssvt_ragged_attention_fwd", heads, hd, float(scale), q.shape[0], ptr(lists.q_win), ptr(lists.key_off),
             ptr(lists.key_mult), ptr(q), D, ptr(kv), 2 * D, _ptr_view(v), 2 * D, ptr(out), D, ptr(lse), stream())
        ctx.save_for_backward(q, kv, out, lse)
        ctx.lists, ctx.heads, ctx.scale = lists, heads, float(scale)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        q, kv, out, lse = ctx.saved_tensors
        lists, heads, D = ctx.lists, ctx.heads, q.shape[1]
        grad_out = grad_out.float().contiguous()
        gq, gkv = torch.empty_like(q), torch.empty_like(kv)
        delta = torch.empty_like(lse)
        call("mssvt_ragged_attention_bwd", heads, D // heads, ctx.scale, q.shape[0], kv.shape[0], ptr(lists.q_win),
             ptr(lists.k_win), ptr(lists.q_off), ptr(lists.key_off), ptr(lists.key_mult), ptr(q), D, ptr(kv), 2 * D,
             _ptr_view(kv[:, D:]), 2 * D, ptr(out), D, ptr(lse), ptr(grad_out), D, ptr(delta), ptr(gq), D,
             ptr(gkv), 2 * D, _ptr_view(gkv[:, D:]), 2 * D, stream())
        return gq, gkv, None, None, None


def _ptr_view(t):
    """device pointer of a column slice of a contiguous matrix (the kernels take the row stride)"""
    if not t.is_cuda:
        raise RuntimeError("mssvt_b200 operators run on CUDA tensors only; there is no CPU path")
    return ctypes.c_void_p(t.data_ptr())


ragged_window_attention = RaggedWindowAttention.apply


class InterpMerge(Function):
    """merged[v] = sum_j w[v, j] rows[src[v, j]] (three_interpolate of mssvt_backbone.py:318-333); src < 0 = zero
    row, src[v, 0] == -2 = voxel outside every window: merged[v] = x[v] (quirk Q5)."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rows, x, src, weights):
        rows, x = rows.float().contiguous(), x.float().contiguous()
        src, weights = src.to(torch.int32).contiguous(), weights.float().contiguous()
        out = torch.empty_like(x)
        call("mssvt_interp_merge_fwd", x.shape[0], x.shape[1], ptr(src), ptr(weights),
             ptr(rows) if rows.numel() else None, ptr(x), ptr(out), stream())
        ctx.save_for_backward(src, weights)
        ctx.num_rows = rows.shape[0]
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        src, weights = ctx.saved_tensors
        grad_out = grad_out.float().contiguous()
        g_rows = torch.empty((ctx.num_rows, grad_out.shape[1]), dtype=torch.float32, device=grad_out.device)
        g_x = torch.empty_like(grad_out)
        call("mssvt_interp_merge_bwd", grad_out.shape[0], grad_out.shape[1], ctx.num_rows, ptr(src), ptr(weights),
             ptr(grad_out), ptr(g_rows) if ctx.num_rows else None, ptr(g_x), stream())
        return g_rows, g_x, None, None


interp_merge = InterpMerge.apply


class EmbedRows(Function):
    """Row sets of one block through mssvt_embed_rows_fwd / _bwd: for every set (rows, win, masked, c0, c1[, part]) the
    tensor xn[rows, c0:c1] + relu(pos_proj([xyz[rows] - centre[win] | centre[win]]))[c0:c1]; part = "x": the gathered
    rows alone, "pos": the embedding alone.  One Function for all sets of a block so that the gradient of xn and of
    pos_proj is accumulated in one buffer each."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, xn, pos_w, pos_b, xyz, centre, *sets):
        xn, pos_w, pos_b = xn.contiguous(), pos_w.contiguous(), pos_b.contiguous()
        xyz, centre = xyz.float().contiguous(), centre.float().contiguous()
        C, outs, packed = xn.shape[1], [], []
        for t in sets:
            rows, win, masked, c0, c1 = t[:5]
            part = t[5] if len(t) > 5 else "both"
            rows, win = rows.to(torch.int32).contiguous(), win.to(torch.int32).contiguous()
            masked = None if masked is None else masked.to(torch.uint8).contiguous()
            out = torch.empty((rows.shape[0], c1 - c0), dtype=torch.float32, device=xn.device)
            use_x, use_pos = part != "pos", part != "x"
            call("mssvt_embed_rows_fwd", rows.shape[0], c0, c1 - c0, C, ptr(rows), ptr(win), ptr(masked),
                 ptr(xn) if use_x else None, ptr(xyz), ptr(centre), ptr(pos_w) if use_pos else None, ptr(pos_b), ptr(out),
                 stream())
            outs.append(out)
            packed.append((rows, win, masked, c0, c1, use_x, use_pos))
        ctx.save_for_backward(pos_w, pos_b, xyz, centre)
        ctx.sets, ctx.xn_shape = packed, xn.shape
        return tuple(outs)

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, *grads):
        pos_w, pos_b, xyz, centre = ctx.saved_tensors
        g_xn = torch.zeros(ctx.xn_shape, dtype=torch.float32, device=pos_w.device)
        g_w, g_b = torch.zeros_like(pos_w), torch.zeros_like(pos_b)
        for (rows, win, masked, c0, c1, use_x, use_pos), g in zip(ctx.sets, grads):
            if g is None:
                continue
            g = g.float().contiguous()
            call("mssvt_embed_rows_bwd", rows.shape[0], c0, c1 - c0, ctx.xn_shape[1], ptr(rows), ptr(win), ptr(masked),
                 ptr(xyz), ptr(centre), ptr(pos_w) if use_pos else None, ptr(pos_b), ptr(g), ptr(g_xn) if use_x else None,
                 ptr(g_w), ptr(g_b), stream())
        return (g_xn, g_w, g_b, None, None) + (None,) * len(ctx.sets)


def embed_rows(xn, pos_w, pos_b, xyz, centre, sets):
    return EmbedRows.apply(xn, pos_w, pos_b, xyz, centre, *sets)


class SegmentMax(Function):
    """max over the compact rows of every window (mssvt_segment_max_fwd / _bwd): rows (R, C) -> (num_windows, C)"""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rows, lists, num_windows):
        rows = rows.contiguous()
        C = rows.shape[1]
        out = torch.empty((num_windows, C), dtype=torch.float32, device=rows.device)
        arg = torch.empty((num_windows, C), dtype=torch.int32, device=rows.device)
        call("mssvt_segment_max_fwd", num_windows, C, ptr(lists.key_off), ptr(rows), ptr(out), ptr(arg), stream())
        ctx.save_for_backward(arg)
        ctx.lists, ctx.num_rows = lists, rows.shape[0]
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        (arg,) = ctx.saved_tensors
        grad_out = grad_out.float().contiguous()
        g = torch.empty((ctx.num_rows, grad_out.shape[1]), dtype=torch.float32, device=grad_out.device)
        call("mssvt_segment_max_bwd", ctx.num_rows, grad_out.shape[1], ptr(ctx.lists.k_win), ptr(arg), ptr(grad_out), ptr(g),
             stream())
        return g, None, None


segment_max = SegmentMax.apply


class LayerNormRows(Function):
    """nn.LayerNorm over (N, C) rows: mssvt_layernorm forward, mssvt_layernorm_bwd backward (C in {64, 128})"""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, weight, bias, eps):
        x, weight, bias = x.contiguous(), weight.contiguous(), bias.contiguous()
        y = torch.empty_like(x)
        call("mssvt_layernorm", x.shape[0], None, x.shape[1], ptr(x), ptr(weight), ptr(bias), float(eps), ptr(y), stream())
        ctx.save_for_backward(x, weight)
        ctx.eps = float(eps)
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad_y):
        x, weight = ctx.saved_tensors
        grad_y = grad_y.float().contiguous()
        g_x, g_w, g_b = torch.empty_like(x), torch.empty_like(weight), torch.empty_like(weight)
        call("mssvt_layernorm_bwd", x.shape[0], x.shape[1], ptr(x), ptr(weight), ctx.eps, ptr(grad_y), ptr(g_x), ptr(g_w),
             ptr(g_b), stream())
        return g_x, g_w, g_b, None


def layer_norm_rows(norm, x):
    """nn.LayerNorm module `norm` applied to rows x through the hand-written kernels where they cover the shape"""
    if x.dim() == 2 and x.shape[1] in (64, 128) and x.shape[0] > 0 and norm.elementwise_affine and norm.bias is not None:
        return LayerNormRows.apply(x, norm.weight, norm.bias, norm.eps)
    return norm(x)


def _rows_view_ok(t):
    return (t.dim() == 2 and t.dtype == torch.float32 and t.stride(1) == 1 and t.stride(0) % 4 == 0
            and t.data_ptr() % 16 == 0)


_WGRAD_WS = {}       # (K, N) -> workspace floats of mssvt_linear_rows_wgrad


def _linear_terms():
    """split-TF32 operands (fp32-grade results) unless the user has allowed plain TF32 matmuls in torch"""
    return 1 if torch.backends.cuda.matmul.allow_tf32 else 3


class LinearRows(Function):
    """nn.Linear over (R, K) rows, K / N in {32, 64, 128}: forward and input gradient in mssvt_linear_rows_fwd, weight /
    bias gradient in mssvt_linear_rows_wgrad (one pass over the rows, deterministic).  x may be a column slice of a wider
    matrix (row stride passed down)."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x, weight, bias, relu):
        if not _rows_view_ok(x):
            x = x.contiguous()
        weight = weight.contiguous()
        N, K = weight.shape
        y = torch.empty((x.shape[0], N), dtype=torch.float32, device=x.device)
        terms = _linear_terms()
        call("mssvt_linear_rows_fwd", x.shape[0], K, N, terms, _ptr_view(x), x.stride(0), ptr(weight),
             None if bias is None else ptr(bias.contiguous()), int(bool(relu)), ptr(y), N, stream())
        ctx.save_for_backward(x, weight, y if relu else None)
        ctx.has_bias, ctx.terms = bias is not None, terms
        return y

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad_y):
        x, weight, y = ctx.saved_tensors
        N, K = weight.shape
        grad_y = grad_y.float()
        if y is not None:
            grad_y = grad_y * (y > 0)                       # ReLU fused into the forward epilogue
        if not _rows_view_ok(grad_y):
            grad_y = grad_y.contiguous()
        R, dev = x.shape[0], x.device
        g_x = None
        if ctx.needs_input_grad[0]:
            g_x = torch.empty((R, K), dtype=torch.float32, device=dev)
            call("mssvt_linear_rows_fwd", R, N, K, ctx.terms, _ptr_view(grad_y), grad_y.stride(0),
                 ptr(weight.t().contiguous()), None, 0, ptr(g_x), K, stream())
        g_w, g_b = torch.empty_like(weight), None
        if ctx.has_bias:
            g_b = torch.empty(N, dtype=torch.float32, device=dev)
        if (K, N) not in _WGRAD_WS:
            _WGRAD_WS[(K, N)] = call("mssvt_linear_rows_wgrad_workspace_floats", K, N)
        ws = torch.empty(_WGRAD_WS[(K, N)], dtype=torch.float32, device=dev)
        call("mssvt_linear_rows_wgrad", R, K, N, ctx.terms, _ptr_view(grad_y), grad_y.stride(0), _ptr_view(x), x.stride(0),
             ptr(ws), ptr(g_w), ptr(g_b), stream())
        return g_x, g_w, g_b, None


def linear_rows(layer, x, relu=False):
    """nn.Linear module `layer` applied to rows x (optionally followed by ReLU) through the hand-written kernels where
    they cover the shape, torch otherwise"""
    if not x.is_cuda:
        raise RuntimeError("mssvt_b200 operators run on CUDA tensors only (got a %s tensor); there is no CPU path" % x.device.type)
    if x.dim() == 2 and x.shape[0] > 0 and layer.in_features in (32, 64, 128) and layer.out_features in (32, 64, 128):
        return LinearRows.apply(x, layer.weight, layer.bias, relu)
    y = layer(x)
    return torch.relu(y) if relu else y
