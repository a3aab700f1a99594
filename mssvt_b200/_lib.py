"""ctypes loader for libmssvt_b200.so (the C-ABI of include/mssvt_b200.h).

There is no CPU fallback and no alternative backend: if the library is missing, or a call
returns an error, this raises.  PyTorch is used by the callers only to own device memory and
streams; every kernel launched here is hand-written CUDA for sm_100a.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# MSSVT_B200_LIB: another build of the same library (tools/build_trace_lib.sh, kernel experiments); never a fallback
LIB_PATH = os.environ.get("MSSVT_B200_LIB") or os.path.join(_HERE, "libmssvt_b200.so")
_lib = None

P, I, L, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float

# name -> argument ctypes (every function returns int unless listed in _RESTYPES)
_SIGNATURES = {
    "mssvt_fill_i32": [P, L, I, P],
    "mssvt_count_samples": [I, I, P, P, P, P],
    "mssvt_exclusive_scan": [I, P, P, I, P, P, P],
    "mssvt_voxel_world_coords": [I, P, P, P, P, P],
    "mssvt_build_hash_table": [I, I, I, I, I, I, P, P, P, P],
    "mssvt_hash_lookup": [I, I, P, P, P, P, P],
    "mssvt_window_partition_workspace_bytes": [I],
    "mssvt_window_partition": [I] * 11 + [P, P, P, P, P, L, P],
    "mssvt_window_list_workspace_bytes": [I] * 5,
    "mssvt_window_list": [I] * 10 + [P, P, P, P, L, P],
    "mssvt_gather_two_window": [I] * 16 + [P] * 14 + [P],
    "mssvt_gather_one_window": [I] * 10 + [P] * 5 + [P],
    "mssvt_group_features": [I, I, I, I, P, P, P, P, P, P],
    "mssvt_group_features_grad": [I, I, I, I, I, P, P, P, P, P, P],
    "mssvt_fps": [I, I, I, P, P, P, P],
    "mssvt_fps_log2_block": [I],
    "mssvt_gather_points": [I, I, I, I, P, P, P, P],
    "mssvt_three_nn": [I, I, I, P, P, P, P, P],
    "mssvt_group_points": [I, I, I, I, I, P, P, P, P],
    "mssvt_group_points_grad": [I, I, I, I, I, P, P, P, P],
    "mssvt_grid_index_words": [I, I, I, I],
    "mssvt_grid_index_build": [I, I, I, I, I, P, P, P, P, P, P],
    "mssvt_block_geometry": [I] * 15 + [P] * 6 + [I, P, P, P, P, P, I] + [P] * 14 + [P],
    "mssvt_block_queries": [I] * 4 + [P] * 9 + [P],
    "mssvt_query_src": [I, P, I, P, P, P, P],
    "mssvt_window_rows": [I] * 8 + [P, I, P, P, P, P, P, P, P],
    "mssvt_layernorm": [I, P, I, P, P, P, F, P, P],
    "mssvt_block_attention": [P, I, P, I] + [P] * 11 + [P],
    "mssvt_attention_tiles": [I] * 4 + [P] * 11 + [P],
    "mssvt_block_attention_tc": [I] * 7 + [F] + [P] * 11 + [I] + [P] * 16 + [I, P, P] + [P],
    "mssvt_compress_attention": [P, I, P, I] + [P] * 6 + [P],
    "mssvt_compress_tiles": [I, I] + [P] * 9 + [P],
    "mssvt_compress_attention_tc": [I, I, I, I, F] + [P] * 12 + [I] + [P] * 11 + [P],
    "mssvt_ffn": [P, I, P, I, P, P, P, P, P, P],
    "mssvt_pack_operand_tf32": [P, I, I, I, P, P],
    "mssvt_pack_operand_bf16": [P, I, I, P, P],
    "mssvt_pack_operand_bf16x2": [P, I, I, P, P],
    "mssvt_tma_copy_rows": [P, P, I, P],
    "mssvt_ffn_tc": [I, I, I, I, F] + [P] * 6 + [I, P, P, P, P, P, P, P, F, P] + [P] * 6 + [I, P],
    "mssvt_dense_scatter": [I, P, I, I, I, I, I, P, P, P, P],
    "mssvt_sizeof_attn_shape": [],
    "mssvt_sizeof_ffn_shape": [],
    "mssvt_vfe_bitmap_words": [I, I, I, I],
    "mssvt_vfe_voxelize": [I, P, I, I, I, I, I] + [P] * 9 + [P],
    "mssvt_vfe_features": [I, P, I, I, I, I, I] + [P] * 5 + [I, P, P, I, P, P, I, P, P] + [P],
    "mssvt_ragged_attention_fwd": [I, I, F, I, P, P, P, P, I, P, I, P, I, P, I, P, P],
    "mssvt_ragged_attention_bwd": [I, I, F, I, I, P, P, P, P, P, P, I, P, I, P, I, P, I, P, P, I, P, P, I, P, I, P, I, P],
    "mssvt_embed_rows_fwd": [I, I, I, I] + [P] * 9 + [P],
    "mssvt_embed_rows_bwd": [I, I, I, I] + [P] * 11 + [P],
    "mssvt_layernorm_bwd": [I, I, P, P, F, P, P, P, P, P],
    "mssvt_linear_rows_fwd": [I, I, I, I, P, I, P, P, I, P, I, P],
    "mssvt_linear_rows_wgrad_workspace_floats": [I, I],
    "mssvt_linear_rows_wgrad": [I, I, I, I, P, I, P, I, P, P, P, P],
    "mssvt_ragged_lists_count": [I, P, P, P, P, P],
    "mssvt_ragged_lists_fill": [I, P, I, I] + [P] * 15 + [P],
    "mssvt_ragged_merge_map": [I, I] + [P] * 7 + [P],
    "mssvt_compress_lists_count": [I, P, I, P, P, P, P],
    "mssvt_compress_lists_fill": [I, P, I] + [P] * 6 + [P],
    "mssvt_segment_max_fwd": [I, I, P, P, P, P, P],
    "mssvt_segment_max_bwd": [I, I, P, P, P, P, P],
    "mssvt_interp_merge_fwd": [I, I, P, P, P, P, P, P],
    "mssvt_interp_merge_bwd": [I, I, I, P, P, P, P, P, P],
    "mssvt_last_cuda_error": [],
    "mssvt_version": [],
    "mssvt_launch_count": [],
}
_RESTYPES = {"mssvt_linear_rows_wgrad_workspace_floats": L, "mssvt_window_partition_workspace_bytes": L, "mssvt_window_list_workspace_bytes": L, "mssvt_grid_index_words": L, "mssvt_version": ctypes.c_char_p,
             "mssvt_vfe_bitmap_words": L,
             "mssvt_launch_count": L}
_NO_STATUS = {"mssvt_linear_rows_wgrad_workspace_floats", "mssvt_window_partition_workspace_bytes", "mssvt_window_list_workspace_bytes", "mssvt_grid_index_words", "mssvt_vfe_bitmap_words", "mssvt_fps_log2_block", "mssvt_version",
              "mssvt_sizeof_attn_shape", "mssvt_sizeof_ffn_shape", "mssvt_last_cuda_error",
              "mssvt_launch_count"}
_ERRORS = {-1: "invalid argument", -2: "CUDA launch/runtime error", -3: "workspace too small"}

EXPORTS = tuple(_SIGNATURES)

MAX_GROUPS = 4


class AttnShape(ctypes.Structure):
    """Mirror of `struct AttnShape` in csrc/block.cu (field for field)."""
    _fields_ = ([(n, I) for n in ("C", "G", "hd", "nq", "nk_total", "nk", "cap1", "interp", "pos_layers")] +
                [("heads", I * MAX_GROUPS), ("sd", I * MAX_GROUPS), ("c0", I * MAX_GROUPS),
                 ("off_pos_w", I), ("off_pos_b", I), ("off_pos2_w", I), ("off_pos2_b", I),
                 ("off_wq", I * MAX_GROUPS), ("off_bq", I * MAX_GROUPS),
                 ("off_wkv", I * MAX_GROUPS), ("off_bkv", I * MAX_GROUPS),
                 ("off_wp", I * MAX_GROUPS), ("off_bp", I * MAX_GROUPS),
                 ("total_floats", I), ("scale", F), ("win_cell", F * 3), ("lo", F * 3)])


class FfnShape(ctypes.Structure):
    """Mirror of `struct FfnShape` in csrc/block.cu."""
    _fields_ = ([(n, I) for n in ("C", "F", "C_out", "mode", "off_ln_g", "off_ln_b", "off_w1", "off_b1",
                                  "off_w2", "off_b2", "off_wo", "off_bo", "total_floats")] + [("eps", F)])


def load():
    """Load the library once; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "mssvt_b200: %s not found. Build it with `make -C mssvt_b200/csrc` (or "
            "`python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU or PyTorch "
            "fallback for this path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, I)
    if lib.mssvt_sizeof_attn_shape() != ctypes.sizeof(AttnShape) or \
            lib.mssvt_sizeof_ffn_shape() != ctypes.sizeof(FfnShape):
        raise RuntimeError("mssvt_b200: descriptor layout mismatch between _lib.py and block.cu")
    _lib = lib
    return lib


# When set to a list, call() brackets every entry point with CUDA events on the current stream
# and appends (name, start_event, end_event): bench.py's per-kernel breakdown.  None = off.
PROFILE = None


def call(name, *args):
    """Invoke an entry point; raise on a non-zero status."""
    fn = getattr(load(), name)
    if PROFILE is not None and name not in _NO_STATUS:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        PROFILE.append((name, e0, e1))
    else:
        rc = fn(*args)
    if name in _NO_STATUS:
        return rc
    if rc != 0:
        extra = ""
        if rc == -2:
            extra = " (cudaError %d)" % load().mssvt_last_cuda_error()
        raise RuntimeError("%s failed: %s%s" % (name, _ERRORS.get(rc, rc), extra))
    return rc


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("mssvt_b200 operators run on CUDA tensors only (got a %s tensor); "
                           "there is no CPU path" % t.device.type)
    if not t.is_contiguous():
        raise RuntimeError("mssvt_b200 operators need contiguous tensors")
    return ctypes.c_void_p(t.data_ptr())


def stream():
    """cudaStream_t of torch's current stream (raw handle: no Stream object is built per call)"""
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


def host_floats(values):
    return (F * len(values))(*[float(v) for v in values])
