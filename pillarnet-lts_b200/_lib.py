"""ctypes binding of libpillarnet_b200.so (include/pillarnet_b200.h).

The shim plays the role of the reference's pybind layer (det3d/ops/pillar_ops/src/pillar_api.cpp:10-21,
det3d/ops/iou3d_nms/src/iou3d_nms_api.cpp:11-19): it validates tensors the way the reference's
CHECK_INPUT does (det3d/ops/pillar_ops/src/cuda_utils.h:19-21 — CUDA + contiguous), unwraps
`data_ptr()` and the current stream, and turns error codes into RuntimeError.  There is no CPU
fallback: if the library is missing the import of any op fails loudly.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_size_t, c_uint32, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpillarnet_b200.so")

PN_F32, PN_BF16 = 0, 1
PN_IMPL_SIMT, PN_IMPL_TCGEN05 = 0, 1
PN_NBR_ANY, PN_NBR_SUBM_SORTED = 0, 1
PN_ERR_UNSUPPORTED = 4


class ConvArgs(Structure):
    _fields_ = [
        ("inp", c_void_p), ("in_dtype", c_int), ("in_ld", c_int),
        ("nbr", c_void_p), ("taps", c_int),
        ("weight", c_void_p), ("k_pad", c_int),
        ("scale", c_void_p), ("shift", c_void_p),
        ("residual", c_void_p), ("res_ld", c_int),
        ("out", c_void_p), ("out_dtype", c_int), ("out_ld", c_int), ("out_coff", c_int),
        ("relu", c_int),
        ("num_rows", c_void_p), ("rows_cap", c_int),
        ("cin", c_int), ("cout", c_int), ("rows_hint", c_int), ("out_hp", c_int), ("out_wp", c_int), ("in_rows", c_int),
        ("deconv_cout", c_int), ("deconv_hp_in", c_int), ("deconv_wp_in", c_int),
        ("nbr_kind", c_int), ("nbr_plan", c_void_p),
    ]


class RulebookLevel(Structure):
    """pn_rulebook_level"""
    _fields_ = [
        ("words", c_void_p), ("prefix", c_void_p), ("coords", c_void_p), ("num_rows", c_void_p),
        ("m_cap", c_int), ("nbr_down", c_void_p), ("nbr_subm", c_void_p),
    ]


class TaskArgs(Structure):
    _fields_ = [
        ("maps", c_void_p), ("ld", c_int),
        ("off_reg", c_int), ("off_height", c_int), ("off_dim", c_int), ("off_rot", c_int),
        ("off_vel", c_int), ("off_iou", c_int), ("off_hm", c_int),
        ("num_cls", c_int), ("H", c_int), ("W", c_int), ("stride", c_int),
        ("seg_base", c_int), ("per_class", c_int), ("activated", c_int),
    ]


# name -> (restype, argtypes); must list every symbol include/pillarnet_b200.h declares
SIGNATURES = {
    "pn_abi_version": (c_int, []),
    "pn_last_error": (c_char_p, []),
    "pn_device_sm_count": (c_int, []),
    "pn_launch_count": (ctypes.c_longlong, []),
    "pn_mask_words": (c_size_t, [c_int, c_int, c_int]),
    "pn_pillarize_scratch_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pn_pillarize": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float,
                             c_float, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                             c_void_p, c_size_t, c_void_p]),
    "pn_pfn_scatter_max": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float,
                                   c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                   c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pn_scatter_max_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "pn_rulebook_subm3x3": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                    c_void_p, c_void_p]),
    "pn_rulebook_down_scratch_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pn_rulebook_pyramid_scratch_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "pn_rulebook_pyramid3x3s2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                         c_size_t, c_void_p]),
    "pn_rulebook_down3x3s2": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                      c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                      c_void_p, c_size_t, c_void_p]),
    "pn_dense_nbr_table": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pn_conv_gather": (c_int, [POINTER(ConvArgs), c_int, c_void_p]),
    "pn_conv_window_plan_bytes": (c_size_t, [c_int]),
    "pn_conv_window_plan": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "pn_sizeof_conv_args": (c_size_t, []),
    "pn_sizeof_task_args": (c_size_t, []),
    "pn_conv3x3_small_cout": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                      c_void_p, c_int, c_void_p]),
    "pn_conv_dense3x3": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pn_conv_dense3x3_grouped": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                         c_void_p]),
    "pn_conv_dense3x3_grouped_shift": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                               c_void_p, c_int, c_void_p]),
    "pn_rulebook_block": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pn_roi_grid_bilinear": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "pn_roi_refine": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                              c_void_p, c_void_p]),
    "pn_bn_stats": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "pn_bn_finalize": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_float, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pn_bn_apply": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int,
                            c_void_p, c_int, c_void_p]),
    "pn_bn_bwd_stats": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "pn_bn_bwd_apply": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "pn_dense_to_sparse": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                   c_int, c_void_p]),
    "pn_conv_pack_weight_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pn_cast_f32_to_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "pn_cast_bf16_to_f32": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "pn_split_bf16x3": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "pn_sparse_to_dense": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                                   c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "pn_decode_candidates": (c_int, [POINTER(TaskArgs), c_int, c_int, c_int, c_float, POINTER(c_float),
                                     c_float, c_float, c_float, POINTER(c_float), c_void_p, c_int,
                                     c_void_p, c_void_p]),
    "pn_double_flip_merge": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "pn_point_features": (c_int, [c_void_p, c_int, c_int, c_float, c_float, c_float, c_float, c_float, c_float,
                                  c_void_p, c_void_p]),
    "pn_scatter_max": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "pn_rulebook_transpose": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pn_conv_wgrad": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int,
                              c_int, c_void_p, c_int, c_int, c_void_p]),
    "pn_boxes_aligned_overlap_bev": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "pn_assign_labels": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float,
                                 c_float, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
    "pn_merge_sweeps_scratch_bytes": (c_size_t, [c_int]),
    "pn_merge_sweeps": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pn_select_topk": (c_int, [POINTER(TaskArgs), c_int, c_int, c_int, POINTER(c_int), c_float, c_float,
                               c_float, POINTER(c_float), c_void_p, c_int, c_void_p, c_void_p, c_int,
                               c_void_p, c_void_p]),
    "pn_nms_scratch_bytes": (c_size_t, [c_int, c_int]),
    "pn_nms": (c_int, [c_int, c_int, c_int, POINTER(c_float), POINTER(c_int), POINTER(c_int), c_void_p,
                       c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_int, c_void_p, c_void_p,
                       c_void_p]),
    "pn_boxes_iou_bev": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "pn_nms_rotated": (c_int, [c_void_p, c_int, c_float, c_void_p, c_size_t, c_void_p, c_void_p,
                               c_void_p]),
    "pn_nms_normal": (c_int, [c_void_p, c_int, c_float, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "pn_boxes_overlap_bev": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "pn_compat_point_pillar_index": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                             c_void_p]),
    "pn_compat_pillar_indices": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pn_compat_gather_indice": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "pn_compat_gather_feature": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "pn_compat_gather_feature_grad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
}

_lib = None


def load():
    """Loads the shared library (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU/PyTorch fallback for the pillarnet_b200 hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().pn_last_error()
        raise RuntimeError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """device pointer of a tensor (None -> NULL)"""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def require_cuda(*tensors):
    """the reference's CHECK_INPUT (cuda_utils.h:19-21): CUDA tensor + contiguous."""
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("pillarnet_b200 ops need CUDA tensors (no CPU fallback)")
        if not t.is_contiguous():
            raise RuntimeError("pillarnet_b200 ops need contiguous tensors")


def require_cuda_rows(*tensors):
    """CUDA (rows, C) matrices whose rows are contiguous (a column slice of a wider matrix is fine: the ops take the
    row stride)."""
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("pillarnet_b200 ops need CUDA tensors (no CPU fallback)")
        if t.dim() != 2 or t.stride(1) != 1:
            raise RuntimeError("pillarnet_b200 ops need (rows, C) matrices with contiguous rows")


def farr(values):
    return (c_float * len(values))(*[float(v) for v in values])


def iarr(values):
    return (c_int * len(values))(*[int(v) for v in values])
