"""`pillar_cuda` with the reference's pybind signatures (det3d/ops/pillar_ops/src/pillar_api.cpp:10-21) over
libpillarnet_b200.so.  Every wrapper takes the tensors the reference's C++ wrapper takes (pillar_ops.cpp:15-55,
group_ops.cpp, scatter_ops.cpp:7-41), checks them like CHECK_INPUT (cuda_utils.h:19-21: CUDA + contiguous),
fills the caller's output tensors in place on torch's current stream and returns 1."""
import torch

from .. import _lib
from .._lib import check, ptr, require_cuda, stream_ptr


def _i32(*ts):
    for t in ts:
        if t.dtype != torch.int32:
            raise RuntimeError("expected an int32 tensor")


def _f32(*ts):
    for t in ts:
        if t.dtype != torch.float32:
            raise RuntimeError("expected a float32 tensor")


def create_point_pillar_index_stack_wrapper(pts_xy, pts_batch_cnt, pillars_mask, point_pillar_index):
    """pillar_ops.cpp:15-36: pts_xy (N,2) i32, pts_batch_cnt (B) i32, pillars_mask (B,H,W) bool, point_pillar_index (N) i32."""
    require_cuda(pts_xy, pts_batch_cnt, pillars_mask, point_pillar_index)
    _i32(pts_xy, pts_batch_cnt, point_pillar_index)
    if pillars_mask.dtype not in (torch.bool, torch.uint8) or pillars_mask.dim() != 3:
        raise RuntimeError("pillars_mask must be a (B,H,W) bool tensor")
    B, H, W = pillars_mask.shape
    if pts_batch_cnt.numel() != B or pts_xy.shape[0] != point_pillar_index.shape[0]:
        raise RuntimeError("inconsistent shapes")
    check(_lib.load().pn_compat_point_pillar_index(ptr(pts_xy), ptr(pts_batch_cnt), pts_xy.shape[0], B, H, W,
                                                  ptr(pillars_mask), ptr(point_pillar_index), stream_ptr()),
          "pn_compat_point_pillar_index")
    return 1


def create_pillar_indices_wrapper(pillars_position, pillar_indices):
    """pillar_ops.cpp:39-55: pillars_position (B,H,W) i32, pillar_indices (M,3) i32."""
    require_cuda(pillars_position, pillar_indices)
    _i32(pillars_position, pillar_indices)
    B, H, W = pillars_position.shape
    check(_lib.load().pn_compat_pillar_indices(ptr(pillars_position), B, H, W, ptr(pillar_indices), stream_ptr()),
          "pn_compat_pillar_indices")
    return 1


def gather_indice_wrapper(index, indices, outs):
    require_cuda(index, indices, outs)
    _i32(index, indices, outs)
    check(_lib.load().pn_compat_gather_indice(ptr(index), ptr(indices), index.shape[0], ptr(outs), stream_ptr()),
          "pn_compat_gather_indice")
    return 1


def gather_feature_wrapper(index, features, outs):
    require_cuda(index, features, outs)
    _i32(index)
    _f32(features, outs)
    check(_lib.load().pn_compat_gather_feature(ptr(index), ptr(features), index.shape[0], features.shape[1], ptr(outs),
                                              stream_ptr()), "pn_compat_gather_feature")
    return 1


def gather_feature_grad_wrapper(index, grad_outs, grad_features):
    require_cuda(index, grad_outs, grad_features)
    _i32(index)
    _f32(grad_outs, grad_features)
    check(_lib.load().pn_compat_gather_feature_grad(ptr(index), ptr(grad_outs), index.shape[0], grad_outs.shape[1],
                                                   ptr(grad_features), stream_ptr()), "pn_compat_gather_feature_grad")
    return 1


def scatter_max_wrapper(index, src, arg, out):
    """scatter_ops.cpp:7-23: index (L) i32, src (L,C) f32, arg (M,C) i32, out (M,C) f32 (zero-initialised by the caller:
    the 0 floor of scatter_ops_gpu.cu:13-22).  pn_scatter_max writes out and arg completely (same 0 floor)."""
    require_cuda(index, src, arg, out)
    _i32(index, arg)
    _f32(src, out)
    L, C = src.shape
    check(_lib.load().pn_scatter_max(ptr(src), ptr(index), L, out.shape[0], C, ptr(out), ptr(arg), stream_ptr()),
          "pn_scatter_max")
    return 1


def scatter_max_grad_wrapper(grad_out, arg, grad_src):
    """scatter_ops.cpp:26-41: grad_src[arg[m,c], c] = grad_out[m,c] (grad_src zero-initialised by the caller)."""
    require_cuda(grad_out, arg, grad_src)
    _i32(arg)
    _f32(grad_out, grad_src)
    M, C = arg.shape
    check(_lib.load().pn_scatter_max_grad(ptr(grad_out), ptr(arg), ptr(None), M, C, ptr(grad_src), stream_ptr()),
          "pn_scatter_max_grad")
    return 1
