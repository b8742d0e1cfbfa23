"""`iou3d_nms_cuda` with the reference's pybind signatures (det3d/ops/iou3d_nms/src/iou3d_nms_api.cpp:11-19,
iou3d_nms.h:10-14) over libpillarnet_b200.so.  Boxes are (n,7) f32 CUDA [x,y,z,dx,dy,dz,heading]; outputs are the
caller's pre-allocated tensors.  `nms_gpu` / `nms_normal_gpu` fill a CPU int64 `keep` and return the number kept,
exactly like iou3d_nms.cpp:113-159 — which costs the one device->host copy the reference's API imposes; the product
path (CenterHead.predict here) keeps everything on the device instead.

The two *_cpu functions of the reference module are CPU code outside the hot path and are not provided (no CPU
fallback in this package): they raise."""
import torch

from .. import _lib
from .._lib import check, ptr, stream_ptr
from ctypes import c_float, c_size_t


def _boxes(*ts):
    for t in ts:
        if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2 or t.shape[1] != 7 or not t.is_contiguous():
            raise RuntimeError("boxes must be a contiguous CUDA float32 (n,7) tensor")   # iou3d_nms.cpp:14-25


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    _boxes(boxes_a, boxes_b)
    check(_lib.load().pn_boxes_iou_bev(ptr(boxes_a), boxes_a.shape[0], ptr(boxes_b), boxes_b.shape[0], ptr(ans_iou),
                                      stream_ptr()), "pn_boxes_iou_bev")
    return 1


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    _boxes(boxes_a, boxes_b)
    check(_lib.load().pn_boxes_overlap_bev(ptr(boxes_a), boxes_a.shape[0], ptr(boxes_b), boxes_b.shape[0],
                                          ptr(ans_overlap), stream_ptr()), "pn_boxes_overlap_bev")
    return 1


def boxes_aligned_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    _boxes(boxes_a, boxes_b)
    if boxes_a.shape[0] != boxes_b.shape[0]:
        raise RuntimeError("aligned overlap needs equally many boxes")
    check(_lib.load().pn_boxes_aligned_overlap_bev(ptr(boxes_a), ptr(boxes_b), boxes_a.shape[0], ptr(ans_overlap),
                                                  stream_ptr()), "pn_boxes_aligned_overlap_bev")
    return 1


def _nms(fn_name, boxes, keep, thr):
    _boxes(boxes)
    if keep.is_cuda or keep.dtype != torch.int64 or not keep.is_contiguous():
        raise RuntimeError("keep must be a contiguous CPU int64 tensor")                 # iou3d_nms.cpp:118-119
    lib = _lib.load()
    n = boxes.shape[0]
    if n == 0:
        return 0
    # the device sweep handles 4096 boxes per call; the reference has no limit, so larger inputs are refused loudly
    if n > 4096:
        raise RuntimeError(f"{fn_name}: {n} boxes exceed the 4096-box capacity of the device-side sweep")
    cap = max(64, (n + 63) // 64 * 64)
    sb = (cap * 12 * 4 + 256) + (lib.pn_nms_scratch_bytes(1, cap) + 256) + (cap * 11 * 4 + 256) + 512
    scratch = torch.empty(sb, dtype=torch.uint8, device=boxes.device)
    keep_d = torch.empty(cap, dtype=torch.int32, device=boxes.device)
    num_d = torch.empty(1, dtype=torch.int32, device=boxes.device)
    check(getattr(lib, fn_name)(ptr(boxes), n, c_float(thr), ptr(scratch), c_size_t(sb), ptr(keep_d), ptr(num_d),
                                stream_ptr()), fn_name)
    num = int(num_d.item())
    keep[:num] = keep_d[:num].to(torch.int64).cpu()
    return num


def nms_gpu(boxes, keep, nms_overlap_thresh):
    """iou3d_nms.cpp:113-159: boxes already sorted by score; returns num_to_keep, keep[:num] filled."""
    return _nms("pn_nms_rotated", boxes, keep, float(nms_overlap_thresh))


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    """iou3d_nms.cpp:162-207 (axis-aligned IoU)."""
    return _nms("pn_nms_normal", boxes, keep, float(nms_overlap_thresh))


def boxes_iou_bev_cpu(*args):
    raise NotImplementedError("CPU function outside the hot path; pillarnet_lts_b200 has no CPU fallback")


def boxes_aligned_iou_bev_cpu(*args):
    raise NotImplementedError("CPU function outside the hot path; pillarnet_lts_b200 has no CPU fallback")
