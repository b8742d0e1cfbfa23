"""Functional wrappers over the C ABI (one Python function per pn_* entry point).

All functions are asynchronous on torch's current CUDA stream, allocate their outputs with torch
(caller-owned memory, as the ABI requires) and never synchronise the host, so a whole forward pass
built from them can be captured in a CUDA graph.  Row counts only the device knows stay on the device
(`num` tensors) next to host capacities.
"""
import ctypes
from ctypes import byref, c_float, c_int, c_size_t, c_void_p

import torch

from . import _lib
from ._lib import (PN_BF16, PN_F32, PN_IMPL_SIMT, PN_IMPL_TCGEN05, ConvArgs, TaskArgs, check, farr,
                   iarr, ptr, require_cuda, stream_ptr)

_DT = {torch.float32: PN_F32, torch.bfloat16: PN_BF16}


def _i32(*shape, device):
    return torch.empty(*shape, dtype=torch.int32, device=device)


class RankTable:
    """Occupancy bitmask + popcount prefix of one (B,H,W) raster and its active rows.

    coords (cap,3) int32 [b,y,x] in ascending cell order; num: device int32 scalar (1,) ; cap: host.
    """

    __slots__ = ("words", "prefix", "coords", "num", "cap", "B", "H", "W", "_nbr_subm", "_plan_subm", "_train", "ready")

    def __init__(self, words, prefix, coords, num, cap, B, H, W):
        self.words, self.prefix, self.coords, self.num = words, prefix, coords, num
        self.cap, self.B, self.H, self.W = cap, B, H, W
        self._nbr_subm = None
        self._plan_subm = None
        self._train = None   # training-path caches (exact-size view, rulebooks): train.py
        self.ready = None    # event recorded when the table was complete (lets a side stream fork right there)

    def count(self):
        """host sync: number of active rows"""
        return min(int(self.num.item()), self.cap)

    def subm_nbr(self):
        """cached 3x3 submanifold neighbour table (cap,9) int32 (the `indice_key` cache of spconv)."""
        if self._nbr_subm is None:
            self._nbr_subm = rulebook_subm3x3(self)
        return self._nbr_subm

    def subm_plan(self):
        """cached tile plans of the submanifold table for the window-staged conv kernel (pn_conv_window_plan)"""
        if self._plan_subm is None:
            self._plan_subm = conv_window_plan(self.subm_nbr(), self.num, self.cap)
        return self._plan_subm


def pillarize(points, frame_offsets, n_frames, H, W, x0, y0, pillar_size, m_cap=None):
    """points (N,D) f32 cuda (frames concatenated), frame_offsets (B+1,) int32 cuda.

    Returns (RankTable, point_pillar (N,) int32).  See pn_pillarize in include/pillarnet_b200.h.
    """
    lib = _lib.load()
    require_cuda(points, frame_offsets)
    if points.dtype != torch.float32 or points.dim() != 2:
        raise RuntimeError("points must be (N,D) float32")
    if frame_offsets.dtype != torch.int32 or frame_offsets.numel() != n_frames + 1:
        raise RuntimeError("frame_offsets must be (B+1,) int32")
    dev = points.device
    N, D = points.shape
    if m_cap is None:
        m_cap = max(1, min(N, n_frames * H * W))
    nw = lib.pn_mask_words(n_frames, H, W)
    words = _i32(nw, device=dev)
    prefix = _i32(nw, device=dev)
    coords = _i32(m_cap, 3, device=dev)
    point_pillar = _i32(max(N, 1), device=dev)
    num = _i32(1, device=dev)
    sb = lib.pn_pillarize_scratch_bytes(n_frames, H, W)
    scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
    # the reference's CUDA expression divides by multiplying with the fp32 reciprocal
    inv = (torch.tensor(1.0, dtype=torch.float32) / torch.tensor(pillar_size, dtype=torch.float32)).item()
    check(lib.pn_pillarize(ptr(points), D, ptr(frame_offsets), N, n_frames, H, W,
                           c_float(_f32(x0)), c_float(_f32(y0)), c_float(inv), ptr(words), ptr(prefix),
                           ptr(coords), m_cap, ptr(point_pillar), ptr(num), ptr(scratch), c_size_t(sb),
                           stream_ptr()), "pn_pillarize")
    table = RankTable(words, prefix, coords, num, m_cap, n_frames, H, W)
    table.ready = torch.cuda.Event()
    table.ready.record()
    return table, point_pillar[:N]


def _f32(v):
    """round a Python double to fp32 the way torch does for scalar operands"""
    return torch.tensor(float(v), dtype=torch.float32).item()


def pfn_scatter_max(points, point_pillar, table, x0, y0, pillar_size, x_offset, y_offset, weight,
                    scale, shift, want_bf16=False, want_arg=False, n_live=None, want_f32=True):
    """Fused offset features + Linear + affine(BN) + ReLU + per-pillar max. Returns (f32, bf16|None, arg|None)."""
    lib = _lib.load()
    require_cuda(points, point_pillar)
    # weight/scale/shift are passed as HOST arrays (they become kernel launch parameters)
    weight, scale, shift = (t.detach().to("cpu", torch.float32).contiguous() for t in (weight, scale, shift))
    N, D = points.shape
    C = weight.shape[0]
    if weight.shape[1] != D + 2:
        raise RuntimeError(f"PFN weight must be (C,{D + 2})")
    dev = points.device
    if not want_f32 and not want_bf16:
        raise RuntimeError("nothing to compute")
    out = torch.empty(table.cap, C, dtype=torch.float32, device=dev) if want_f32 else None
    out_bf = torch.empty(table.cap, C, dtype=torch.bfloat16, device=dev) if want_bf16 else None
    arg = _i32(table.cap, C, device=dev) if want_arg else None
    inv = (torch.tensor(1.0, dtype=torch.float32) / torch.tensor(pillar_size, dtype=torch.float32)).item()
    check(lib.pn_pfn_scatter_max(ptr(points), D, N, ptr(n_live), ptr(point_pillar), ptr(table.num), table.cap,
                                 c_float(_f32(x0)), c_float(_f32(y0)), c_float(inv),
                                 c_float(_f32(pillar_size)), c_float(_f32(x_offset)),
                                 c_float(_f32(y_offset)), ptr(weight), ptr(scale), ptr(shift), C,
                                 ptr(out), ptr(out_bf), ptr(arg), stream_ptr()), "pn_pfn_scatter_max")
    return out, out_bf, arg


def scatter_max_grad(grad_out, arg, table, n_points):
    lib = _lib.load()
    require_cuda(grad_out, arg)
    C = grad_out.shape[1]
    grad_src = torch.zeros(n_points, C, dtype=torch.float32, device=grad_out.device)
    check(lib.pn_scatter_max_grad(ptr(grad_out), ptr(arg), ptr(table.num), table.cap, C, ptr(grad_src),
                                  stream_ptr()), "pn_scatter_max_grad")
    return grad_src


def rulebook_subm3x3(table):
    lib = _lib.load()
    nbr = _i32(table.cap, 9, device=table.coords.device)
    check(lib.pn_rulebook_subm3x3(ptr(table.words), ptr(table.prefix), ptr(table.coords), ptr(table.num),
                                  table.cap, table.H, table.W, ptr(nbr), stream_ptr()),
          "pn_rulebook_subm3x3")
    return nbr


def conv_window_plan(nbr, num, rows_cap):
    """pn_conv_window_plan: tile plans of a (rows_cap, 9) rulebook, to be passed as conv_gather(nbr_plan=...)."""
    lib = _lib.load()
    require_cuda(nbr)
    nb = lib.pn_conv_window_plan_bytes(rows_cap)
    plan = torch.empty(nb, dtype=torch.uint8, device=nbr.device)
    check(lib.pn_conv_window_plan(ptr(nbr), ptr(num), rows_cap, ptr(plan), c_size_t(nb), stream_ptr()),
          "pn_conv_window_plan")
    return plan


def rulebook_down3x3s2(table, out_cap=None):
    """Strided 3x3/s2/p1 rulebook. Returns (out RankTable, nbr (out_cap,9) int32 into input rows)."""
    lib = _lib.load()
    dev = table.coords.device
    Ho, Wo = (table.H + 2 - 3) // 2 + 1, (table.W + 2 - 3) // 2 + 1
    if out_cap is None:
        # every active input can activate up to 4 outputs (2 per axis)
        out_cap = max(1, min(4 * table.cap, table.B * Ho * Wo))
    nw = lib.pn_mask_words(table.B, Ho, Wo)
    words, prefix = _i32(nw, device=dev), _i32(nw, device=dev)
    coords = _i32(out_cap, 3, device=dev)
    num = _i32(1, device=dev)
    nbr = _i32(out_cap, 9, device=dev)
    sb = lib.pn_rulebook_down_scratch_bytes(table.B, Ho, Wo)
    scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
    check(lib.pn_rulebook_down3x3s2(ptr(table.words), ptr(table.prefix), ptr(table.coords),
                                    ptr(table.num), table.cap, table.B, table.H, table.W, ptr(words),
                                    ptr(prefix), ptr(coords), ptr(num), out_cap, ptr(nbr), ptr(scratch),
                                    c_size_t(sb), stream_ptr()), "pn_rulebook_down3x3s2")
    return RankTable(words, prefix, coords, num, out_cap, table.B, Ho, Wo), nbr


def rulebook_block(table, s, out_cap=None):
    """SparseConv2d(k = s, stride = s) rulebook (pn_rulebook_block): (out RankTable on the (H // s, W // s) grid,
    nbr (out_cap, s*s) int32 into the input rows)."""
    lib = _lib.load()
    dev = table.coords.device
    Ho, Wo = table.H // s, table.W // s
    if out_cap is None:
        out_cap = max(1, min(table.cap, table.B * Ho * Wo))     # blocks do not overlap: at most one output per input
    nw = lib.pn_mask_words(table.B, Ho, Wo)
    words, prefix = _i32(nw, device=dev), _i32(nw, device=dev)
    coords, num = _i32(out_cap, 3, device=dev), _i32(1, device=dev)
    nbr = _i32(out_cap, s * s, device=dev)
    sb = lib.pn_rulebook_down_scratch_bytes(table.B, Ho, Wo)
    scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
    check(lib.pn_rulebook_block(ptr(table.words), ptr(table.prefix), table.B, table.H, table.W, s, ptr(words),
                                ptr(prefix), ptr(coords), ptr(num), out_cap, ptr(nbr), ptr(scratch), c_size_t(sb),
                                stream_ptr()), "pn_rulebook_block")
    return RankTable(words, prefix, coords, num, out_cap, table.B, Ho, Wo), nbr


def roi_grid_bilinear(rois, grid_size, feat_rows, n_frames, H, W, C, x0, y0, cell, feat_coff=0, padded=False,
                      want_points=True):
    """rois (B, N, >=7) f32 -> (features (B, N, G*G, C) in feat's dtype, points (B, N, G*G, 2) f32 or None): RoI grid
    points + bilinear interpolation of the NHWC map rows (pn_roi_grid_bilinear)."""
    lib = _lib.load()
    require_cuda(rois, feat_rows)
    B, N, D = rois.shape
    if rois.dtype != torch.float32 or not rois.is_contiguous() or B != n_frames:
        raise RuntimeError("rois must be a contiguous (B, N, D) float32 tensor")
    P = grid_size * grid_size
    out = torch.empty(B, N, P, C, dtype=feat_rows.dtype, device=rois.device)
    pts = torch.empty(B, N, P, 2, dtype=torch.float32, device=rois.device) if want_points else None
    check(lib.pn_roi_grid_bilinear(ptr(rois), D, D - 1, B * N, N, grid_size, ptr(feat_rows),   # yaw = rois[:, -1] as the reference
                                   _DT[feat_rows.dtype], feat_rows.stride(0), feat_coff, n_frames, H, W,
                                   1 if padded else 0, C, c_float(_f32(x0)), c_float(_f32(y0)), c_float(_f32(cell)),
                                   ptr(pts), ptr(out), stream_ptr()), "pn_roi_grid_bilinear")
    return out, pts


def roi_refine(rois, reg, cls, roi_scores, roi_labels):
    """(boxes (B, N, code), scores (B, N), valid (B, N) bool) from the RoI head's outputs (pn_roi_refine)."""
    lib = _lib.load()
    require_cuda(rois, reg, cls, roi_scores)
    B, N, D = rois.shape
    code = reg.shape[-1]
    reg = reg.reshape(B * N, code).float().contiguous()
    cls = cls.reshape(B * N).float().contiguous()
    roi_scores = roi_scores.reshape(B * N).float().contiguous()
    labels = roi_labels.reshape(B * N).to(torch.int64).contiguous() if roi_labels is not None else None
    boxes = torch.empty(B, N, code, dtype=torch.float32, device=rois.device)
    scores = torch.empty(B, N, dtype=torch.float32, device=rois.device)
    valid = torch.empty(B, N, dtype=torch.uint8, device=rois.device)
    check(lib.pn_roi_refine(ptr(rois.contiguous()), D, ptr(reg), code, ptr(cls), ptr(roi_scores), ptr(labels), B * N,
                            ptr(boxes), ptr(scores), ptr(valid), stream_ptr()), "pn_roi_refine")
    return boxes, scores, valid.bool()


def rulebook_pyramid(table, n_levels):
    """All strided levels below `table` in n_levels + 2 launches (pn_rulebook_pyramid3x3s2).
    Returns [(out RankTable with its submanifold table attached, nbr_down (cap,9))] per level — the same
    tensors as rulebook_down3x3s2 + RankTable.subm_nbr() level by level."""
    lib = _lib.load()
    dev = table.coords.device
    B, H, W, cap = table.B, table.H, table.W, table.cap
    levels = (_lib.RulebookLevel * n_levels)()
    out = []
    for l in range(n_levels):
        H, W = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
        cap = max(1, min(4 * cap, B * H * W))      # every active input can activate up to 4 outputs
        nw = lib.pn_mask_words(B, H, W)
        words, prefix = _i32(nw, device=dev), _i32(nw, device=dev)
        coords, num = _i32(cap, 3, device=dev), _i32(1, device=dev)
        nbr_down, nbr_subm = _i32(cap, 9, device=dev), _i32(cap, 9, device=dev)
        lv = levels[l]
        lv.words, lv.prefix, lv.coords, lv.num_rows = ptr(words), ptr(prefix), ptr(coords), ptr(num)
        lv.m_cap, lv.nbr_down, lv.nbr_subm = cap, ptr(nbr_down), ptr(nbr_subm)
        t = RankTable(words, prefix, coords, num, cap, B, H, W)
        t._nbr_subm = nbr_subm
        out.append((t, nbr_down))
    sb = lib.pn_rulebook_pyramid_scratch_bytes(B, table.H, table.W, n_levels)
    scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
    check(lib.pn_rulebook_pyramid3x3s2(ptr(table.words), ptr(table.prefix), B, table.H, table.W, n_levels,
                                       ctypes.byref(levels), ptr(scratch), c_size_t(sb), stream_ptr()),
          "pn_rulebook_pyramid3x3s2")
    return out


_dense_nbr_cache = {}


def dense_nbr_table(mode, n_frames, H, W, stride, device, in_pad=False, out_pad=False):
    """Static gather table for dense NHWC convs; cached per (mode,B,H,W,stride,padding,device).
    in_pad/out_pad: the input/output rows index a zero-padded (H+2,W+2) map."""
    key = (mode, n_frames, H, W, stride, bool(in_pad), bool(out_pad), str(device))
    t = _dense_nbr_cache.get(key)
    if t is None:
        lib = _lib.load()
        po = 2 if out_pad else 0
        if mode == 0:
            Ho, Wo = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
            t = _i32(n_frames * (Ho + po) * (Wo + po), 9, device=device)
        else:
            t = _i32(n_frames * (2 * H + po) * (2 * W + po), 4, device=device)
        flags = (1 if in_pad else 0) | (2 if out_pad else 0)
        check(lib.pn_dense_nbr_table(mode, n_frames, H, W, stride, flags, ptr(t), stream_ptr()),
              "pn_dense_nbr_table")
        _dense_nbr_cache[key] = t
    return t


def conv_gather(inp, weight, nbr, taps, cin, cout, out, *, in_ld=None, k_pad=None, scale=None,
                shift=None, residual=None, res_ld=None, out_ld=None, out_coff=0, relu=False, num=None,
                rows_cap=None, impl=PN_IMPL_SIMT, in_ptr_offset=0, rows_hint=0, out_hw_pad=None, deconv=None,
                nbr_kind=0, nbr_plan=None):
    """out[o, coff:coff+cout] = act((sum_t W_t . in[nbr[o,t]]) * scale + shift + residual).
    deconv = (cout_per_tap, Hp_in, Wp_in): the 2x2/s2 transposed conv as one GEMM (pn_conv_args.deconv_*).

    `inp`/`out` are 2-D channels-last tensors (possibly wider than cin/cout: in_ld/out_ld are the
    row strides in elements; in_ptr_offset selects a channel slice of the input)."""
    lib = _lib.load()
    a = ConvArgs()
    es = inp.element_size()
    a.inp = c_void_p(inp.data_ptr() + in_ptr_offset * es)
    a.in_dtype = _DT[inp.dtype]
    a.in_ld = in_ld if in_ld is not None else inp.stride(0)
    a.nbr = ptr(nbr).value
    a.taps = taps
    if weight.dtype != inp.dtype:
        raise RuntimeError("weight dtype must match the input dtype")
    a.weight = weight.data_ptr()
    a.k_pad = k_pad if k_pad is not None else weight.stride(0)
    a.scale = ptr(scale).value
    a.shift = ptr(shift).value
    a.residual = ptr(residual).value
    a.res_ld = res_ld if res_ld is not None else (residual.stride(0) if residual is not None else 0)
    a.out = out.data_ptr()
    a.out_dtype = _DT[out.dtype]
    a.out_ld = out_ld if out_ld is not None else out.stride(0)
    a.out_coff = out_coff
    a.relu = 1 if relu else 0
    a.num_rows = ptr(num).value
    a.rows_cap = rows_cap if rows_cap is not None else out.shape[0]
    a.cin, a.cout = cin, cout
    a.rows_hint = int(rows_hint)
    a.out_hp, a.out_wp = (out_hw_pad if out_hw_pad is not None else (0, 0))
    a.in_rows = inp.shape[0]
    a.deconv_cout, a.deconv_hp_in, a.deconv_wp_in = deconv if deconv is not None else (0, 0, 0)
    a.nbr_kind = int(nbr_kind)
    a.nbr_plan = ptr(nbr_plan).value
    if residual is not None and residual.dtype != out.dtype:
        raise RuntimeError("residual dtype must match the output dtype")
    check(lib.pn_conv_gather(byref(a), impl, stream_ptr()), "pn_conv_gather")
    return out


def conv3x3_small_cout(inp, in_ld, cin, n_frames, H, W, groups, n_groups, wbuf, out, in_padded=False):
    """grouped tiny-Cout dense 3x3 conv (all CenterHead final convs in one launch); see the C header."""
    lib = _lib.load()
    check(lib.pn_conv3x3_small_cout(ptr(inp), in_ld, cin, n_frames, H, W, 1 if in_padded else 0, ptr(groups),
                                    n_groups, ptr(wbuf),
                                    ptr(out), out.stride(0), stream_ptr()), "pn_conv3x3_small_cout")
    return out


def conv_dense3x3(inp, in_coff, cin, n_frames, H, W, weight, cout, out, *, scale=None, shift=None, out_coff=0,
                  out_compact=False, relu=False, tile_hint=0, out_group_cols=0):
    """3x3/s1/p1 conv on zero-padded NHWC rows (B*(H+2)*(W+2), ld) -> padded (or compact) rows; see the C header.
    out_group_cols = gc: `out` is the planar (cout/gc * rows, gc) matrix (one contiguous map per channel group)."""
    lib = _lib.load()
    check(lib.pn_conv_dense3x3(ptr(inp), inp.stride(0), in_coff, cin, n_frames, H, W, ptr(weight), weight.stride(0),
                               cout, ptr(scale), ptr(shift), ptr(out), _DT[out.dtype], out.stride(0), out_coff,
                               1 if out_compact else 0, out_group_cols, 1 if relu else 0, tile_hint, stream_ptr()),
          "pn_conv_dense3x3")
    return out


def conv_dense3x3_grouped(inp, in_coff, cin, n_groups, n_frames, H, W, weight, shift, group_tab, out, *,
                          scale=None, out_compact=True, relu=False, in_planar=False):
    """n_groups independent small-Cout 3x3 convs on padded NHWC rows in one tensor-core launch (C header)."""
    lib = _lib.load()
    check(lib.pn_conv_dense3x3_grouped(ptr(inp), inp.stride(0), in_coff, cin, n_groups, n_frames, H, W, ptr(weight),
                                       weight.stride(0), ptr(scale), ptr(shift), ptr(group_tab), ptr(out),
                                       _DT[out.dtype], out.stride(0), 1 if out_compact else 0, 1 if relu else 0,
                                       1 if in_planar else 0, stream_ptr()), "pn_conv_dense3x3_grouped")
    return out


def conv_dense3x3_grouped_shift(inp, n_groups, n_frames, H, W, weight, shift, group_tab, out):
    """the grouped last conv of the CenterHead branches as one 1x1 GEMM per branch + nine shifted sums (C header);
    raises NotImplementedError when the map is too wide for the kernel's halo (callers fall back to the grouped conv)"""
    lib = _lib.load()
    rc = lib.pn_conv_dense3x3_grouped_shift(ptr(inp), n_groups, n_frames, H, W, ptr(weight), ptr(shift), ptr(group_tab),
                                            ptr(out), out.stride(0), stream_ptr())
    if rc == _lib.PN_ERR_UNSUPPORTED:
        raise NotImplementedError("pn_conv_dense3x3_grouped_shift: map too wide")
    check(rc, "pn_conv_dense3x3_grouped_shift")
    return out


def pack_weight_bf16(w_f32_2d, k_pad=None):
    """(Cout,K) f32 -> (Cout,k_pad) bf16 with K zero-padded to a multiple of 64."""
    lib = _lib.load()
    require_cuda(w_f32_2d)
    cout, k = w_f32_2d.shape
    if k_pad is None:
        k_pad = (k + 63) // 64 * 64
    out = torch.empty(cout, k_pad, dtype=torch.bfloat16, device=w_f32_2d.device)
    check(lib.pn_conv_pack_weight_bf16(ptr(w_f32_2d), cout, k, k_pad, ptr(out), stream_ptr()),
          "pn_conv_pack_weight_bf16")
    return out


def cast_rows(inp, dtype, num=None):
    lib = _lib.load()
    rows, cols = inp.shape
    out = torch.empty(rows, cols, dtype=dtype, device=inp.device)
    if inp.dtype == torch.float32 and dtype == torch.bfloat16:
        fn, name = lib.pn_cast_f32_to_bf16, "pn_cast_f32_to_bf16"
    elif inp.dtype == torch.bfloat16 and dtype == torch.float32:
        fn, name = lib.pn_cast_bf16_to_f32, "pn_cast_bf16_to_f32"
    else:
        raise RuntimeError("unsupported cast")
    check(fn(ptr(inp), inp.stride(0), ptr(out), out.stride(0), cols, ptr(num), rows, stream_ptr()), name)
    return out


def split_bf16x3(inp, cols, col0=0, num=None):
    """(rows, ld) f32 -> (rows, 3*cols) bf16 [hi | lo | hi] of columns [col0, col0 + cols) (C header)."""
    lib = _lib.load()
    require_cuda(inp)
    if inp.dtype != torch.float32:
        raise RuntimeError("split_bf16x3 wants f32 rows")
    rows = inp.shape[0]
    out = torch.empty(rows, 3 * cols, dtype=torch.bfloat16, device=inp.device)
    check(lib.pn_split_bf16x3(c_void_p(inp.data_ptr() + 4 * col0), inp.stride(0), ptr(out), cols, ptr(num), rows,
                              stream_ptr()), "pn_split_bf16x3")
    return out


def sparse_to_dense(feat, table, C, out=None, out_coff=0, padded=False):
    """NHWC densify: returns (B*H*W, out_ld) rows whose [coff,coff+C) columns hold the features
    (padded: (B*(H+2)*(W+2), out_ld) rows of the zero-bordered map)."""
    lib = _lib.load()
    n_cells = table.B * (table.H + (2 if padded else 0)) * (table.W + (2 if padded else 0))
    if out is None:
        out = torch.empty(n_cells, C, dtype=feat.dtype, device=feat.device)
    check(lib.pn_sparse_to_dense(ptr(feat), _DT[feat.dtype], feat.stride(0), ptr(table.words),
                                 ptr(table.prefix), table.B, table.H, table.W, C, ptr(out),
                                 out.stride(0), out_coff, 1 if padded else 0, stream_ptr()), "pn_sparse_to_dense")
    return out


def double_flip_merge(rows, offsets, num_cls, n_frames_out, H, W):
    """center_head.py:233-304: merge the 4 flipped views of every frame into one activated map (same columns)."""
    lib = _lib.load()
    if not rows.is_cuda or rows.dtype != torch.float32 or rows.stride(1) != 1:
        raise RuntimeError("double_flip_merge: rows must be a CUDA f32 (n, cols) view with unit column stride")
    n_cols = rows.shape[1]
    t = make_task_args(rows, offsets, num_cls, H, W, 1, 0, False)
    out = torch.empty(n_frames_out * H * W, n_cols, dtype=torch.float32, device=rows.device)
    check(lib.pn_double_flip_merge(byref(t), n_frames_out, n_cols, ptr(out), out.stride(0), stream_ptr()),
          "pn_double_flip_merge")
    return out


def make_task_args(maps, offsets, num_cls, H, W, stride, seg_base, per_class, activated=False):
    t = TaskArgs()
    t.maps = maps.data_ptr()
    t.ld = maps.stride(0)
    t.off_reg = offsets.get("reg", -1)
    t.off_height = offsets.get("height", -1)
    t.off_dim = offsets.get("dim", -1)
    t.off_rot = offsets.get("rot", -1)
    t.off_vel = offsets.get("vel", -1)
    t.off_iou = offsets.get("iou", -1)
    t.off_hm = offsets.get("hm", -1)
    t.num_cls, t.H, t.W, t.stride = num_cls, H, W, stride
    t.seg_base, t.per_class = seg_base, 1 if per_class else 0
    t.activated = 1 if activated else 0
    return t


def boxes_iou_bev(boxes_a, boxes_b):
    """drop-in for iou3d_nms_cuda.boxes_iou_bev_gpu (iou3d_nms.cpp:90-110)."""
    lib = _lib.load()
    require_cuda(boxes_a, boxes_b)
    out = torch.empty(boxes_a.shape[0], boxes_b.shape[0], dtype=torch.float32, device=boxes_a.device)
    check(lib.pn_boxes_iou_bev(ptr(boxes_a), boxes_a.shape[0], ptr(boxes_b), boxes_b.shape[0], ptr(out),
                               stream_ptr()), "pn_boxes_iou_bev")
    return out


def nms_rotated(boxes, thr):
    """drop-in for iou3d_nms_cuda.nms_gpu (iou3d_nms.cpp:113-159): boxes (n,7) already score-sorted.

    Returns (keep int32 (n,), num_keep int32 (1,)) on the device (no host round trip)."""
    lib = _lib.load()
    require_cuda(boxes)
    n = boxes.shape[0]
    cap = max(64, (n + 63) // 64 * 64)
    sb = (cap * 12 * 4 + 256) + (lib.pn_nms_scratch_bytes(1, cap) + 256) + (cap * 11 * 4 + 256) + 512
    scratch = torch.empty(sb, dtype=torch.uint8, device=boxes.device)
    keep = _i32(cap, device=boxes.device)
    num = _i32(1, device=boxes.device)
    check(lib.pn_nms_rotated(ptr(boxes), n, c_float(thr), ptr(scratch), c_size_t(sb), ptr(keep), ptr(num),
                             stream_ptr()), "pn_nms_rotated")
    return keep, num


# ---- training-path operators ------------------------------------------------------

def point_features(points, x0, y0, pillar_size, x_offset, y_offset):
    """(n, 2+D) f32 [x-ctr_x, y-ctr_y, p...] (pillar_utils.py:51-56); points must be in range."""
    lib = _lib.load()
    require_cuda(points)
    n, D = points.shape
    out = torch.empty(n, D + 2, dtype=torch.float32, device=points.device)
    inv = (torch.tensor(1.0, dtype=torch.float32) / torch.tensor(pillar_size, dtype=torch.float32)).item()
    check(lib.pn_point_features(ptr(points), D, n, c_float(_f32(x0)), c_float(_f32(y0)), c_float(inv),
                                c_float(_f32(pillar_size)), c_float(_f32(x_offset)), c_float(_f32(y_offset)),
                                ptr(out), stream_ptr()), "pn_point_features")
    return out


def scatter_max(src, index, n_pillars, want_arg=True):
    """drop-in for pillar_cuda.scatter_max_wrapper: returns (out (M,C) f32, arg (M,C) i32 | None)."""
    lib = _lib.load()
    require_cuda(src, index)
    if src.dtype != torch.float32 or index.dtype != torch.int32 or not src.is_contiguous() or not index.is_contiguous():
        raise RuntimeError("scatter_max: src must be contiguous f32 (L,C), index contiguous int32 (L)")
    L, C = src.shape
    out = torch.empty(n_pillars, C, dtype=torch.float32, device=src.device)
    arg = _i32(n_pillars, C, device=src.device) if want_arg else None
    check(lib.pn_scatter_max(ptr(src), ptr(index), L, n_pillars, C, ptr(out), ptr(arg), stream_ptr()),
          "pn_scatter_max")
    return out, arg


def scatter_max_grad_flat(grad_out, arg, n_points):
    """drop-in for pillar_cuda.scatter_max_grad_wrapper with an exact pillar count (M = arg.shape[0])."""
    lib = _lib.load()
    require_cuda(grad_out, arg)
    M, C = arg.shape
    grad_out = grad_out.float().contiguous()
    grad_src = torch.zeros(n_points, C, dtype=torch.float32, device=grad_out.device)
    check(lib.pn_scatter_max_grad(ptr(grad_out), ptr(arg), ptr(None), M, C, ptr(grad_src), stream_ptr()),
          "pn_scatter_max_grad")
    return grad_src


def rulebook_transpose(nbr, n_in, num_out=None):
    """(n_in, taps) int32 input-stationary table of an output-stationary rulebook (see the C header)."""
    lib = _lib.load()
    require_cuda(nbr)
    n_out, taps = nbr.shape
    nbr_t = _i32(n_in, taps, device=nbr.device)
    check(lib.pn_rulebook_transpose(ptr(nbr), ptr(num_out), n_out, taps, n_in, ptr(nbr_t), stream_ptr()),
          "pn_rulebook_transpose")
    return nbr_t


def conv_wgrad(x, dy, nbr, taps, cin, cout, *, num=None, rows=None, impl=PN_IMPL_SIMT):
    """dW (cout, taps*cin) f32 = sum_o dy[o]^T x[nbr[o,t]]  (see pn_conv_wgrad)."""
    lib = _lib.load()
    require_cuda(x, dy)
    if x.dtype != dy.dtype:
        raise RuntimeError("conv_wgrad: x and dy must have the same dtype")
    rows = dy.shape[0] if rows is None else rows
    dw = torch.empty(cout, taps * cin, dtype=torch.float32, device=x.device)
    check(lib.pn_conv_wgrad(ptr(x), _DT[x.dtype], x.stride(0), ptr(dy), _DT[dy.dtype], dy.stride(0), ptr(nbr), taps,
                            ptr(num), rows, cin, cout, ptr(dw), dw.stride(0), impl, stream_ptr()), "pn_conv_wgrad")
    return dw


def boxes_aligned_overlap_bev(boxes_a, boxes_b):
    """drop-in for iou3d_nms_cuda.boxes_aligned_overlap_bev_gpu: (n,) BEV intersection areas (pcdet boxes)."""
    lib = _lib.load()
    require_cuda(boxes_a, boxes_b)
    n = boxes_a.shape[0]
    out = torch.empty(n, dtype=torch.float32, device=boxes_a.device)
    check(lib.pn_boxes_aligned_overlap_bev(ptr(boxes_a.contiguous()), ptr(boxes_b.contiguous()), n, ptr(out),
                                           stream_ptr()), "pn_boxes_aligned_overlap_bev")
    return out


def assign_labels_task(gt_boxes, gt_cls, num_cls, H, W, x0, y0, cell, gaussian_overlap, min_radius):
    """pn_assign_labels for one task.  gt_boxes (B,M,9|7) f32, gt_cls (B,M) int32 (1-based, 0 = empty).
    Returns dict(hm (B,H,W,K) f32, ind/cat (B,M) i64, mask (B,M) u8, anno_box (B,M,10), gt_box (B,M,7))."""
    lib = _lib.load()
    require_cuda(gt_boxes, gt_cls)
    if gt_boxes.dtype != torch.float32 or gt_cls.dtype != torch.int32:
        raise RuntimeError("assign_labels: gt_boxes float32, gt_cls int32")
    B, M, D = gt_boxes.shape
    dev = gt_boxes.device
    out = dict(hm=torch.empty(B, H, W, num_cls, dtype=torch.float32, device=dev),
               ind=torch.empty(B, M, dtype=torch.int64, device=dev),
               mask=torch.empty(B, M, dtype=torch.uint8, device=dev),
               cat=torch.empty(B, M, dtype=torch.int64, device=dev),
               anno_box=torch.empty(B, M, 10, dtype=torch.float32, device=dev),
               gt_box=torch.empty(B, M, 7, dtype=torch.float32, device=dev))
    mr = list(min_radius) if isinstance(min_radius, (list, tuple)) else [int(min_radius)]
    check(lib.pn_assign_labels(ptr(gt_boxes), D, ptr(gt_cls), B, M, num_cls, H, W, c_float(_f32(x0)), c_float(_f32(y0)),
                               c_float(cell), c_float(_f32(gaussian_overlap)), iarr(mr), len(mr), ptr(out["hm"]),
                               ptr(out["ind"]), ptr(out["mask"]), ptr(out["cat"]), ptr(out["anno_box"]),
                               ptr(out["gt_box"]), stream_ptr()), "pn_assign_labels")
    return out


def merge_sweeps(raw, sweep_offsets, transforms, time_lags, min_distance=1.0, n_feat=4, out=None, out_base=None):
    """pn_merge_sweeps: raw (n_raw, in_dim) f32 CUDA (key frame + sweeps concatenated), sweep_offsets host ints,
    transforms list of 4x4 / 3x4 arrays or None per sweep, time_lags host floats.
    Returns (out (cap, n_feat+1) f32, n_total (1,) int32 device)."""
    import numpy as np
    lib = _lib.load()
    require_cuda(raw)
    n_sweeps = len(sweep_offsets) - 1
    n_raw = int(sweep_offsets[-1])
    T = np.full((n_sweeps, 12), np.nan, np.float64)
    for k, m in enumerate(transforms):
        if m is not None:
            T[k] = np.asarray(m, np.float64)[:3, :4].reshape(12)
    lag = np.asarray(time_lags, np.float32)
    if out is None:
        out = torch.empty(max(n_raw, 1), n_feat + 1, dtype=torch.float32, device=raw.device)
    n_total = _i32(1, device=raw.device)
    sb = lib.pn_merge_sweeps_scratch_bytes(n_raw)
    scratch = torch.empty(sb, dtype=torch.uint8, device=raw.device)
    check(lib.pn_merge_sweeps(ptr(raw), raw.shape[1], n_feat, iarr([int(v) for v in sweep_offsets]), n_sweeps,
                              T.ctypes.data_as(c_void_p), lag.ctypes.data_as(c_void_p), c_float(_f32(min_distance)),
                              ptr(out_base), ptr(out), out.shape[0], ptr(n_total), ptr(scratch), c_size_t(sb),
                              stream_ptr()), "pn_merge_sweeps")
    return out, n_total


# ---- training: batch-statistics BN on the live rows (bn_train.cu), sync-free -----------------------------------------

def bn_train_forward(x, num, gamma, beta, running_mean, running_var, eps, momentum, residual=None, relu=False):
    """y = act(BN_batch(x) + residual) over the rows < *num (the rows of the capacity beyond them are left untouched);
    running statistics are updated in place like nn.BatchNorm1d in train mode.  Returns (y, mean, rstd)."""
    lib = _lib.load()
    _lib.require_cuda_rows(x, residual)
    rows_cap, C = x.shape
    dev = x.device
    sums = torch.empty(2 * C, dtype=torch.float32, device=dev)
    stats = torch.empty(4, C, dtype=torch.float32, device=dev)     # mean, rstd, scale, shift
    check(lib.pn_bn_stats(ptr(x), _DT[x.dtype], x.stride(0), ptr(num), rows_cap, C, ptr(sums), stream_ptr()),
          "pn_bn_stats")
    check(lib.pn_bn_finalize(ptr(sums), ptr(num), rows_cap, C, ptr(gamma), ptr(beta), c_float(eps), c_float(momentum),
                             ptr(running_mean), ptr(running_var), ptr(stats[0]), ptr(stats[1]), ptr(stats[2]),
                             ptr(stats[3]), stream_ptr()), "pn_bn_finalize")
    y = torch.empty(rows_cap, C, dtype=x.dtype, device=dev)
    if residual is not None and residual.dtype != x.dtype:
        raise RuntimeError("residual dtype must match x")
    check(lib.pn_bn_apply(ptr(x), _DT[x.dtype], x.stride(0), ptr(stats[2]), ptr(stats[3]), ptr(residual),
                          residual.stride(0) if residual is not None else 0, 1 if relu else 0, ptr(num), rows_cap, C,
                          ptr(y), y.stride(0), stream_ptr()), "pn_bn_apply")
    return y, stats[0], stats[1]


def bn_train_backward(dy, y, x, mean, rstd, gamma, relu, num, want_dres, zero_tail=False):
    """Returns (dx, dres | None, dgamma (C) f32, dbeta (C) f32).  zero_tail: the rows of dx at or above *num are zeros
    (for consumers that read the whole capacity, e.g. a torch matmul) instead of unwritten."""
    lib = _lib.load()
    _lib.require_cuda_rows(x, y)
    rows_cap, C = x.shape
    dev = x.device
    dy = dy.contiguous()
    require_cuda(dy)
    if dy.dtype != x.dtype:
        dy = dy.to(x.dtype)
    sums = torch.empty(2 * C, dtype=torch.float32, device=dev)
    check(lib.pn_bn_bwd_stats(ptr(dy), _DT[x.dtype], dy.stride(0), ptr(y), y.stride(0) if y is not None else 0, ptr(x),
                              x.stride(0), ptr(mean), ptr(rstd), 1 if relu else 0, ptr(num), rows_cap, C, ptr(sums),
                              stream_ptr()), "pn_bn_bwd_stats")
    dx = (torch.zeros if zero_tail else torch.empty)(rows_cap, C, dtype=x.dtype, device=dev)
    dres = torch.empty(rows_cap, C, dtype=x.dtype, device=dev) if want_dres else None
    check(lib.pn_bn_bwd_apply(ptr(dy), _DT[x.dtype], dy.stride(0), ptr(y), y.stride(0) if y is not None else 0, ptr(x),
                              x.stride(0), ptr(mean), ptr(rstd), ptr(gamma), ptr(sums), 1 if relu else 0, ptr(num),
                              rows_cap, C, ptr(dx), dx.stride(0), ptr(dres), dres.stride(0) if want_dres else 0,
                              stream_ptr()), "pn_bn_bwd_apply")
    return dx, dres, sums[C:], sums[:C]


def dense_to_sparse(dense_rows, table, C):
    """rows (cap, C) of a compact NHWC map (B*H*W, ld) at the table's sites (rows beyond the live count untouched)"""
    lib = _lib.load()
    require_cuda(dense_rows)
    out = torch.empty(table.cap, C, dtype=dense_rows.dtype, device=dense_rows.device)
    check(lib.pn_dense_to_sparse(ptr(dense_rows), _DT[dense_rows.dtype], dense_rows.stride(0), ptr(table.coords),
                                 ptr(table.num), table.cap, table.H, table.W, C, ptr(out), out.stride(0), stream_ptr()),
          "pn_dense_to_sparse")
    return out
