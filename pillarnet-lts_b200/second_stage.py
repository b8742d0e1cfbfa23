"""Pillar R-CNN second stage, inference path (SURVEY §8 f rank 3): BEV feature fusion + RoI grid pooling.

Mirrors det3d/models/second_stage/bev_interpolation.py:17-308 (`BEVFeature`, `BEVStrideFeature`: same constructor
kwargs, module tree and state_dict keys): a top-down `ConvTranspose2d(k = s, stride = s)` on the neck's last map, one
lateral layer per backbone stage (dense `ConvTranspose2d(k = s, stride = s)` or spconv `SparseConv2d(k = s, stride = s)`),
channel concat, a 3x3 fusion conv, then 7x7 rotated grid points per RoI and bilinear interpolation.

Here: every `ConvTranspose2d(k = s, stride = s)` + BN + ReLU is ONE GEMM over the input pixels (N = s*s*Cout, folded
affine repeated per tap) followed by a pixel shuffle; the sparse lateral conv runs on `pn_rulebook_block` +
`pn_conv_gather`; the fusion conv on the dense tensor-core kernel; grid points + interpolation in `pn_roi_grid_bilinear`.
"""
import numpy as np
import torch
from torch import nn

from . import config, ops
from .layers import (DenseMap, Lowered, SparseConv2d, SparseSequential, build_norm_layer, dense_conv3x3, lower,
                     run_conv, use_padded_layout, weight_matrix)
from .registry import SECOND_STAGE
from .sparse import SparseConvTensor


def _lower_deconv_ks(conv, bn):
    """ConvTranspose2d(k = s, stride = s, bias=False) + BN as one GEMM: weight rows (dy*s + dx)*Cout + o over K = Cin"""
    base = lower(conv, bn)                       # folded affine (+ cache key)
    cache = conv.__dict__.setdefault("_pn_deconv_ks", {})
    prec = config.get_precision()
    hit = cache.get(prec)
    if hit is not None and hit.key == base.key:
        return hit
    s = conv.kernel_size[0]
    cin, cout = conv.in_channels, conv.out_channels
    # weight (Cin, Cout, s, s): out(s*y + dy, s*x + dx, o) = sum_c in(y, x, c) W[c, o, dy, dx]
    w = conv.weight.detach().float().permute(2, 3, 1, 0).reshape(s * s * cout, cin).contiguous()
    lw = Lowered()
    if prec == "bf16":
        lw.weight = ops.pack_weight_bf16(w)
    elif prec == "bf16x3":
        from .layers import split_weight_bf16x3
        lw.weight = split_weight_bf16x3(w, cin)
    else:
        lw.weight = w
    lw.k_pad = lw.weight.shape[1]
    lw.scale, lw.shift = base.scale.repeat(s * s).contiguous(), base.shift.repeat(s * s).contiguous()
    lw.key = base.key
    cache[prec] = lw
    return lw


def deconv_ks(x, conv, bn, relu=True):
    """x: DenseMap -> (B, s*H, s*W, Cout) NHWC tensor of `ConvTranspose2d(k = s, stride = s)` + BN + ReLU"""
    s = conv.kernel_size[0]
    assert conv.kernel_size == (s, s) and conv.stride == (s, s) and conv.bias is None
    cout = conv.out_channels
    lw = _lower_deconv_ks(conv, bn)
    rows_in = x.n_rows
    y = run_conv(x.rows, lw, None, 1, x.C, s * s * cout, rows_in, relu=relu, in_ld=x.rows.stride(0),
                 in_ptr_offset=x.coff)
    p = x.pad
    y = y.view(x.B, x.H + 2 * p, x.W + 2 * p, s, s, cout)
    if p:
        y = y[:, 1:-1, 1:-1]
    if s == 1:
        return y.reshape(x.B, x.H, x.W, cout)
    return y.permute(0, 1, 3, 2, 4, 5).reshape(x.B, x.H * s, x.W * s, cout)


def sparse_block_conv(sp, seq):
    """SparseSequential(SparseConv2d(k = s, stride = s, bias=True), BN1d, ReLU) -> SparseConvTensor on the coarser grid"""
    conv, bn = seq[0], seq[1]
    s = conv.stride
    table, nbr = ops.rulebook_block(sp.table, s)
    lw = lower(conv, bn)
    taps = s * s
    if config.get_precision() == "bf16" and taps > 9:
        # the tensor-core gather kernel holds nine taps per tile: wider blocks run on the FMA kernel
        out = torch.empty(table.cap, conv.out_channels, dtype=sp.feat.dtype, device=sp.feat.device)
        ops.conv_gather(sp.feat, lw.weight, nbr, taps, conv.in_channels, conv.out_channels, out, k_pad=lw.k_pad,
                        scale=lw.scale, shift=lw.shift, relu=True, num=table.num, rows_cap=table.cap,
                        impl=ops.PN_IMPL_SIMT)
    else:
        out = run_conv(sp.feat, lw, nbr, taps, conv.in_channels, conv.out_channels, table.cap, num=table.num, relu=True)
    return SparseConvTensor(out, table, (table.H, table.W), sp.batch_size)


class _BEVFusion(nn.Module):
    """common forward of BEVFeature / BEVStrideFeature (bev_interpolation.py:125-159, 273-308)"""

    def _as_map(self, t):
        if isinstance(t, SparseConvTensor):
            d = t.dense_nhwc(padded=use_padded_layout())
            return DenseMap(d, t.batch_size, t.table.H, t.table.W, t.feat.shape[1], 0, 1 if use_padded_layout() else 0)
        return DenseMap.from_nchw(t)

    def _fused_map_stride1(self, bev_feature, backbone_features):
        """every transposed conv has k = stride = 1 (the shipped config): each is a 1x1 GEMM over the rows of its input
        map, written straight into its channel slice of the concatenated (padded) map — no shuffle, cat or pad copies"""
        srcs = [(self._as_map(bev_feature), self.top_down_conv)]
        for k, src in enumerate(self.lat_conv_name):
            srcs.append((self._as_map(backbone_features[src]), self.lat_conv[k]))
        x0 = srcs[0][0]
        cout = self.top_down_conv[0].out_channels
        for m, _ in srcs:
            if (m.B, m.H, m.W, m.pad) != (x0.B, x0.H, x0.W, x0.pad):
                raise RuntimeError("second-stage feature maps disagree")
        cat = torch.empty(x0.n_rows, cout * len(srcs), dtype=x0.rows.dtype, device=x0.rows.device)
        for k, (m, seq) in enumerate(srcs):
            run_conv(m.rows, _lower_deconv_ks(seq[0], seq[1]), None, 1, m.C, cout, m.n_rows, relu=True, out=cat,
                     out_coff=k * cout, in_ld=m.rows.stride(0), in_ptr_offset=m.coff,
                     out_hw_pad=(x0.H + 2, x0.W + 2) if x0.pad else None)      # border rows stay zero
        x = DenseMap(cat, x0.B, x0.H, x0.W, cat.shape[1], 0, x0.pad)
        return dense_conv3x3(x, self.fusion_conv[0], self.fusion_conv[1], relu=True)

    def fused_map(self, bev_feature, backbone_features):
        """-> DenseMap of the fusion conv's output (B, H_out, W_out, share_channels)"""
        if (config.get_precision() != "bf16x3" and self.top_down_conv[0].kernel_size == (1, 1)
                and all(t == "dense" and self.lat_conv[k][0].kernel_size == (1, 1)
                        for k, t in enumerate(self.lat_tensor_type))):
            return self._fused_map_stride1(bev_feature, backbone_features)
        parts = [deconv_ks(self._as_map(bev_feature), self.top_down_conv[0], self.top_down_conv[1])]
        for k, src in enumerate(self.lat_conv_name):
            cur = backbone_features[src]
            if self.lat_tensor_type[k] == "dense":
                parts.append(deconv_ks(self._as_map(cur), self.lat_conv[k][0], self.lat_conv[k][1]))
            else:
                if not isinstance(cur, SparseConvTensor):
                    raise RuntimeError(f"lateral layer {src}: a sparse tensor is expected")
                o = sparse_block_conv(cur, self.lat_conv[k])
                parts.append(o.dense_nhwc().view(o.batch_size, o.table.H, o.table.W, -1))
        B, H, W, _ = parts[0].shape
        for p in parts:
            if tuple(p.shape[:3]) != (B, H, W):
                raise RuntimeError(f"second-stage feature maps disagree: {[tuple(q.shape) for q in parts]}")
        cat = torch.cat(parts, dim=-1)
        x = DenseMap.from_nchw(cat.permute(0, 3, 1, 2))
        return dense_conv3x3(x, self.fusion_conv[0], self.fusion_conv[1], relu=True)

    def forward(self, example):
        rois = example["rois"]
        B, N = rois.shape[:2]
        fused = self.fused_map(example["bev_feature"], example["backbone_features"])
        cell = np.float32(self.out_stride * self.pillar_size)      # `bev_stride * self.pillar_size`, then an fp32 division
        feats, pts = ops.roi_grid_bilinear(rois.float().contiguous(), self.grid_size, fused.rows, B, fused.H, fused.W,
                                           fused.C, self.point_cloud_range[0], self.point_cloud_range[1], float(cell),
                                           feat_coff=fused.coff, padded=bool(fused.pad))
        example["roi_features"] = feats.view(B, N, -1)
        example["point_features"] = feats                  # (B, N, G*G, C)
        example["point_coords"] = pts                      # (B, N, G*G, 2)
        return example


def _lateral(in_channels, out_channels, stride, sparse):
    if not sparse:
        return nn.Sequential(nn.ConvTranspose2d(in_channels, out_channels, stride, stride=stride, bias=False),
                             build_norm_layer(dict(type="BN", momentum=0.01, eps=1e-3), out_channels)[1], nn.ReLU())
    return SparseSequential(SparseConv2d(in_channels, out_channels, kernel_size=stride, stride=stride, padding=0, bias=True),
                            build_norm_layer(dict(type="BN1d", momentum=0.01, eps=1e-3), out_channels)[1], nn.ReLU())


@SECOND_STAGE.register_module
class BEVFeature(_BEVFusion):
    """bev_interpolation.py:17-159 (top-down from conv4's stride; lateral dense iff stride > 1 or out_stride == 8)"""

    def __init__(self, feature_sources, pillar_size, pc_range, out_stride=4, grid_size=7, in_channels=256,
                 share_channels=64, backbone_channels=None, backbone_strides=None):
        super().__init__()
        self.pillar_size, self.point_cloud_range, self.grid_size = pillar_size, pc_range, grid_size
        self.lat_conv, self.lat_conv_name, self.lat_tensor_type = nn.ModuleList(), [], []
        opt_strides, names, chans = [1, 2, 4, 8], ["conv1", "conv2", "conv3", "conv4"], [32, 64, 128, 256]
        assert out_stride in opt_strides
        out_channels = chans[opt_strides.index(out_stride)]
        assert out_channels <= backbone_channels[names[opt_strides.index(out_stride)]]
        stride = int(backbone_strides["conv4"] / out_stride)
        self.top_down_conv = _lateral(in_channels, out_channels, stride, sparse=False)
        c_in = out_channels
        for src in feature_sources:
            if src not in names:
                continue
            stride = backbone_strides[src] / out_stride
            if stride > 1 or (out_stride == 8 and stride == 1):
                self.lat_conv.append(_lateral(backbone_channels[src], out_channels, int(stride), sparse=False))
                self.lat_tensor_type.append("dense")
            else:
                self.lat_conv.append(_lateral(backbone_channels[src], out_channels, int(np.round(1 / stride)), sparse=True))
                self.lat_tensor_type.append("sparse")
            c_in += out_channels
            self.lat_conv_name.append(src)
        self.fusion_conv = nn.Sequential(nn.Conv2d(c_in, share_channels, 3, stride=1, padding=1, bias=True),
                                         build_norm_layer(dict(type="BN", momentum=0.01, eps=1e-3), share_channels)[1],
                                         nn.ReLU())
        self.out_stride = out_stride


@SECOND_STAGE.register_module
class BEVStrideFeature(_BEVFusion):
    """bev_interpolation.py:162-308 (top-down from conv3's stride; lateral dense iff stride >= 1)"""

    def __init__(self, feature_sources, pillar_size, pc_range, out_stride=4, grid_size=7, in_channels=128,
                 share_channels=64, backbone_channels=None, backbone_strides=None):
        super().__init__()
        self.pillar_size, self.point_cloud_range, self.grid_size = pillar_size, pc_range, grid_size
        self.lat_conv, self.lat_conv_name, self.lat_tensor_type = nn.ModuleList(), [], []
        opt_strides, names, chans = [1, 2, 4], ["conv1", "conv2", "conv3"], [32, 64, 128]
        assert out_stride in opt_strides
        out_channels = chans[opt_strides.index(out_stride)]
        assert out_channels <= backbone_channels[names[opt_strides.index(out_stride)]]
        stride = int(backbone_strides["conv3"] / out_stride)
        self.top_down_conv = _lateral(in_channels, out_channels, stride, sparse=False)
        c_in = out_channels
        for src in feature_sources:
            if src not in ["conv1", "conv2", "conv3", "conv4"]:
                continue
            stride = backbone_strides[src] / out_stride
            if stride >= 1:
                self.lat_conv.append(_lateral(backbone_channels[src], out_channels, int(stride), sparse=False))
                self.lat_tensor_type.append("dense")
            else:
                self.lat_conv.append(_lateral(backbone_channels[src], out_channels, int(np.round(1 / stride)), sparse=True))
                self.lat_tensor_type.append("sparse")
            c_in += out_channels
            self.lat_conv_name.append(src)
        self.fusion_conv = nn.Sequential(nn.Conv2d(c_in, share_channels, 3, stride=1, padding=1, bias=True),
                                         build_norm_layer(dict(type="BN", eps=1e-3, momentum=0.01), share_channels)[1],
                                         nn.ReLU())
        self.out_stride = out_stride
