"""Parameter containers and the fused conv(+BN+ReLU+residual) runner shared by backbone/neck/head.

The containers keep the reference's parameter names and layouts so checkpoints load unchanged
(det3d/torchie/trainer/checkpoint.py:67-137): sparse conv weights are
(Cout,kH,kW,Cin) as in spconv 2.x, dense convs are plain nn.Conv2d / nn.ConvTranspose2d /
nn.BatchNorm{1,2}d modules used as *containers* — their torch forward is never called on the
product path; the math runs in libpillarnet_b200 through ops.conv_gather.
"""
import math

import torch
from torch import nn

from . import config, ops


class SparseReLU(nn.ReLU):
    """spconv.pytorch.SparseReLU stand-in (backbones/base.py:3-4,96,105): marker module; the ReLU is
    fused into the producing conv's epilogue."""


class _SparseConvBase(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1,
                 bias=True, indice_key=None):
        super().__init__()
        if dilation != 1:
            raise NotImplementedError("only dilation 1 sparse convs are on the PillarNet hot path")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding, self.indice_key = kernel_size, stride, padding, indice_key
        # spconv 2.x layout (Cout, kH, kW, Cin); checkpoint.py:78-87
        self.weight = nn.Parameter(torch.empty(out_channels, kernel_size, kernel_size, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.in_channels * self.kernel_size * self.kernel_size
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def weight_2d(self):
        return self.weight.detach().reshape(self.out_channels, -1)


class SubMConv2d(_SparseConvBase):
    """spconv.pytorch.SubMConv2d container (backbones/base.py:43-52)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1,
                 bias=True, indice_key=None):
        if kernel_size != 3:
            raise NotImplementedError("only 3x3 submanifold convs are supported")
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, bias, indice_key)


class SparseConv2d(_SparseConvBase):
    """spconv.pytorch.SparseConv2d container: 3x3 / stride 2 / padding 1 (PillarResNet.py:87,95,103) or the
    non-overlapping kernel = stride = s, padding 0 of the second stage's lateral layers
    (second_stage/bev_interpolation.py:66-72)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=2, padding=1, dilation=1,
                 bias=False, indice_key=None):
        block = kernel_size == stride and padding == 0 and 1 <= stride <= 8
        if not block and not (kernel_size == 3 and stride == 2 and padding == 1):
            raise NotImplementedError("strided sparse convs: 3x3 / stride 2 / padding 1, or kernel = stride <= 8 / padding 0")
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, bias, indice_key)


class SparseSequential(nn.Sequential):
    """spconv.pytorch.SparseSequential container: children keep their integer names."""


class Sequential(nn.Sequential):
    """det3d.models.utils.Sequential (models/utils/misc.py:21-95): nn.Sequential with .add()."""

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
        self.add_module(name, module)


def build_norm_layer(cfg, num_features):
    """det3d/models/utils/norm.py:68-109 for the types the PillarNet configs use."""
    cfg = dict(cfg)
    t = cfg.pop("type")
    cfg.setdefault("eps", 1e-5)
    if t == "BN":
        layer = nn.BatchNorm2d(num_features, **cfg)
    elif t == "BN1d":
        layer = nn.BatchNorm1d(num_features, **cfg)
    elif t == "LN":
        layer = nn.LayerNorm(num_features, **cfg)
    else:
        raise NotImplementedError(f"norm type {t}")
    return t.lower(), layer


# ---- lowering: conv (+bias) + BN(eval) -> packed weight, per-channel scale/shift -----------------

def _versions(*tensors):
    return tuple((t.data_ptr(), t._version) if t is not None else None for t in tensors)


def weight_matrix(conv):
    """(Cout, taps*Cin) fp32 matrix in tap-major/channel-minor order for any supported container."""
    if isinstance(conv, _SparseConvBase):
        return conv.weight.detach().reshape(conv.out_channels, -1)
    if isinstance(conv, nn.ConvTranspose2d):
        # weight (Cin,Cout,2,2): out(2y+dy,2x+dx,o) = sum_c in(y,x,c) W[c,o,dy,dx]; tap = dy*2+dx
        return conv.weight.detach().permute(1, 2, 3, 0).reshape(conv.out_channels, -1)
    if isinstance(conv, nn.Conv2d):
        return conv.weight.detach().permute(0, 2, 3, 1).reshape(conv.out_channels, -1)
    if isinstance(conv, nn.Conv1d) and conv.kernel_size == (1,):        # the RoI head's per-RoI "FC" layers
        return conv.weight.detach().reshape(conv.out_channels, -1)
    if isinstance(conv, nn.Linear):
        return conv.weight.detach()
    raise TypeError(type(conv))


def fold_affine(conv_bias, bn, cout, device):
    """y = conv*scale + shift with eval-mode BN folded: scale = g/sqrt(var+eps), shift = b + (bias-mean)*scale."""
    if bn is not None:
        inv = torch.rsqrt(bn.running_var.detach().double() + bn.eps)
        g = bn.weight.detach().double() if bn.weight is not None else torch.ones_like(inv)
        b = bn.bias.detach().double() if bn.bias is not None else torch.zeros_like(inv)
        scale = g * inv
        shift = b - bn.running_mean.detach().double() * scale
        if conv_bias is not None:
            shift = shift + conv_bias.detach().double() * scale
    else:
        scale = torch.ones(cout, dtype=torch.float64, device=device)
        shift = conv_bias.detach().double() if conv_bias is not None else torch.zeros(cout, dtype=torch.float64, device=device)
    return scale.float().contiguous(), shift.float().contiguous()


def invalidate(module):
    """Drops every cached lowering under `module`.  The caches key on tensor._version, which in-place
    edits through `.data` do not bump (load_state_dict / optimiser steps / copy_ do) — call this after
    such an edit."""
    for m in module.modules():
        for k in ("_pn_lowered", "_pn_group", "_pn_folded", "_pn_final_groups", "_pn_final_groups_tc",
                  "_pn_deconv_gemm"):
            m.__dict__.pop(k, None)


def split_weight_bf16x3(w2d, cin):
    """(Cout, taps*cin) f32 -> (Cout, k_pad) bf16 with every tap laid out [hi_w | hi_w | lo_w] (3*cin columns): the
    partner of ops.split_bf16x3's [hi_x | lo_x | hi_x] rows, so one bf16 GEMM sums hi*hi + lo*hi + hi*lo."""
    cout = w2d.shape[0]
    w = w2d.reshape(cout, -1, cin)
    hi = w.to(torch.bfloat16).float()
    lo = (w - hi).to(torch.bfloat16).float()
    return ops.pack_weight_bf16(torch.cat([hi, hi, lo], dim=2).reshape(cout, -1).contiguous())


class Lowered:
    __slots__ = ("weight", "k_pad", "scale", "shift", "key")


def lower(conv, bn, precision=None):
    """Cached (per module, per precision) packed weight + folded affine; refreshed when a parameter or
    running statistic changes (tensor._version), e.g. after load_state_dict or an optimiser step."""
    precision = precision or config.get_precision()
    bias = getattr(conv, "bias", None)
    key = (precision,) + _versions(conv.weight, bias,
                                   *(() if bn is None else (bn.weight, bn.bias, bn.running_mean, bn.running_var)))
    cache = conv.__dict__.setdefault("_pn_lowered", {})
    hit = cache.get(precision)
    if hit is not None and hit.key == key:
        return hit
    if bn is not None and bn.training:
        raise NotImplementedError("training-mode (batch statistics) BN is lowered by the training path only")
    w2d = weight_matrix(conv).float().contiguous()
    lw = Lowered()
    if precision == "bf16":
        lw.weight = ops.pack_weight_bf16(w2d)
    elif precision == "bf16x3":
        lw.weight = split_weight_bf16x3(w2d, conv.in_features if isinstance(conv, nn.Linear) else conv.in_channels)
    else:
        lw.weight = w2d
    lw.k_pad = lw.weight.shape[1]
    lw.scale, lw.shift = fold_affine(bias, bn, w2d.shape[0], w2d.device)
    lw.key = key
    cache[precision] = lw
    return lw


def lower_group(convs, bns, precision=None):
    """Horizontal fusion: several convs reading the same input become one conv with concatenated Cout."""
    precision = precision or config.get_precision()
    parts = [lower(c, b, precision) for c, b in zip(convs, bns)]
    key = tuple(p.key for p in parts)
    holder = convs[0].__dict__.setdefault("_pn_group", {})
    hit = holder.get(precision)
    if hit is not None and hit.key == key:
        return hit
    lw = Lowered()
    lw.weight = torch.cat([p.weight for p in parts], 0).contiguous()
    lw.k_pad = lw.weight.shape[1]
    lw.scale = torch.cat([p.scale for p in parts]).contiguous()
    lw.shift = torch.cat([p.shift for p in parts]).contiguous()
    lw.key = key
    holder[precision] = lw
    return lw


def run_conv(x2d, lw, nbr, taps, cin, cout, rows_cap, *, num=None, relu=False, residual=None, out=None,
             out_coff=0, out_dtype=None, in_ld=None, in_ptr_offset=0, rows_hint=0, out_hw_pad=None, deconv=None,
             nbr_kind=0, nbr_plan=None):
    """One fused conv launch on channels-last rows."""
    if out is None:
        out = torch.empty(rows_cap, cout, dtype=out_dtype or x2d.dtype, device=x2d.device)
    if config.get_precision() == "bf16x3":
        # fp32 rows in and out; the conv itself on the tensor cores over split-bf16 operands (three bf16 products per
        # fp32 product).  Rows are split for this launch only (6 bytes per value written, then gathered).
        if deconv is not None or x2d.dtype != torch.float32 or out.dtype != torch.float32:
            raise RuntimeError("bf16x3 mode runs fp32 rows through the gather conv")
        xs = ops.split_bf16x3(x2d, cin, col0=in_ptr_offset)
        ops.conv_gather(xs, lw.weight, nbr, taps, 3 * cin, cout, out, k_pad=lw.k_pad, scale=lw.scale, shift=lw.shift,
                        residual=residual, out_coff=out_coff, relu=relu, num=num, rows_cap=rows_cap,
                        impl=config.conv_impl(), rows_hint=rows_hint, out_hw_pad=out_hw_pad)
        return out
    ops.conv_gather(x2d, lw.weight, nbr, taps, cin, cout, out, in_ld=in_ld, k_pad=lw.k_pad, scale=lw.scale,
                    shift=lw.shift, residual=residual, out_coff=out_coff, relu=relu, num=num,
                    rows_cap=rows_cap, impl=config.conv_impl(), in_ptr_offset=in_ptr_offset, rows_hint=rows_hint,
                    out_hw_pad=out_hw_pad, deconv=deconv, nbr_kind=nbr_kind, nbr_plan=nbr_plan)
    return out


def use_padded_layout():
    """bf16 tensor-core mode stores dense BEV maps with a one-pixel zero border (see conv_dense_tc.cu)."""
    return config.get_precision() == "bf16"


class DenseMap:
    """Channels-last dense BEV map: rows (B*Hs*Ws, ld) with the logical channels at [coff, coff+C).
    pad = 1: storage is the zero-bordered map, Hs = H+2, Ws = W+2 (the layout pn_conv_dense3x3 uses)."""

    __slots__ = ("rows", "B", "H", "W", "C", "coff", "pad")

    def __init__(self, rows, B, H, W, C, coff=0, pad=0):
        self.rows, self.B, self.H, self.W, self.C, self.coff, self.pad = rows, B, H, W, C, coff, pad

    @property
    def n_rows(self):
        return self.B * (self.H + 2 * self.pad) * (self.W + 2 * self.pad)

    def nchw(self):
        """(B,C,H,W) view (channels_last strides), the reference's dense layout."""
        p = self.pad
        v = self.rows.view(self.B, self.H + 2 * p, self.W + 2 * p, -1)
        if p:
            v = v[:, 1:-1, 1:-1]
        v = v[..., self.coff:self.coff + self.C].permute(0, 3, 1, 2)
        v._pn_dense = self  # lets the next module recover the NHWC rows without a copy
        return v

    @staticmethod
    def from_nchw(t):
        d = getattr(t, "_pn_dense", None)
        if d is not None and (d.pad == 1) == use_padded_layout():
            return d
        B, C, H, W = t.shape
        x = t.permute(0, 2, 3, 1)
        if x.dtype != config.act_dtype():
            x = x.to(config.act_dtype())
        if use_padded_layout():
            x = torch.nn.functional.pad(x, (0, 0, 1, 1, 1, 1))
            return DenseMap(x.contiguous().view(B * (H + 2) * (W + 2), C), B, H, W, C, 0, 1)
        return DenseMap(x.contiguous().view(B * H * W, C), B, H, W, C)


def new_dense_rows(B, H, W, width, dtype, device, pad):
    return torch.empty(B * (H + 2 * pad) * (W + 2 * pad), width, dtype=dtype, device=device)


def dense_conv3x3(x, conv, bn, relu=True, stride=1, out=None, out_coff=0, out_dtype=None, out_compact=False,
                  lowered=None, planar_cols=0):
    """3x3 pad-1 dense conv (also ZeroPad2d(1)+valid conv, necks/rpn.py:172-176) on a DenseMap.
    Padded maps + stride 1 run on the TMA-fed dense tensor-core kernel; everything else on the gather conv."""
    Ho, Wo = (x.H + 2 - 3) // stride + 1, (x.W + 2 - 3) // stride + 1
    lw = lowered if lowered is not None else lower(conv, bn)
    cout = conv.out_channels
    opad = 0 if out_compact else x.pad
    tc_dense = x.pad and stride == 1 and x.C % 64 == 0 and x.coff % 8 == 0
    if planar_cols:
        # planar output (one contiguous padded map per `planar_cols` channels) for a following grouped conv
        if not (tc_dense and opad and out is None and cout % planar_cols == 0):
            raise RuntimeError("planar output needs the padded tensor-core path")
        n_rows = x.B * (Ho + 2) * (Wo + 2)
        out = torch.empty(cout // planar_cols * n_rows, planar_cols, dtype=out_dtype or x.rows.dtype,
                          device=x.rows.device)
        ops.conv_dense3x3(x.rows, x.coff, x.C, x.B, x.H, x.W, lw.weight, cout, out, scale=lw.scale, shift=lw.shift,
                          relu=relu, out_group_cols=planar_cols)
        return DenseMap(out, x.B, Ho, Wo, planar_cols, 0, opad)
    if out is None:
        out = new_dense_rows(x.B, Ho, Wo, cout, out_dtype or x.rows.dtype, x.rows.device, opad)
    if tc_dense:
        ops.conv_dense3x3(x.rows, x.coff, x.C, x.B, x.H, x.W, lw.weight, cout, out, scale=lw.scale, shift=lw.shift,
                          out_coff=out_coff, out_compact=out_compact, relu=relu)
    else:
        nbr = ops.dense_nbr_table(0, x.B, x.H, x.W, stride, x.rows.device, in_pad=bool(x.pad), out_pad=bool(opad))
        rows_cap = x.B * (Ho + 2 * opad) * (Wo + 2 * opad)
        run_conv(x.rows, lw, nbr, 9, x.C, cout, rows_cap, relu=relu, out=out, out_coff=out_coff,
                 in_ld=x.rows.stride(0), in_ptr_offset=x.coff,
                 out_hw_pad=(Ho + 2, Wo + 2) if opad else None)
    return DenseMap(out, x.B, Ho, Wo, cout, out_coff, opad)


def lower_deconv_gemm(conv, bn):
    """ConvTranspose2d(k=2,s=2) as one GEMM: weight rows (dy*2+dx)*Cout + o over K = Cin, affine repeated per tap."""
    base = lower(conv, bn)
    cache = conv.__dict__.setdefault("_pn_deconv_gemm", {})
    hit = cache.get("bf16")
    if hit is not None and hit.key == base.key:
        return hit
    cout, cin = conv.out_channels, conv.in_channels
    w = weight_matrix(conv).float().reshape(cout, 4, cin).permute(1, 0, 2).reshape(4 * cout, cin).contiguous()
    lw = Lowered()
    lw.weight = ops.pack_weight_bf16(w)
    lw.k_pad = lw.weight.shape[1]
    lw.scale, lw.shift = base.scale.repeat(4).contiguous(), base.shift.repeat(4).contiguous()
    lw.key = base.key
    cache["bf16"] = lw
    return lw


def dense_deconv2x2(x, conv, bn, relu=True, out=None, out_coff=0):
    """ConvTranspose2d(k=2,s=2)+BN+ReLU (necks/rpn.py:150-154).  On padded bf16 maps: one GEMM over the input
    pixels whose epilogue scatters each tap's 2x-upsampled position (pn_conv_args.deconv_*); otherwise a 4-tap
    gather conv over the output pixels (three of the four taps of every output pixel are empty)."""
    cout = conv.out_channels
    Ho, Wo = 2 * x.H, 2 * x.W
    if out is None:
        out = new_dense_rows(x.B, Ho, Wo, cout, x.rows.dtype, x.rows.device, x.pad)
    if (x.pad and config.get_precision() == "bf16" and config.deconv_gemm() and x.C % 64 == 0 and cout % 16 == 0
            and x.coff % 8 == 0):
        lw = lower_deconv_gemm(conv, bn)
        rows_in = x.B * (x.H + 2) * (x.W + 2)
        run_conv(x.rows, lw, None, 1, x.C, 4 * cout, rows_in, relu=relu, out=out, out_coff=out_coff,
                 in_ld=x.rows.stride(0), in_ptr_offset=x.coff, out_hw_pad=(Ho + 2, Wo + 2),
                 deconv=(cout, x.H + 2, x.W + 2))
        return DenseMap(out, x.B, Ho, Wo, cout, out_coff, x.pad)
    nbr = ops.dense_nbr_table(1, x.B, x.H, x.W, 2, x.rows.device, in_pad=bool(x.pad), out_pad=bool(x.pad))
    lw = lower(conv, bn)
    rows_cap = x.B * (Ho + 2 * x.pad) * (Wo + 2 * x.pad)
    run_conv(x.rows, lw, nbr, 4, x.C, cout, rows_cap, relu=relu, out=out, out_coff=out_coff,
             in_ld=x.rows.stride(0), in_ptr_offset=x.coff, out_hw_pad=(Ho + 2, Wo + 2) if x.pad else None)
    return DenseMap(out, x.B, Ho, Wo, cout, out_coff, x.pad)
