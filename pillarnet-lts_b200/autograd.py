"""Autograd bindings of the training-path kernels.

`ScatterMaxFunction` mirrors det3d/ops/pillar_ops/scatter_utils.py:7-37 (same forward/backward contract, same
flat `arg` convention).  `SparseConvFunction` is the autograd node spconv supplies for SubMConv2d /
SparseConv2d in the reference (external; used by backbones/base.py:38-63, PillarResNet.py:87,95,103):
forward and data-gradient are the gather-GEMM kernel (pn_conv_gather) on the output- and input-stationary
rulebook, the weight gradient is pn_conv_wgrad.  No torch fallback: every path calls the C ABI.
"""
import torch
from torch.autograd import Function

from . import config, ops


class ScatterMaxFunction(Function):
    @staticmethod
    def forward(ctx, src, index, M):
        """src (L,C) f32, index (L,) int32 -> out (M,C) f32: max(0, max over the pillar's points)."""
        out, arg = ops.scatter_max(src.contiguous(), index.contiguous(), int(M), want_arg=True)
        ctx.for_backwards = (src.shape[0], src.shape[1], arg)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        L, C, arg = ctx.for_backwards
        return ops.scatter_max_grad_flat(grad_out, arg, L), None, None


scatter_max = ScatterMaxFunction.apply


class Rulebook:
    """Exact-size (training) rulebook: nbr (n_out,taps) output-stationary, nbr_t (n_in,taps) input-stationary."""

    __slots__ = ("nbr", "n_in", "n_out", "taps", "_nbr_t")

    def __init__(self, nbr, n_in, n_out, taps=9):
        self.nbr, self.n_in, self.n_out, self.taps = nbr, n_in, n_out, taps
        self._nbr_t = None

    @property
    def nbr_t(self):
        if self._nbr_t is None:
            self._nbr_t = ops.rulebook_transpose(self.nbr[:self.n_out], self.n_in)
        return self._nbr_t


def _pack(w2d):
    """weight matrix in the dtype/layout the active conv implementation wants"""
    w2d = w2d.float().contiguous()
    return ops.pack_weight_bf16(w2d) if config.get_precision() == "bf16" else w2d


class SparseConvFunction(Function):
    """y[o] = bias + sum_t W_t x[nbr[o,t]]   (weight (Cout,kH,kW,Cin) = spconv 2.x layout)."""

    @staticmethod
    def forward(ctx, x, weight, bias, rb):
        cout, cin = weight.shape[0], weight.shape[-1]
        taps = rb.taps
        act = config.act_dtype()
        xq = x.detach().to(act).contiguous()
        wq = _pack(weight.detach().reshape(cout, taps * cin))
        out = torch.empty(rb.n_out, cout, dtype=act, device=x.device)
        if rb.n_out:
            ops.conv_gather(xq, wq, rb.nbr, taps, cin, cout, out, k_pad=wq.shape[1],
                            shift=bias.detach().float().contiguous() if bias is not None else None,
                            rows_cap=rb.n_out, impl=config.conv_impl())
        ctx.save_for_backward(xq, weight)
        ctx.rb, ctx.has_bias, ctx.x_dtype = rb, bias is not None, x.dtype
        return out

    @staticmethod
    def backward(ctx, dy):
        xq, weight = ctx.saved_tensors
        rb = ctx.rb
        cout, cin = weight.shape[0], weight.shape[-1]
        taps = rb.taps
        act = config.act_dtype()
        dyq = dy.detach().to(act).contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # W^T[ci][t*cout + co] = W[co][t][ci]
            wt = _pack(weight.detach().reshape(cout, taps, cin).permute(2, 1, 0).reshape(cin, taps * cout))
            dx = torch.empty(rb.n_in, cin, dtype=act, device=dy.device)
            if rb.n_in:
                ops.conv_gather(dyq, wt, rb.nbr_t, taps, cout, cin, dx, k_pad=wt.shape[1], rows_cap=rb.n_in,
                                impl=config.conv_impl())
            dx = dx.to(ctx.x_dtype)
        if ctx.needs_input_grad[1]:
            dw = ops.conv_wgrad(xq, dyq, rb.nbr, taps, cin, cout, rows=rb.n_out, impl=config.conv_impl())
            dw = dw.view(weight.shape).to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dyq.float().sum(0)
        return dx, dw, db, None


sparse_conv = SparseConvFunction.apply
