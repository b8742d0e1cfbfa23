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


# ---- sync-free (CUDA-graph capturable) training nodes -----------------------------------------------------------------

class StaticRulebook:
    """Capacity-sized rulebook with device-resident row counts: nbr (cap_out, taps) output-stationary, nbr_t
    (cap_in, taps) input-stationary; plan / plan_t = tile plans of the window-staged kernel (bf16, submanifold only)."""

    __slots__ = ("nbr", "nbr_t", "num_in", "num_out", "cap_in", "cap_out", "taps", "plan", "plan_t", "kind")

    def __init__(self, nbr, nbr_t, num_in, num_out, cap_in, cap_out, plan=None, plan_t=None, kind=0, taps=9):
        self.nbr, self.nbr_t, self.num_in, self.num_out = nbr, nbr_t, num_in, num_out
        self.cap_in, self.cap_out, self.taps, self.plan, self.plan_t, self.kind = cap_in, cap_out, taps, plan, plan_t, kind


class SparseConvBNFunction(Function):
    """y = act(BN_batch(conv(x) + bias) + residual) on capacity-sized rows with the live counts on the device:
    SparseSequential(SubMConv2d | SparseConv2d, BN1d[, SparseReLU]) (+ the block's residual add and ReLU) of
    backbones/base.py:145-213 in train mode as ONE autograd node — conv (pn_conv_gather), batch statistics, normalise
    + residual + ReLU (pn_bn_*), and their backward: BN backward, data gradient (gather conv on the input-stationary
    rulebook), weight gradient (pn_conv_wgrad).  No host synchronisation anywhere."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, residual, rb, bn, relu):
        cout, cin = weight.shape[0], weight.shape[-1]
        act = config.act_dtype()
        xq = x.detach().to(act).contiguous()
        wq = _pack(weight.detach().reshape(cout, rb.taps * cin))
        xc = torch.empty(rb.cap_out, cout, dtype=act, device=x.device)
        ops.conv_gather(xq, wq, rb.nbr, rb.taps, cin, cout, xc, k_pad=wq.shape[1],
                        shift=bias.detach().float().contiguous() if bias is not None else None, num=rb.num_out,
                        rows_cap=rb.cap_out, impl=config.conv_impl(), nbr_kind=rb.kind, nbr_plan=rb.plan)
        res = residual.detach().to(act).contiguous() if residual is not None else None
        y, mean, rstd = ops.bn_train_forward(xc, rb.num_out, gamma.detach().float(), beta.detach().float(),
                                             bn.running_mean, bn.running_var, bn.eps, bn.momentum, res, relu)
        if bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        ctx.save_for_backward(xq, weight, xc, y, mean, rstd, gamma)
        ctx.rb, ctx.relu, ctx.has_bias, ctx.has_res = rb, relu, bias is not None, residual is not None
        ctx.x_dtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        xq, weight, xc, y, mean, rstd, gamma = ctx.saved_tensors
        rb = ctx.rb
        cout, cin = weight.shape[0], weight.shape[-1]
        act = config.act_dtype()
        dconv, dres, dgamma, dbeta = ops.bn_train_backward(dy, y, xc, mean, rstd, gamma.detach().float(), ctx.relu,
                                                           rb.num_out, ctx.has_res)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            wt = _pack(weight.detach().reshape(cout, rb.taps, cin).permute(2, 1, 0).reshape(cin, rb.taps * cout))
            dx = torch.empty(rb.cap_in, cin, dtype=act, device=dy.device)
            ops.conv_gather(dconv, wt, rb.nbr_t, rb.taps, cout, cin, dx, k_pad=wt.shape[1], num=rb.num_in,
                            rows_cap=rb.cap_in, impl=config.conv_impl(), nbr_kind=rb.kind, nbr_plan=rb.plan_t)
            dx = dx.to(ctx.x_dtype)
        if ctx.needs_input_grad[1]:
            dw = ops.conv_wgrad(xq, dconv, rb.nbr, rb.taps, cin, cout, num=rb.num_out, rows=rb.cap_out,
                                impl=config.conv_impl()).view(weight.shape).to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            # a bias in front of a batch-statistics BN has no gradient: sum_r dconv[r] = -gamma*rstd*(sum_r xhat)*(..)/n
            # and sum_r xhat = 0 (the dynamic path's dy.sum(0) is that zero plus rounding noise)
            db = torch.zeros(cout, dtype=weight.dtype, device=dy.device)
        return dx, dw, db, dgamma.to(gamma.dtype), dbeta.to(gamma.dtype), dres, None, None, None


sparse_conv_bn = SparseConvBNFunction.apply


class DenseFromSparseFunction(Function):
    """SparseConvTensor.dense() (PillarResNet.py:139) with a device-resident row count: forward pn_sparse_to_dense,
    backward pn_dense_to_sparse."""

    @staticmethod
    def forward(ctx, feat, table):
        C = feat.shape[1]
        rows = ops.sparse_to_dense(feat.detach().contiguous(), table, C)
        ctx.table, ctx.C = table, C
        return rows.view(table.B, table.H, table.W, C).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        t = ctx.table
        rows = g.permute(0, 2, 3, 1).contiguous().view(t.B * t.H * t.W, ctx.C)
        return ops.dense_to_sparse(rows, t, ctx.C), None


dense_from_sparse_static = DenseFromSparseFunction.apply


class DenseBNFunction(Function):
    """nn.BatchNorm2d in train mode (+ the ReLU behind it) on a channels-last (B, C, H, W) map as one autograd node on the
    row kernels of csrc/bn_train.cu (a dense NHWC map is a (B*H*W, C) row matrix): statistics + finalize (running stats
    updated in place like torch) + apply, and the two-kernel backward.  Replaces PyTorch's four channels-last BN kernels
    + a separate ReLU per layer of the dense conv5 / neck / head in the static training path (det3d/models/necks/rpn.py:
    172-185, center_head.py:27-33 in train mode)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, bn, relu, conv_bias=None):
        """conv_bias: bias of the conv that produced x, NOT yet added (batch normalisation removes a per-channel shift
        exactly, so the add — and the (B, H, W) reduction of its gradient — are skipped; only the running mean sees
        it).  Its gradient is exactly zero."""
        B, C, H, W = x.shape
        # channels-last maps and channel slices of them flatten to a (B*H*W, C) row view (row stride = all channels of
        # the parent) without a copy; anything else is copied
        rows = x.permute(0, 2, 3, 1).reshape(B * H * W, C)
        if rows.stride(1) != 1 or rows.stride(0) % 8 != 0 or rows.data_ptr() % 16 != 0:
            rows = rows.contiguous()
        y, mean, rstd = ops.bn_train_forward(rows, None, gamma.detach().float(), beta.detach().float(), bn.running_mean,
                                             bn.running_var, bn.eps, bn.momentum, None, relu)
        if bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        if conv_bias is not None:
            bn.running_mean.add_(conv_bias.detach().to(bn.running_mean.dtype), alpha=bn.momentum)
        ctx.save_for_backward(rows, y if relu else None, mean, rstd, gamma)
        ctx.relu, ctx.shape = relu, (B, C, H, W)
        ctx.bias_like = conv_bias
        return y.view(B, H, W, C).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        rows, y, mean, rstd, gamma = ctx.saved_tensors
        B, C, H, W = ctx.shape
        dr = dy.permute(0, 2, 3, 1)
        if not dr.is_contiguous():
            dr = dr.contiguous()
        dx, _, dgamma, dbeta = ops.bn_train_backward(dr.view(B * H * W, C), y, rows, mean, rstd, gamma.detach().float(),
                                                     ctx.relu, None, False)
        dbias = torch.zeros_like(ctx.bias_like) if ctx.bias_like is not None else None
        return (dx.view(B, H, W, C).permute(0, 3, 1, 2), dgamma.to(gamma.dtype), dbeta.to(gamma.dtype), None, None, dbias)


class RowBNFunction(Function):
    """nn.BatchNorm1d in train mode (+ ReLU) over the first *num rows of a (rows, C) matrix (device-resident count) on the
    library's BN kernels — the PFN's BatchNorm over the in-range points once they are compacted to a prefix
    (det3d/models/readers/pillar_modules.py:26-33 in train mode).  Rows past the count are left as they are in the output
    buffer (zeros) and get zero gradient."""

    @staticmethod
    def forward(ctx, x, gamma, beta, bn, relu, num):
        x = x.contiguous()
        y, mean, rstd = ops.bn_train_forward(x, num, gamma.detach().float(), beta.detach().float(), bn.running_mean,
                                             bn.running_var, bn.eps, bn.momentum, None, relu)
        if bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        ctx.save_for_backward(x, y if relu else None, mean, rstd, gamma, num)
        ctx.relu = relu
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, mean, rstd, gamma, num = ctx.saved_tensors
        dx, _, dgamma, dbeta = ops.bn_train_backward(dy, y, x, mean, rstd, gamma.detach().float(), ctx.relu, num, False,
                                                     zero_tail=True)      # the Linear's backward reads every row
        return dx, dgamma.to(gamma.dtype), dbeta.to(gamma.dtype), None, None, None
