"""Training-mode forward of the hot path: reader and sparse backbone run on the library's
kernels through autograd nodes (autograd.py); the dense conv5 / neck / head and the loss stay PyTorch (cuDNN),
as §8 a25 scopes them ("loss stays PyTorch").  BatchNorm uses batch statistics over the active rows
(nn.BatchNorm1d on the (M,C) feature matrix, backbones/base.py:155-213 in train mode).

Training synchronises once per rulebook (exact row counts are read back), like the reference does
(`.item()` at pillar_utils.py:45 and inside spconv); inference stays sync-free.
"""
import contextlib

import torch
import torch.nn.functional as F

from . import config, ops
from .autograd import Rulebook, scatter_max, sparse_conv
from .sparse import SparseConvTensor


def _exact(table):
    """RankTable view with cap == live rows (one host sync, cached on the table)."""
    ex = table._train.get("exact") if table._train is not None else None
    if ex is None:
        n = table.count()
        ex = ops.RankTable(table.words, table.prefix, table.coords[:n], table.num, n, table.B, table.H, table.W)
        ex._train = {"exact": ex}
        if table._train is None:
            table._train = {}
        table._train["exact"] = ex
    return ex


def subm_rulebook(table):
    """cached per table, like spconv's indice_key (backbones/base.py:43-52)"""
    rb = table._train.get("subm")
    if rb is None:
        nbr = ops.rulebook_subm3x3(table) if table.cap else torch.empty(0, 9, dtype=torch.int32,
                                                                         device=table.coords.device)
        rb = Rulebook(nbr, table.cap, table.cap)
        # a submanifold rulebook is its own transpose with the taps mirrored: nbr_t[i,t] = nbr[i,8-t]
        rb._nbr_t = nbr.flip(1).contiguous()
        table._train["subm"] = rb
    return rb


def down_rulebook(table):
    out_table, nbr = ops.rulebook_down3x3s2(table)
    ex = _exact(out_table)
    return ex, Rulebook(nbr[:ex.cap].contiguous(), table.cap, ex.cap)


def autocast_ctx():
    if config.get_precision() == "bf16":
        return torch.autocast(device_type="cuda", dtype=torch.bfloat16)
    return contextlib.nullcontext()


# ---- reader ---------------------------------------------------------------------------------------

def reader_forward(pfn, points, frame_offsets, batch_size):
    """PillarMaxPooling.forward in train mode (pillar_modules.py:56-74): the MLP is torch (autograd),
    pillarization / point features / scatter-max (+ its backward) are library kernels."""
    table, point_pillar = ops.pillarize(points, frame_offsets, batch_size, pfn.height, pfn.width,
                                        pfn.point_cloud_range[0], pfn.point_cloud_range[1], pfn.pillar_size)
    ex = _exact(table)
    sel = torch.nonzero(point_pillar >= 0).squeeze(1)          # in-range points, input order (reference order)
    pts = points.index_select(0, sel).contiguous()
    idx = point_pillar.index_select(0, sel).contiguous()
    feats = ops.point_features(pts, pfn.point_cloud_range[0], pfn.point_cloud_range[1], pfn.pillar_size,
                               pfn.x_offset, pfn.y_offset)
    h = pfn.shared_mlps(feats)                                  # Linear -> BN1d(batch stats) -> ReLU
    pooled = scatter_max(h.float(), idx, ex.cap)
    sp = SparseConvTensor(pooled.to(config.act_dtype()), ex, (pfn.height, pfn.width), batch_size)
    sp._count = ex.cap
    sp.point_pillar = point_pillar
    return sp


# ---- sparse backbone ----------------------------------------------------------------------------------

def conv_bn(feat, conv, bn, rb, relu, residual=None):
    y = sparse_conv(feat, conv.weight, conv.bias, rb)
    y = bn(y)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y


def subm_block(sp, seq, relu, residual=None):
    rb = subm_rulebook(sp.table)
    out = conv_bn(sp.feat, seq[0], seq[1], rb, relu, residual)
    return _wrap(out, sp.table, sp)


def _wrap(feat, table, like):
    sp = SparseConvTensor(feat, table, (table.H, table.W), like.batch_size)
    sp._count = table.cap
    return sp


def down_block(sp, conv, bn):
    out_table, rb = down_rulebook(sp.table)
    out = conv_bn(sp.feat, conv, bn, rb, relu=True)
    return _wrap(out, out_table, sp)


def dense_from_sparse(sp):
    """differentiable SparseConvTensor.dense() (PillarResNet.py:139): (B,C,H,W)"""
    t = sp.table
    c = t.coords.long()
    d = sp.feat.new_zeros(t.B, t.H, t.W, sp.feat.shape[1])
    d = d.index_put((c[:, 0], c[:, 1], c[:, 2]), sp.feat)
    return d.permute(0, 3, 1, 2)


def to_dense(x):
    return dense_from_sparse(x) if isinstance(x, SparseConvTensor) else x
