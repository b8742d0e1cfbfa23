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


# ---- synthetic supervision + one optimisation step ---------------------------------------------------

def synthetic_targets(head, n_frames, bev_h, bev_w, rng, max_objs=500, device="cuda"):
    """Random CenterPoint-style targets with the shapes AssignLabel produces (datasets/pipelines/preprocess.py:
    249-345): per task hm (B,H,W,K) with unit peaks at the object centres, ind/mask/cat (B,M), anno_box (B,M,10)
    [reg2, height, log-dim3, vel2, sin, cos], gt_box (B,M,7).  Benchmark / test input only."""
    import numpy as np
    out = {k: [] for k in ("hm", "ind", "mask", "cat", "anno_box", "gt_box")}
    for t, K in enumerate(head.num_classes):
        s = head.task_strides[t]
        H, W = bev_h // s, bev_w // s
        M = max_objs
        hm = np.zeros((n_frames, H, W, K), np.float32)
        ind = np.zeros((n_frames, M), np.int64)
        mask = np.zeros((n_frames, M), np.uint8)
        cat = np.zeros((n_frames, M), np.int64)
        anno = np.zeros((n_frames, M, 10), np.float32)
        gt = np.zeros((n_frames, M, 7), np.float32)
        for b in range(n_frames):
            n = int(rng.integers(max(1, M // 8), max(2, M // 2)))
            pix = rng.choice(H * W, n, replace=False)
            ind[b, :n], mask[b, :n] = pix, 1
            cat[b, :n] = rng.integers(0, K, n)
            ys, xs = pix // W, pix % W
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    yy, xx = np.clip(ys + dy, 0, H - 1), np.clip(xs + dx, 0, W - 1)
                    v = 1.0 if (dy == 0 and dx == 0) else 0.4
                    np.maximum.at(hm, (b, yy, xx, cat[b, :n]), v)
            reg = rng.uniform(0, 1, (n, 2)).astype(np.float32)
            dim = rng.uniform(0.5, 5.0, (n, 3)).astype(np.float32)
            rot = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
            z = rng.uniform(-2, 1, n).astype(np.float32)
            anno[b, :n, 0:2], anno[b, :n, 2], anno[b, :n, 3:6] = reg, z, np.log(dim)
            anno[b, :n, 6:8] = rng.normal(0, 1, (n, 2))
            anno[b, :n, 8], anno[b, :n, 9] = np.sin(rot), np.cos(rot)
            gt[b, :n, 0] = (xs + reg[:, 0]) * s * head.pillar_size + head.point_cloud_range[0]
            gt[b, :n, 1] = (ys + reg[:, 1]) * s * head.pillar_size + head.point_cloud_range[1]
            gt[b, :n, 2], gt[b, :n, 3:6], gt[b, :n, 6] = z, dim, rot
        for k, v in (("hm", hm), ("ind", ind), ("mask", mask), ("cat", cat), ("anno_box", anno), ("gt_box", gt)):
            out[k].append(torch.from_numpy(v).to(device))
    return out


def train_step(model, example, optimizer, averager=None):
    """forward + loss + backward (+ overlapped gradient all-reduce, dist.GradientAverager) + optimiser step;
    returns the summed loss (device scalar)."""
    if averager is not None:
        averager.zero_grad()
    else:
        optimizer.zero_grad(set_to_none=True)
    losses = model(example, return_loss=True)
    loss = sum(l.sum() for l in losses["loss"])
    loss.backward()
    if averager is not None:
        averager.finish()
    optimizer.step()
    return loss.detach()
