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
from ._lib import PN_NBR_SUBM_SORTED
from .autograd import (Rulebook, StaticRulebook, dense_from_sparse_static, scatter_max, sparse_conv, sparse_conv_bn)
from .sparse import SparseConvTensor

# Static mode: every tensor of the step has a capacity-derived shape and the live row counts stay on the device, so the
# whole step (forward, loss, backward, optimiser) is free of host synchronisation and can be captured in a CUDA graph
# (TrainEngine).  The dynamic mode below sizes every matrix exactly and reads the counts back, like the reference.
_static = False


def set_static(on):
    global _static
    _static = bool(on)


def is_static():
    return _static


def run_dense_seq(seq, x):
    """A Sequential of the dense conv5 / neck / head in TRAIN mode: convs (and anything else) run as PyTorch modules under
    autograd (cuDNN); in the static path every BatchNorm2d and the ReLU behind it run as one node on the library's BN
    kernels (autograd.DenseBNFunction) when the map is bf16 / f32 with C % 8 == 0."""
    from torch import nn
    from .autograd import DenseBNFunction
    mods = list(seq)

    def fusable_bn(m, x):
        return (_static and isinstance(m, nn.BatchNorm2d) and m.training and m.track_running_stats and m.affine
                and m.momentum is not None and x.is_cuda and x.shape[1] % 8 == 0)

    i = 0
    while i < len(mods):
        m = mods[i]
        if (isinstance(m, nn.Conv2d) and m.bias is not None and m.padding_mode == "zeros" and i + 1 < len(mods)
                and isinstance(mods[i + 1], nn.BatchNorm2d) and m.out_channels % 8 == 0 and fusable_bn(mods[i + 1], x)):
            # conv + bias followed by batch statistics: the bias add and the reduction of its gradient are dropped
            bn = mods[i + 1]
            y = torch.nn.functional.conv2d(x, m.weight, None, m.stride, m.padding, m.dilation, m.groups)
            relu = i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
            x = DenseBNFunction.apply(y, bn.weight, bn.bias, bn, relu, m.bias)
            i += 3 if relu else 2
        elif (_static and isinstance(m, nn.BatchNorm2d) and m.training and m.track_running_stats and m.affine
                and m.momentum is not None and x.is_cuda and x.dtype in (torch.bfloat16, torch.float32)
                and x.shape[1] % 8 == 0):
            relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
            x = DenseBNFunction.apply(x, m.weight, m.bias, m, relu)
            i += 2 if relu else 1
        elif isinstance(m, nn.Sequential):
            x = run_dense_seq(m, x)
            i += 1
        else:
            x = m(x)
            i += 1
    return x


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

def _masked_bn_relu(h, valid, bn):
    """BatchNorm1d (batch statistics over the rows with valid == 1, running stats updated as torch does) + ReLU, written
    with shape-static torch ops so it needs no compaction of the valid rows."""
    w = valid.to(h.dtype).unsqueeze(1)
    n = w.sum().clamp(min=1.0)
    mean = (h * w).sum(0) / n
    d = (h - mean) * w
    var = (d * d).sum(0) / n
    y = (h - mean) * torch.rsqrt(var + bn.eps) * bn.weight + bn.bias
    with torch.no_grad():
        m = bn.momentum
        bn.running_mean.mul_(1 - m).add_(mean.detach() * m)
        bn.running_var.mul_(1 - m).add_(var.detach() * (n / (n - 1).clamp(min=1.0)) * m)
        bn.num_batches_tracked.add_(1)
    return F.relu(y)


def _reader_forward_static(pfn, points, frame_offsets, batch_size):
    table, point_pillar = ops.pillarize(points, frame_offsets, batch_size, pfn.height, pfn.width,
                                        pfn.point_cloud_range[0], pfn.point_cloud_range[1], pfn.pillar_size)
    feats = ops.point_features(points, pfn.point_cloud_range[0], pfn.point_cloud_range[1], pfn.pillar_size,
                               pfn.x_offset, pfn.y_offset)
    # points past the live count (the tail of the capacity) were not visited by pn_pillarize: index -1 like the
    # out-of-range ones
    live = frame_offsets[batch_size:batch_size + 1]
    point_pillar = torch.where(torch.arange(points.shape[0], device=points.device, dtype=torch.int32) < live,
                               point_pillar, torch.full_like(point_pillar, -1))
    bn = pfn.shared_mlps[1]
    if isinstance(bn, torch.nn.BatchNorm1d) and bn.track_running_stats and bn.affine and bn.momentum is not None:
        # The in-range points are moved to a prefix (stable; the others fill the tail) with shape-static index math, so
        # the BatchNorm + ReLU over them is the library's fused row kernel with a device-resident count instead of ~40
        # masked elementwise / reduction passes over the (capacity, 32) matrix.  The pooling is order independent.
        from .autograd import RowBNFunction
        valid = point_pillar >= 0
        n_pts = points.shape[0]
        c_valid = torch.cumsum(valid, 0, dtype=torch.int32)
        live_valid = c_valid[-1:].contiguous()                      # (1,) int32 on the device
        pos = torch.arange(1, n_pts + 1, device=points.device, dtype=torch.int32)
        dest = torch.where(valid, c_valid - 1, n_pts - (pos - c_valid)).long()     # a permutation of 0..n_pts-1
        feats_c = torch.empty_like(feats).index_copy_(0, dest, feats)
        pp_c = torch.empty_like(point_pillar).index_copy_(0, dest, point_pillar)
        h = pfn.shared_mlps[0](feats_c)                             # Linear on every point of the capacity
        h = RowBNFunction.apply(h.float(), bn.weight, bn.bias, bn, True, live_valid)
        pooled = scatter_max(h, pp_c, table.cap)                    # index -1 (the tail) is skipped
    else:
        h = pfn.shared_mlps[0](feats)                               # Linear on every point of the capacity
        h = _masked_bn_relu(h.float(), point_pillar >= 0, bn)
        pooled = scatter_max(h.contiguous(), point_pillar, table.cap)   # index -1 is skipped
    sp = SparseConvTensor(pooled.to(config.act_dtype()), table, (pfn.height, pfn.width), batch_size)
    sp._count = table.cap
    sp.point_pillar = point_pillar
    return sp


def reader_forward(pfn, points, frame_offsets, batch_size):
    """PillarMaxPooling.forward in train mode (pillar_modules.py:56-74): the MLP is torch (autograd),
    pillarization / point features / scatter-max (+ its backward) are library kernels."""
    if _static:
        return _reader_forward_static(pfn, points, frame_offsets, batch_size)
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


def _subm_rulebook_static(table):
    if table._train is None:
        table._train = {}
    rb = table._train.get("subm_static")
    if rb is None:
        nbr = table.subm_nbr()
        # a submanifold rulebook is its own transpose with the taps mirrored: nbr_t[i,t] = nbr[i,8-t]
        nbr_t = nbr.flip(1).contiguous()
        bf16 = config.get_precision() == "bf16"
        plan = table.subm_plan() if bf16 else None
        plan_t = ops.conv_window_plan(nbr_t, table.num, table.cap) if bf16 else None
        rb = StaticRulebook(nbr, nbr_t, table.num, table.num, table.cap, table.cap, plan, plan_t, PN_NBR_SUBM_SORTED)
        table._train["subm_static"] = rb
    return rb


def _down_rulebook_static(table):
    out_table, nbr = ops.rulebook_down3x3s2(table)
    nbr_t = ops.rulebook_transpose(nbr, table.cap, out_table.num)
    return out_table, StaticRulebook(nbr, nbr_t, table.num, out_table.num, table.cap, out_table.cap)


def subm_block(sp, seq, relu, residual=None):
    if _static:
        rb = _subm_rulebook_static(sp.table)
        conv, bn = seq[0], seq[1]
        out = sparse_conv_bn(sp.feat, conv.weight, conv.bias, bn.weight, bn.bias, residual, rb, bn, relu)
        return _wrap(out, sp.table, sp)
    rb = subm_rulebook(sp.table)
    out = conv_bn(sp.feat, seq[0], seq[1], rb, relu, residual)
    return _wrap(out, sp.table, sp)


def _wrap(feat, table, like):
    sp = SparseConvTensor(feat, table, (table.H, table.W), like.batch_size)
    sp._count = table.cap
    return sp


def down_block(sp, conv, bn):
    if _static:
        out_table, rb = _down_rulebook_static(sp.table)
        out = sparse_conv_bn(sp.feat, conv.weight, conv.bias, bn.weight, bn.bias, None, rb, bn, True)
        return _wrap(out, out_table, sp)
    out_table, rb = down_rulebook(sp.table)
    out = conv_bn(sp.feat, conv, bn, rb, relu=True)
    return _wrap(out, out_table, sp)


def dense_from_sparse(sp):
    """differentiable SparseConvTensor.dense() (PillarResNet.py:139): (B,C,H,W)"""
    if _static:
        return dense_from_sparse_static(sp.feat, sp.table)
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


class TrainEngine:
    """The whole training step of BASELINE config 4 without host synchronisation, as two CUDA graphs:

        graph A: zero the gradients, forward (reader, sparse backbone, dense conv5 / neck / head), loss, backward
        (data-parallel runs: one bucketed NCCL all-reduce of the flat gradient buffers, dist.GradientAverager)
        graph B: the optimiser step (a `capturable=True` torch optimiser)

    Inputs live in fixed buffers sized by a capacity (points, frame offsets, the target tensors of
    CenterHead.loss); `step(example)` copies a batch in and replays.  Everything in between uses capacity-derived
    shapes with the live counts on the device (set_static above).  The reference synchronises the host several times
    per layer (`.item()` at pillar_utils.py:45, spconv's indice-pair counts, `mask.sum() == 0` in the losses) and
    issues ~5000 launches per step from Python.

        eng = TrainEngine(model, optimizer, n_frames=4, points_cap=1_200_000, example=first_batch).prepare()
        loss = eng.step(batch)        # device scalar
    """

    TARGET_KEYS = ("hm", "ind", "mask", "cat", "anno_box", "gt_box")

    def __init__(self, model, optimizer, n_frames, points_cap, example, averager=None, point_dim=5, use_graph=True):
        self.model, self.opt, self.avg = model.train(), optimizer, averager
        self.B, self.cap = n_frames, int(points_cap)
        self.dev = next(model.parameters()).device
        self.points = torch.zeros(self.cap, point_dim, dtype=torch.float32, device=self.dev)
        self.offsets = torch.zeros(n_frames + 1, dtype=torch.int32, device=self.dev)
        self.targets = {k: [t.clone() for t in example[k]] for k in self.TARGET_KEYS if k in example}
        self.stream = torch.cuda.Stream(device=self.dev)
        self.use_graph = use_graph
        self.graph_fb = self.graph_opt = None
        self.loss = None
        self.load(example)

    def load(self, example):
        """copies one batch into the fixed input buffers (async on the engine stream)"""
        pts, off = example["points_batched"]
        n = pts.shape[0]
        if n > self.cap:
            raise RuntimeError(f"{n} points exceed the engine capacity {self.cap}")
        with torch.cuda.stream(self.stream):
            self.points[:n].copy_(pts, non_blocking=True)
            self.offsets.copy_(off, non_blocking=True)
            for k, lst in self.targets.items():
                for dst, src in zip(lst, example[k]):
                    dst.copy_(src, non_blocking=True)

    def _example(self):
        ex = {"points_batched": (self.points, self.offsets), "points": None, "metadata": [None] * self.B}
        ex.update(self.targets)
        return ex

    def _forward_backward(self):
        if self.avg is not None:
            self.avg.zero_grad()
        else:
            # None, not zeros: autograd then writes each gradient instead of accumulating into it (saves a fill and an
            # add per parameter, ~1200 launches); under capture the new gradient tensors come from the graph's private
            # pool at fixed addresses, which the optimiser graph captured right after sees
            self.opt.zero_grad(set_to_none=True)
        losses = self.model(self._example(), return_loss=True)
        loss = sum(l.sum() for l in losses["loss"])
        loss.backward()
        return loss.detach()

    def prepare(self, warmup=3):
        """eager warm-up steps (allocator, lowering caches, optimiser state) on the engine stream, then capture"""
        set_static(True)
        if self.avg is not None:
            self.avg.defer = True          # gradients are reduced after the captured backward, in bucket order
        with torch.cuda.stream(self.stream):
            for _ in range(max(1, warmup)):
                self.loss = self._forward_backward()
                if self.avg is not None:
                    self.avg.finish()
                self.opt.step()
            self.stream.synchronize()
            from . import _lib
            l0 = _lib.load().pn_launch_count()
            self.loss = self._forward_backward()       # one more eager pass, counted: the library kernels of a step
            if self.avg is not None:
                self.avg.finish()
            self.opt.step()
            self.stream.synchronize()
            self.launches_per_step = int(_lib.load().pn_launch_count() - l0)
            if self.use_graph:
                self.graph_fb = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_fb, stream=self.stream):
                    self.loss = self._forward_backward()
                self.graph_opt = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_opt, stream=self.stream):
                    self.opt.step()
        self.stream.synchronize()
        return self

    def step(self, example=None):
        """one training step; returns the summed loss (device scalar, valid once the engine stream has reached it)"""
        if example is not None:
            self.load(example)
        with torch.cuda.stream(self.stream):
            if self.graph_fb is not None:
                self.graph_fb.replay()
            else:
                self.loss = self._forward_backward()
            if self.avg is not None:
                self.avg.finish()
            if self.graph_opt is not None:
                self.graph_opt.replay()
            else:
                self.opt.step()
        return self.loss
