"""PillarNet sparse 2D ResNet encoders behind det3d's backbone interface.

Mirrors det3d/models/backbones/PillarResNet.py:8-309 (PillarResNet18S/18/34S/34: same constructor,
same module tree => same state_dict keys, `backbone_channels` / `backbone_strides`) and
det3d/models/backbones/base.py:145-213 (Sparse2DBasicBlockV / Sparse2DBasicBlock).  Every
conv+BN(+ReLU)(+residual) is one gather-GEMM launch (ops.conv_gather) with the epilogue fused; the
rulebook of a stage is built once (spconv's `indice_key` cache) and shared by its SubM convs.
"""
import torch
from torch import nn

from . import config, ops, train
from ._lib import PN_NBR_SUBM_SORTED
from .layers import (DenseMap, SparseConv2d, SparseReLU, SparseSequential, SubMConv2d,
                     build_norm_layer, dense_conv3x3, lower, new_dense_rows, run_conv, use_padded_layout)
from .registry import BACKBONES
from .sparse import SparseConvTensor


def conv2D3x3(in_planes, out_planes, stride=1, dilation=1, indice_key=None, bias=True):
    """backbones/base.py:38-63."""
    assert stride >= 1
    if stride == 1:
        return SubMConv2d(in_planes, out_planes, 3, stride=1, dilation=dilation, padding=dilation,
                          bias=bias, indice_key=indice_key)
    return SparseConv2d(in_planes, out_planes, 3, stride=stride, dilation=dilation, padding=dilation,
                        bias=bias, indice_key=indice_key)


# Row counts seen on a real batch, per raster shape (B, H, W).  The live count of a stage stays on the device and a
# captured graph freezes each conv's tile shape, so the engine measures one eager warm-up pass
# (observe_rows_begin/end, one host sync) before it captures: with the static guess below, stage 4 of a nuScenes
# frame (10.3 k rows against the guessed 8.1 k) got 162 tiles for 148 SMs, i.e. two waves of which the second is 9 % full.
_observed_rows = {}
_observing = None


def observe_rows_begin():
    global _observing
    _observing = []


def observe_rows_end():
    """host sync: records the active-row counts of the tables the last forward pass went through"""
    global _observing
    tables, _observing = _observing or [], None
    for t in tables:
        n = t.count()
        if n > 0:
            _observed_rows[(t.B, t.H, t.W)] = n
    return dict(_observed_rows)


def _rows_hint(table):
    """Expected active rows of a stage (the live count stays on the device); only steers the conv tile shape.
    Observed counts when the engine has measured a batch, else LiDAR BEV occupancy guessed as 25 % of the cells."""
    seen = _observed_rows.get((table.B, table.H, table.W))
    if seen:
        return min(table.cap, seen)
    cells = table.B * table.H * table.W
    return min(table.cap, max(1, cells // 4))


def _subm(sp, seq, relu, residual=None):
    """SparseSequential(SubMConv2d, BN[, SparseReLU]) as one launch."""
    conv, bn = seq[0], seq[1]
    if bn.training:
        return train.subm_block(sp, seq, relu, residual)
    t = sp.table
    lw = lower(conv, bn)
    out = run_conv(sp.feat, lw, t.subm_nbr(), 9, conv.in_channels, conv.out_channels, t.cap, num=t.num,
                   relu=relu, residual=residual, rows_hint=_rows_hint(t), nbr_kind=PN_NBR_SUBM_SORTED,
                   nbr_plan=t.subm_plan() if config.get_precision() == "bf16" else None)
    return SparseConvTensor(out, t, sp.spatial_shape, sp.batch_size)


class Sparse2DBasicBlockV(nn.Module):
    expansion = 1

    def __init__(self, planes, norm_cfg=None, indice_key=None):
        super().__init__()
        if norm_cfg is None:
            norm_cfg = dict(type="BN1d", momentum=0.01, eps=1e-3)
        bias = norm_cfg is not None
        self.conv0 = SparseSequential(conv2D3x3(planes, planes, indice_key=indice_key, bias=bias),
                                      build_norm_layer(norm_cfg, planes)[1])
        self.conv1 = SparseSequential(conv2D3x3(planes, planes, indice_key=indice_key, bias=bias),
                                      build_norm_layer(norm_cfg, planes)[1], SparseReLU())
        self.conv2 = SparseSequential(conv2D3x3(planes, planes, indice_key=indice_key, bias=bias),
                                      build_norm_layer(norm_cfg, planes)[1])
        self.relu = nn.ReLU()

    def forward(self, x):
        x = _subm(x, self.conv0, relu=False)
        out = _subm(x, self.conv1, relu=True)
        return _subm(out, self.conv2, relu=True, residual=x.feat)  # relu(bn(conv) + identity)


class Sparse2DBasicBlock(nn.Module):
    expansion = 1

    def __init__(self, planes, norm_cfg=None, indice_key=None):
        super().__init__()
        if norm_cfg is None:
            norm_cfg = dict(type="BN1d", momentum=0.01, eps=1e-3)
        bias = norm_cfg is not None
        self.conv1 = SparseSequential(conv2D3x3(planes, planes, indice_key=indice_key, bias=bias),
                                      build_norm_layer(norm_cfg, planes)[1], SparseReLU())
        self.conv2 = SparseSequential(conv2D3x3(planes, planes, indice_key=indice_key, bias=bias),
                                      build_norm_layer(norm_cfg, planes)[1])
        self.relu = nn.ReLU()

    def forward(self, x):
        out = _subm(x, self.conv1, relu=True)
        return _subm(out, self.conv2, relu=True, residual=x.feat)


_side_streams = {}


def _prefetch_rulebooks(table, n_levels):
    """Rulebooks depend only on the active-site sets, not on features: all strided levels (and their submanifold
    tables) are built on a side stream while the main stream runs the first stage's convs.  Returns
    [(out_table, nbr_down, ready_event)] per level.  Fork/join is by events, so it is CUDA-graph capturable."""
    main = torch.cuda.current_stream()
    dev = table.coords.device
    side = _side_streams.get(dev)
    if side is None:
        # high priority: its small kernels must squeeze in beside the persistent conv CTAs of the main stream
        # (measured: at default priority k_emit of level 2 took 49 us beside the stage-1 convs and stage 2 waited 48 us)
        side = _side_streams[dev] = torch.cuda.Stream(device=dev, priority=-1)
    # Fork where the level-0 table became complete (right after pn_pillarize), not at the current end of the main
    # stream: the rulebook chain then runs beside the PFN and the first conv instead of beside every conv of stages
    # 1-3, whose persistent CTAs starved it (each small kernel took 20-30 us) and were slowed down by it in turn.
    if table.ready is not None:
        side.wait_event(table.ready)
    else:
        side.wait_stream(main)
    out = []
    with torch.cuda.stream(side):
        levels = ops.rulebook_pyramid(table, n_levels)
        if config.get_precision() == "bf16":
            for ot, _ in levels:
                ot.subm_plan()           # tile plans of the level's submanifold table, off the main stream too
        ev = torch.cuda.Event()
        ev.record(side)
        for ot, nbr in levels:
            for x in (ot.words, ot.prefix, ot.coords, ot.num, nbr, ot._nbr_subm, ot._plan_subm):
                if x is None:
                    continue
                x.record_stream(main)       # allocated on the side stream, consumed on the main stream
            out.append((ot, nbr, ev))
    return out


def _run_stage(sp, stage, pre=None):
    """Runs one `convN` SparseSequential: optional [SparseConv2d, BN, SparseReLU] head then blocks.
    pre = (out_table, nbr, event) when the level's rulebook was prefetched on the side stream."""
    mods = list(stage)
    i = 0
    if isinstance(mods[0], SparseConv2d):
        conv, bn = mods[0], mods[1]
        if bn.training:
            sp = train.down_block(sp, conv, bn)
            for m in mods[3:]:
                sp = m(sp)
            return sp
        if pre is not None:
            out_table, nbr, ev = pre
            torch.cuda.current_stream().wait_event(ev)
        else:
            out_table, nbr = ops.rulebook_down3x3s2(sp.table)
        lw = lower(conv, bn)
        feat = run_conv(sp.feat, lw, nbr, 9, conv.in_channels, conv.out_channels, out_table.cap,
                        num=out_table.num, relu=True, rows_hint=_rows_hint(out_table))
        sp = SparseConvTensor(feat, out_table, (out_table.H, out_table.W), sp.batch_size)
        i = 3
    for m in mods[i:]:
        sp = m(sp)
    return sp


def post_act_block_dense(in_channels, kernel_size, stride=1, padding=0, dilation=1, norm_cfg=None):
    """backbones/base.py:100-108."""
    return nn.Sequential(
        nn.Conv2d(in_channels, in_channels, kernel_size, stride, padding=padding, dilation=dilation, bias=False),
        build_norm_layer(norm_cfg, in_channels)[1],
        SparseReLU(),
    )


class _PillarResNet(nn.Module):
    """Shared constructor/forward; subclasses give the block counts and whether conv5 exists."""
    BLOCKS = (1, 2, 2, 2)   # extra Sparse2DBasicBlocks in conv1 (after BlockV), then blocks in conv2..4
    DENSE = True

    def __init__(self, in_channels=32, **kwargs):
        super().__init__()
        c = in_channels
        norm_cfg = dict(type="BN1d", momentum=0.01, eps=1e-3)
        self.conv1 = SparseSequential(
            Sparse2DBasicBlockV(c, norm_cfg=norm_cfg, indice_key="res1"),
            *[Sparse2DBasicBlock(c, norm_cfg=norm_cfg, indice_key="res1") for _ in range(self.BLOCKS[0])])

        def down(cin, cout, n, key):
            return SparseSequential(
                SparseConv2d(cin, cout, 3, 2, padding=1, bias=False),
                build_norm_layer(norm_cfg, cout)[1],
                SparseReLU(),
                *[Sparse2DBasicBlock(cout, norm_cfg=norm_cfg, indice_key=key) for _ in range(n)])

        self.conv2 = down(c, c * 2, self.BLOCKS[1], "res2")
        self.conv3 = down(c * 2, c * 4, self.BLOCKS[2], "res3")
        self.conv4 = down(c * 4, c * 8, self.BLOCKS[3], "res4")
        self.backbone_channels = {"conv1": 32, "conv2": 64, "conv3": 128, "conv4": 256}
        self.backbone_strides = {"conv1": 1, "conv2": 2, "conv3": 4, "conv4": 8}
        if self.DENSE:
            norm_cfg2 = dict(type="BN", momentum=0.01, eps=1e-3)
            self.conv5 = nn.Sequential(
                nn.Conv2d(256, 256, 3, 2, padding=1, bias=False),
                build_norm_layer(norm_cfg2, 256)[1],
                nn.ReLU(),
                post_act_block_dense(256, 3, padding=1, norm_cfg=norm_cfg2),
                post_act_block_dense(256, 3, padding=1, norm_cfg=norm_cfg2),
            )
            self.backbone_channels["conv5"] = 256
            self.backbone_strides["conv5"] = 16

    def forward(self, sp_tensor):
        pre = [None, None, None]
        if not self.training and config.overlap_rulebooks():
            sp_tensor.table.subm_nbr()                      # needed at once by conv1: main stream
            if config.get_precision() == "bf16":
                sp_tensor.table.subm_plan()
            pre = _prefetch_rulebooks(sp_tensor.table, 3)
        x1 = _run_stage(sp_tensor, self.conv1)
        x2 = _run_stage(x1, self.conv2, pre[0])
        x3 = _run_stage(x2, self.conv3, pre[1])
        x4 = _run_stage(x3, self.conv4, pre[2])
        feats = {"conv1": x1, "conv2": x2, "conv3": x3, "conv4": x4}
        if _observing is not None:
            _observing.extend(x.table for x in (x1, x2, x3, x4))
        if self.training:
            if self.DENSE:
                # training: dense conv5 runs in PyTorch (cuDNN) under autograd
                with train.autocast_ctx():
                    d4 = train.dense_from_sparse(x4)
                    feats["conv4"] = d4
                    feats["conv5"] = train.run_dense_seq(self.conv5, d4)
            return feats
        if self.DENSE:
            # x_conv4.dense() lands in the left half of a 2C-wide buffer so the neck's channel concat
            # (necks/rpn.py:201-205) needs no copy: the up-sampled branch writes the right half.
            t = x4.table
            C = x4.feat.shape[1]
            pad = 1 if use_padded_layout() else 0
            cat_rows = new_dense_rows(t.B, t.H, t.W, 2 * C, x4.feat.dtype, x4.feat.device, pad)
            x4.dense_nhwc(out=cat_rows, out_coff=0, padded=bool(pad))
            d4 = DenseMap(cat_rows, t.B, t.H, t.W, C, 0, pad)
            c5 = self.conv5
            d5 = dense_conv3x3(d4, c5[0], c5[1], relu=True, stride=2)
            d5 = dense_conv3x3(d5, c5[3][0], c5[3][1], relu=True)
            d5 = dense_conv3x3(d5, c5[4][0], c5[4][1], relu=True)
            feats["conv4"] = d4.nchw()
            feats["conv5"] = d5.nchw()
        return feats


@BACKBONES.register_module
class PillarResNet18S(_PillarResNet):
    BLOCKS, DENSE = (1, 2, 2, 2), False


@BACKBONES.register_module
class PillarResNet18(_PillarResNet):
    BLOCKS, DENSE = (1, 2, 2, 2), True


@BACKBONES.register_module
class PillarResNet34S(_PillarResNet):
    BLOCKS, DENSE = (2, 4, 6, 3), False


@BACKBONES.register_module
class PillarResNet34(_PillarResNet):
    BLOCKS, DENSE = (2, 4, 6, 3), True
