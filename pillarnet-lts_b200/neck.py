"""Dense BEV necks behind det3d's neck interface: RPNV1, RPNV2, RPNG, RPNGV2.

Mirrors det3d/models/necks/rpn.py:137-207 (RPNV1), :210-272 (RPNV2), :275-355 (RPNG), :358-450
(RPNGV2): same constructor kwargs and module tree (state_dict keys block_5.1.weight, deblock_5.0.weight,
...).  The nn.Conv2d / nn.ConvTranspose2d / BatchNorm2d children are parameter containers; every
conv+BN+ReLU runs as one NHWC gather-GEMM launch, and torch.cat is realised by writing both branches
into one wide buffer (channel offset in the epilogue).
"""
import logging

import torch
from torch import nn

from . import config, train
from .layers import (DenseMap, Sequential, build_norm_layer, dense_conv3x3, dense_deconv2x2, new_dense_rows,
                     use_padded_layout)
from .registry import NECKS
from .sparse import SparseConvTensor


def _to_dense_map(x, cat_room=False):
    """SparseConvTensor / NCHW tensor -> DenseMap. With cat_room the rows are 2C wide (left half filled)."""
    if isinstance(x, SparseConvTensor):
        t = x.table
        C = x.feat.shape[1]
        width = 2 * C if cat_room else C
        pad = 1 if use_padded_layout() else 0
        rows = new_dense_rows(t.B, t.H, t.W, width, x.feat.dtype, x.feat.device, pad)
        x.dense_nhwc(out=rows, out_coff=0, padded=bool(pad))
        return DenseMap(rows, t.B, t.H, t.W, C, 0, pad)
    return DenseMap.from_nchw(x)


def _cat_buffer(left, right_channels):
    """Returns (rows, coff_right): a buffer holding `left` in its first channels with room for the right branch."""
    if left.coff == 0 and left.rows.shape[1] >= left.C + right_channels:
        return left.rows, left.C
    rows = torch.empty(left.rows.shape[0], left.C + right_channels, dtype=left.rows.dtype, device=left.rows.device)
    rows[:, :left.C] = left.rows[:, left.coff:left.coff + left.C]
    return rows, left.C


class _RPNBase(nn.Module):
    def _build_layer(self, inplanes, planes, num_blocks, stride=1):
        """necks/rpn.py:172-185: ZeroPad2d(1)+valid 3x3 conv, then num_blocks x (3x3 pad 1)."""
        block = Sequential(
            nn.ZeroPad2d(1),
            nn.Conv2d(inplanes, planes, 3, stride=stride, bias=False),
            build_norm_layer(self.norm_cfg, planes)[1],
            nn.ReLU())
        for _ in range(num_blocks):
            block.add(nn.Conv2d(planes, planes, 3, padding=1, bias=False))
            block.add(build_norm_layer(self.norm_cfg, planes)[1])
            block.add(nn.ReLU())
        return block

    @staticmethod
    def _run_block(x, block, out=None, out_coff=0):
        """block = [ZeroPad2d, Conv, BN, ReLU, (Conv, BN, ReLU)*]."""
        mods = list(block)
        convs = [(mods[i], mods[i + 1]) for i in range(1, len(mods), 3)]
        for k, (conv, bn) in enumerate(convs):
            last = k == len(convs) - 1
            x = dense_conv3x3(x, conv, bn, relu=True, stride=conv.stride[0],
                              out=out if last else None, out_coff=out_coff if last else 0)
        return x

    @property
    def downsample_factor(self):
        return 1

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)


@NECKS.register_module
class RPNV1(_RPNBase):
    def __init__(self, layer_nums, num_filters, in_channels, norm_cfg=None, logger=None, **kwargs):
        super().__init__()
        self.norm_cfg = norm_cfg or dict(type="BN", momentum=0.01, eps=1e-3)
        self.block_5 = self._build_layer(in_channels[0], in_channels[0], layer_nums[0], stride=1)
        self.deblock_5 = Sequential(
            nn.ConvTranspose2d(in_channels[0], in_channels[1], 2, stride=2, bias=False),
            build_norm_layer(self.norm_cfg, in_channels[1])[1],
            nn.ReLU())
        self.block_4 = self._build_layer(in_channels[1] * 2, num_filters, layer_nums[1], stride=1)
        (logger or logging.getLogger("RPN")).info("Finish RPN Initialization")

    def _forward_train(self, f):
        """necks/rpn.py:196-207 with the torch modules themselves (training: cuDNN + autograd)"""
        with train.autocast_ctx():
            x4, x5 = train.to_dense(f["conv4"]), train.to_dense(f["conv5"])
            run = train.run_dense_seq
            up = run(self.deblock_5, run(self.block_5, x5))
            return tuple([run(self.block_4, torch.cat([x4, up], dim=1))])

    def forward(self, pillar_features, **kwargs):
        if self.training:
            return self._forward_train(pillar_features)
        x4 = _to_dense_map(pillar_features["conv4"], cat_room=True)
        x5 = _to_dense_map(pillar_features["conv5"])
        x = self._run_block(x5, self.block_5)
        up_c = self.deblock_5[0].out_channels
        rows, coff = _cat_buffer(x4, up_c)
        dense_deconv2x2(x, self.deblock_5[0], self.deblock_5[1], relu=True, out=rows, out_coff=coff)
        cat = DenseMap(rows, x4.B, x4.H, x4.W, x4.C + up_c, 0, x4.pad)
        x = self._run_block(cat, self.block_4)
        return tuple([x.nchw()])


@NECKS.register_module
class RPNV2(_RPNBase):
    def __init__(self, layer_nums, in_channels, num_filters, norm_cfg=None, logger=None):
        super().__init__()
        self.norm_cfg = norm_cfg or dict(type="BN", momentum=0.01, eps=1e-3)
        self.block_4 = self._build_layer(in_channels[0], in_channels[0], layer_nums[0], stride=1)
        self.deblock_4 = Sequential(
            nn.ConvTranspose2d(in_channels[0], in_channels[1], 2, stride=2, bias=False),
            build_norm_layer(self.norm_cfg, in_channels[1])[1],
            nn.ReLU())
        self.block_3 = self._build_layer(in_channels[1] * 2, num_filters, layer_nums[1], stride=1)
        (logger or logging.getLogger("RPN")).info("Finish RPN Initialization")

    def _forward_train(self, f):
        """necks/rpn.py:262-272"""
        with train.autocast_ctx():
            x3, x4 = train.to_dense(f["conv3"]), train.to_dense(f["conv4"])
            run = train.run_dense_seq
            up = run(self.deblock_4, run(self.block_4, x4))
            return tuple([run(self.block_3, torch.cat([x3, up], dim=1))])

    def forward(self, pillar_features, **kwargs):
        if self.training:
            return self._forward_train(pillar_features)
        x3 = _to_dense_map(pillar_features["conv3"], cat_room=True)
        x4 = _to_dense_map(pillar_features["conv4"])
        x = self._run_block(x4, self.block_4)
        up_c = self.deblock_4[0].out_channels
        rows, coff = _cat_buffer(x3, up_c)
        dense_deconv2x2(x, self.deblock_4[0], self.deblock_4[1], relu=True, out=rows, out_coff=coff)
        cat = DenseMap(rows, x3.B, x3.H, x3.W, x3.C + up_c, 0, x3.pad)
        x = self._run_block(cat, self.block_3)
        return tuple([x.nchw()])


@NECKS.register_module
class RPNG(_RPNBase):
    def __init__(self, layer_nums, in_channels, num_filters, norm_cfg=None, logger=None):
        super().__init__()
        self.norm_cfg = norm_cfg or dict(type="BN", eps=1e-3, momentum=0.01)
        self.block_5 = self._build_layer(in_channels[0], in_channels[0], layer_nums[0], stride=1)
        self.top_down_54 = Sequential(
            nn.ConvTranspose2d(in_channels[0], in_channels[1], 2, stride=2, bias=False),
            build_norm_layer(self.norm_cfg, in_channels[1])[1],
            nn.ReLU())
        self.block_4 = self._build_layer(in_channels[1] * 2, num_filters[0], layer_nums[0])
        self.top_down_43 = Sequential(
            nn.ConvTranspose2d(num_filters[0], in_channels[2], 2, stride=2, bias=False),
            build_norm_layer(self.norm_cfg, in_channels[2])[1],
            nn.ReLU())
        self.block_3 = self._build_layer(in_channels[2] * 2, num_filters[1], layer_nums[1])
        (logger or logging.getLogger("RPN")).info("Finish RPN Initialization")

    def _forward_train(self, f):
        """necks/rpn.py:336-355"""
        with train.autocast_ctx():
            x3, x4, x5 = (train.to_dense(f[k]) for k in ("conv3", "conv4", "conv5"))
            run = train.run_dense_seq
            x5 = run(self.block_5, x5)
            x4 = run(self.block_4, torch.cat([x4, run(self.top_down_54, x5)], dim=1))
            x3 = run(self.block_3, torch.cat([x3, run(self.top_down_43, x4)], dim=1))
            return tuple([x4, x3])

    def forward(self, pillar_features, **kwargs):
        if self.training:
            return self._forward_train(pillar_features)
        x3 = _to_dense_map(pillar_features["conv3"], cat_room=True)
        x4 = _to_dense_map(pillar_features["conv4"], cat_room=True)
        x5 = _to_dense_map(pillar_features["conv5"])
        # head stride 8
        x5 = self._run_block(x5, self.block_5)
        up_c = self.top_down_54[0].out_channels
        rows, coff = _cat_buffer(x4, up_c)
        dense_deconv2x2(x5, self.top_down_54[0], self.top_down_54[1], relu=True, out=rows, out_coff=coff)
        x4o = self._run_block(DenseMap(rows, x4.B, x4.H, x4.W, x4.C + up_c, 0, x4.pad), self.block_4)
        # head stride 4
        up_c = self.top_down_43[0].out_channels
        rows, coff = _cat_buffer(x3, up_c)
        dense_deconv2x2(x4o, self.top_down_43[0], self.top_down_43[1], relu=True, out=rows, out_coff=coff)
        x3o = self._run_block(DenseMap(rows, x3.B, x3.H, x3.W, x3.C + up_c, 0, x3.pad), self.block_3)
        return tuple([x4o.nchw(), x3o.nchw()])


@NECKS.register_module
class RPNGV2(_RPNBase):
    def __init__(self, layer_nums, in_channels, num_filters, norm_cfg=None, logger=None):
        super().__init__()
        self.norm_cfg = norm_cfg or dict(type="BN", eps=1e-3, momentum=0.01)
        self.block_5 = self._build_layer(in_channels[0], in_channels[0], layer_nums[0], stride=1)
        self.top_down_54 = Sequential(
            nn.ConvTranspose2d(in_channels[0], num_filters[0] // 2, 2, stride=2, bias=False),
            build_norm_layer(self.norm_cfg, num_filters[0] // 2)[1],
            nn.ReLU())
        self.reduce_4 = Sequential(
            nn.Conv2d(in_channels[1], num_filters[0] // 2, 3, padding=1, bias=False),
            build_norm_layer(self.norm_cfg, num_filters[0] // 2)[1],
            nn.ReLU())
        self.block_4 = self._build_layer(num_filters[0], num_filters[0], layer_nums[0])
        self.top_down_43 = Sequential(
            nn.ConvTranspose2d(num_filters[0], num_filters[1] // 2, 2, stride=2, bias=False),
            build_norm_layer(self.norm_cfg, num_filters[1] // 2)[1],
            nn.ReLU())
        self.reduce_3 = Sequential(
            nn.Conv2d(in_channels[2], num_filters[1] // 2, 3, padding=1, bias=False),
            build_norm_layer(self.norm_cfg, num_filters[1] // 2)[1],
            nn.ReLU())
        self.block_3 = self._build_layer(num_filters[1], num_filters[1], layer_nums[1])
        (logger or logging.getLogger("RPN")).info("Finish RPN Initialization")

    def _forward_train(self, f):
        """necks/rpn.py:428-450"""
        with train.autocast_ctx():
            x3, x4, x5 = (train.to_dense(f[k]) for k in ("conv3", "conv4", "conv5"))
            run = train.run_dense_seq
            x5 = run(self.block_5, x5)
            x4 = run(self.block_4, torch.cat([run(self.reduce_4, x4), run(self.top_down_54, x5)], dim=1))
            x3 = run(self.block_3, torch.cat([run(self.reduce_3, x3), run(self.top_down_43, x4)], dim=1))
            return tuple([x4, x3])

    def forward(self, pillar_features, **kwargs):
        if self.training:
            return self._forward_train(pillar_features)
        x3 = _to_dense_map(pillar_features["conv3"])
        x4 = _to_dense_map(pillar_features["conv4"])
        x5 = _to_dense_map(pillar_features["conv5"])
        dt, dev = x4.rows.dtype, x4.rows.device
        # head stride 8
        half = self.reduce_4[0].out_channels
        rows = new_dense_rows(x4.B, x4.H, x4.W, 2 * half, dt, dev, x4.pad)
        dense_conv3x3(x4, self.reduce_4[0], self.reduce_4[1], relu=True, out=rows, out_coff=0)
        x5 = self._run_block(x5, self.block_5)
        dense_deconv2x2(x5, self.top_down_54[0], self.top_down_54[1], relu=True, out=rows, out_coff=half)
        x4o = self._run_block(DenseMap(rows, x4.B, x4.H, x4.W, 2 * half, 0, x4.pad), self.block_4)
        # head stride 4
        half = self.reduce_3[0].out_channels
        rows = new_dense_rows(x3.B, x3.H, x3.W, 2 * half, dt, dev, x3.pad)
        dense_conv3x3(x3, self.reduce_3[0], self.reduce_3[1], relu=True, out=rows, out_coff=0)
        dense_deconv2x2(x4o, self.top_down_43[0], self.top_down_43[1], relu=True, out=rows, out_coff=half)
        x3o = self._run_block(DenseMap(rows, x3.B, x3.H, x3.W, 2 * half, 0, x3.pad), self.block_3)
        return tuple([x4o.nchw(), x3o.nchw()])
