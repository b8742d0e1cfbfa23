"""Pillar R-CNN RoI / point heads, inference path (SURVEY §8 f rank 3).

Mirrors det3d/models/roi_heads/roi_mix_head.py:16-122 (`RoIMIXHead`), roi_head_template.py:14-39,189-219
(`make_fc_layers`, `generate_predicted_boxes`), mlp_layers.py:23-115 (`MLPMixer`, `ResMLPLayer`) and
det3d/models/point_heads/point_head_simple.py:14-96 (`PointHead`): same constructor kwargs, module trees and
state_dict keys.  The per-RoI "FC" stacks (`Conv1d(k=1)` / `Linear` + BN1d + ReLU) run as GEMMs on the library's conv
kernels (`pn_conv_gather`, taps = 1, folded BN in the epilogue); the box refinement + score fusion is `pn_roi_refine`.
Target assignment and the losses (training) are not part of the inference path and raise.
"""
import torch
from torch import nn

from . import config, ops
from .layers import build_norm_layer, lower, run_conv
from .registry import POINT_HEAD, ROI_HEAD, ConfigDict


class MLPMixer(nn.Module):
    """mlp_layers.py:23-60"""

    def __init__(self, in_channels, num_patches, expansion_factor=4, expansion_factor_token=0.5):
        super().__init__()
        inner = int(num_patches * expansion_factor)
        self.token_mixer = nn.Sequential(build_norm_layer(dict(type="LN"), in_channels)[1],
                                         nn.Conv1d(num_patches, inner, kernel_size=1), nn.GELU(),
                                         nn.Conv1d(inner, num_patches, kernel_size=1))
        inner = int(in_channels * expansion_factor_token)
        self.channel_mixer = nn.Sequential(build_norm_layer(dict(type="LN"), in_channels)[1],
                                           nn.Linear(in_channels, inner), nn.GELU(), nn.Linear(inner, in_channels))

    def forward(self, x):
        x = self.token_mixer(x) + x
        return self.channel_mixer(x) + x


class Affine(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.alpha = nn.Parameter(torch.ones((1, 1, dim)))
        self.beta = nn.Parameter(torch.zeros((1, 1, dim)))

    def forward(self, x):
        return self.alpha * x + self.beta


class ResMLPLayer(nn.Module):
    """mlp_layers.py:76-115 (the Rearrange pair of the token mixer is a transpose on either side of the Linear)"""

    def __init__(self, in_channels, num_patches, expansion_factor=2, layer_scale_init=1e-4):
        super().__init__()
        self.token_aff = Affine(in_channels)
        self.token_scale = nn.Parameter(layer_scale_init * torch.ones(in_channels))
        self.token_mixer = nn.Sequential(nn.Identity(), nn.Linear(num_patches, num_patches), nn.Identity())
        self.channel_aff = Affine(in_channels)
        self.channel_scale = nn.Parameter(layer_scale_init * torch.ones(in_channels))
        self.channel_mixer = nn.Sequential(nn.Linear(in_channels, in_channels * expansion_factor), nn.GELU(),
                                           nn.Linear(in_channels * expansion_factor, in_channels))
        self.post_aff = Affine(in_channels)

    def forward(self, x):
        x = self.token_aff(x)
        x = x + self.token_scale * self.token_mixer[1](x.transpose(1, 2)).transpose(1, 2)
        x = self.channel_aff(x)
        x = x + self.channel_scale * self.channel_mixer(x)
        return self.post_aff(x)


def run_fc_stack(x2d, layers):
    """(n, C) rows through a Sequential of [Conv1d(k=1) | Linear, BN1d, ReLU, (Dropout)]* + optional last Conv1d / Linear:
    every conv/linear (+ its BN + ReLU) is one GEMM launch with the folded affine in the epilogue.  Returns f32 rows
    when the last layer has no BN (a head output), else rows in the activation dtype."""
    mods = list(layers)
    act = config.act_dtype()
    if x2d.dtype != act:
        x2d = x2d.to(act)
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Dropout):
            i += 1
            continue
        if not isinstance(m, (nn.Conv1d, nn.Linear)):
            raise TypeError(f"unexpected layer {type(m)} in an FC stack")
        bn = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d) else None
        relu = bn is not None and i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
        cin = m.in_features if isinstance(m, nn.Linear) else m.in_channels
        cout = m.out_features if isinstance(m, nn.Linear) else m.out_channels
        last = bn is None
        x2d = run_conv(x2d.contiguous(), lower(m, bn), None, 1, cin, cout, x2d.shape[0], relu=relu,
                       out_dtype=torch.float32 if last else None)
        i += 1 + (1 if bn is not None else 0) + (1 if relu else 0)
    return x2d


class RoIHeadTemplate(nn.Module):
    """roi_head_template.py:14-39,189-219 without the proposal-target layer and the losses (training)"""

    def __init__(self, num_class, model_cfg):
        super().__init__()
        self.model_cfg = ConfigDict.wrap(model_cfg)
        self.num_class = num_class
        self.forward_ret_dict = None

    def make_fc_layers(self, input_channels, output_channels, fc_list):
        fc_layers, pre = [], input_channels
        norm_cfg = dict(type="BN1d", eps=1e-3, momentum=0.01)
        for k in range(len(fc_list)):
            fc_layers += [nn.Conv1d(pre, fc_list[k], kernel_size=1, bias=False), build_norm_layer(norm_cfg, fc_list[k])[1],
                          nn.ReLU()]
            pre = fc_list[k]
            if self.model_cfg.DP_RATIO >= 0 and k == 0:
                fc_layers.append(nn.Dropout(self.model_cfg.DP_RATIO))
        fc_layers.append(nn.Conv1d(pre, output_channels, kernel_size=1, bias=True))
        return nn.Sequential(*fc_layers)

    def assign_targets(self, batch_dict):
        raise NotImplementedError("RoI target assignment (training) is outside the inference path")

    def get_loss(self, tb_dict=None):
        raise NotImplementedError("RoI head losses (training) are outside the inference path")


@ROI_HEAD.register_module
class RoIMIXHead(RoIHeadTemplate):
    def __init__(self, in_channels, model_cfg, num_class=1, code_size=7, add_box_param=False, test_cfg=None,
                 mixer_type=None, num_patches=49, **kwargs):
        super().__init__(num_class=num_class, model_cfg=model_cfg)
        self.test_cfg, self.code_size, self.add_box_param, self.num_patches = test_cfg, code_size, add_box_param, num_patches
        pre = in_channels * num_patches
        if mixer_type == "MLPMixer":
            self.mlp_mixer = MLPMixer(in_channels=in_channels, num_patches=num_patches)
        elif mixer_type == "ResMLP":
            self.mlp_mixer = ResMLPLayer(in_channels=in_channels, num_patches=num_patches)
        else:
            self.mlp_mixer = nn.Sequential()
        shared, norm_cfg = [], dict(type="BN1d", eps=1e-3, momentum=0.01)
        fcs = self.model_cfg.SHARED_FC
        for k in range(len(fcs)):
            shared += [nn.Conv1d(pre, fcs[k], kernel_size=1, bias=False), build_norm_layer(norm_cfg, fcs[k])[1], nn.ReLU()]
            pre = fcs[k]
            if k != len(fcs) - 1 and self.model_cfg.DP_RATIO > 0:
                shared.append(nn.Dropout(self.model_cfg.DP_RATIO))
        self.shared_fc_layer = nn.Sequential(*shared)
        self.cls_layers = self.make_fc_layers(pre, self.num_class, self.model_cfg.CLS_FC)
        self.reg_layers = self.make_fc_layers(pre, code_size, self.model_cfg.REG_FC)
        self.init_weights()

    def init_weights(self, weight_init="xavier"):
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Conv1d)):
                nn.init.xavier_normal_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        nn.init.normal_(self.reg_layers[-1].weight, mean=0, std=0.001)

    def forward(self, batch_dict, training=True):
        if training or self.training:
            raise NotImplementedError("RoIMIXHead: only the inference path is implemented")
        rois = batch_dict["rois"]
        B, N = rois.shape[:2]
        feats = batch_dict["roi_features"].view(B * N, self.num_patches, -1)
        if not (isinstance(self.mlp_mixer, nn.Sequential) and len(self.mlp_mixer) == 0):
            feats = self.mlp_mixer(feats.float())            # LayerNorm / GELU mixers: PyTorch (not in the shipped config)
        if self.add_box_param:
            raise NotImplementedError("add_box_param concatenates tensors of different rank in the reference")
        pooled = feats.reshape(B * N, -1)
        shared = run_fc_stack(pooled, self.shared_fc_layer)
        rcnn_cls = run_fc_stack(shared, self.cls_layers)      # (B*N, num_class)
        rcnn_reg = run_fc_stack(shared, self.reg_layers)      # (B*N, code_size)
        batch_dict["rcnn_cls"], batch_dict["rcnn_reg"] = rcnn_cls, rcnn_reg
        batch_dict["batch_cls_preds"] = rcnn_cls.view(B, N, -1)
        boxes, scores, valid = ops.roi_refine(rois.float().contiguous(), rcnn_reg, rcnn_cls[:, 0], batch_dict["roi_scores"],
                                              batch_dict.get("roi_labels"))
        batch_dict["batch_box_preds"] = boxes                 # generate_predicted_boxes
        batch_dict["refined_scores"], batch_dict["refined_valid"] = scores, valid
        batch_dict["cls_preds_normalized"] = False
        return batch_dict


@POINT_HEAD.register_module
class PointHead(nn.Module):
    """point_head_simple.py:14-96: an auxiliary per-grid-point foreground classifier; at inference it only adds
    `point_cls_scores` (and re-weights the features when ATT_MODEL is set)"""

    def __init__(self, in_channels, num_class, model_cfg, **kwargs):
        super().__init__()
        self.model_cfg, self.num_class = ConfigDict.wrap(model_cfg), num_class
        layers, c_in = [], in_channels
        for c in self.model_cfg.CLS_FC:
            layers += [nn.Linear(c_in, c, bias=False), build_norm_layer(dict(type="BN1d", eps=1e-3, momentum=0.01), c)[1],
                       nn.ReLU()]
            c_in = c
        layers.append(nn.Linear(c_in, 1, bias=True))
        self.cls_layers = nn.Sequential(*layers)
        self.forward_ret_dict = None

    def forward(self, batch_dict):
        if self.training:
            raise NotImplementedError("PointHead target assignment (training) is outside the inference path")
        pf = batch_dict["point_features"]
        preds = run_fc_stack(pf.reshape(-1, pf.shape[-1]), self.cls_layers)
        self.forward_ret_dict = {"point_cls_preds": preds}
        batch_dict["point_cls_scores"] = torch.sigmoid(preds)
        if self.model_cfg.get("ATT_MODEL", False):
            batch_dict["point_features"] = pf * batch_dict["point_cls_scores"].view(*pf.shape[:-1], 1).to(pf.dtype)
        return batch_dict

    def get_loss(self, tb_dict=None):
        raise NotImplementedError("PointHead loss (training) is outside the inference path")
