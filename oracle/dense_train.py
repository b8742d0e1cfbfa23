"""TEST INFRASTRUCTURE ONLY — differentiable dense-equivalent TRAINING forward of the sparse backbone, the checker of
the whole-step gradient parity test (tests/test_gpu_train.py).

Restates, with plain torch ops under autograd (fp32, TF32 off), what spconv + nn.BatchNorm1d compute in the reference's
train mode (det3d/models/backbones/base.py:145-213, PillarResNet.py:72-149,224-309):
  * SubMConv2d        = F.conv2d(x, W, bias, padding=1) evaluated at the active sites
  * SparseConv2d (s2) = F.conv2d(x, W, None, stride=2, padding=1) at the sites of max_pool2d(mask, 3, 2, 1)
  * BatchNorm1d       = batch statistics over the ACTIVE rows only (F.batch_norm on the gathered (M, C) matrix)
  * out = relu(bn(conv2(relu(bn(conv1(x))))) + identity); BlockV: identity = bn(conv0(x)) without ReLU
spconv itself is not available (parity unpinned); none of this library's kernels is used here.
"""
import torch
import torch.nn.functional as F


def _bn_active(y, mask, bn):
    """y (B,C,H,W), mask (B,1,H,W) in {0,1}: BatchNorm1d over the active positions, zeros elsewhere"""
    B, C, H, W = y.shape
    idx = mask.view(B, H, W).nonzero(as_tuple=True)
    rows = y.permute(0, 2, 3, 1)[idx]                           # (M, C)
    rows = F.batch_norm(rows, None, None, bn.weight, bn.bias, True, 0.0, bn.eps)
    out = torch.zeros(B, H, W, C, dtype=y.dtype, device=y.device).index_put(idx, rows)
    return out.permute(0, 3, 1, 2)


def backbone_train(backbone, x, mask):
    """x (B,C,H,W) zero at inactive sites, mask (B,1,H,W). Returns dict conv1..conv4 (dense, masked) [+ conv5]."""

    def subm(seq, x, mask, relu, res=None):
        conv, bn = seq[0], seq[1]
        y = _bn_active(F.conv2d(x, conv.weight.permute(0, 3, 1, 2), conv.bias, padding=1), mask, bn)
        if res is not None:
            y = y + res
        if relu:
            y = F.relu(y)
        return y * mask

    feats = {}
    for name in ("conv1", "conv2", "conv3", "conv4"):
        mods = list(getattr(backbone, name))
        i = 0
        if not hasattr(mods[0], "conv1"):
            conv, bn = mods[0], mods[1]
            mask = (F.max_pool2d(mask, 3, 2, 1) > 0).float()
            x = F.relu(_bn_active(F.conv2d(x, conv.weight.permute(0, 3, 1, 2), None, stride=2, padding=1), mask, bn)) * mask
            i = 3
        for b in mods[i:]:
            if hasattr(b, "conv0"):
                x = subm(b.conv0, x, mask, False)
            out = subm(b.conv1, x, mask, True)
            x = subm(b.conv2, out, mask, True, res=x)
        feats[name] = x
    if hasattr(backbone, "conv5"):
        feats["conv5"] = backbone.conv5(feats["conv4"])
    return feats


def loss_dense_equivalent(model, sp_feat, indices, B, H, W, example, train_cfg):
    """sum of the task losses with the sparse backbone replaced by its dense equivalent.  sp_feat (M,C) fp32 requires
    grad through the reader; indices (M,3) [b,y,x]."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    idx = (indices[:, 0].long(), indices[:, 1].long(), indices[:, 2].long())
    x = torch.zeros(B, H, W, sp_feat.shape[1], device=sp_feat.device).index_put(idx, sp_feat).permute(0, 3, 1, 2)
    mask = torch.zeros(B, H, W, 1, device=sp_feat.device).index_put(idx, torch.ones_like(sp_feat[:, :1])).permute(0, 3, 1, 2)
    feats = backbone_train(model.backbone, x, mask)
    bev = model.neck._forward_train(feats)
    preds = model.bbox_head._forward_train(bev)
    losses = model.bbox_head.loss(example, preds, train_cfg)
    return sum(l.sum() for l in losses["loss"])
