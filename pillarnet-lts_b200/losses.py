"""CenterHead training losses in PyTorch ("loss stays PyTorch").

Restates det3d/models/losses/centernet_loss.py:9-125 (RegLoss, FastFocalLoss, IouLoss, IouRegLoss) and the
axis-aligned IoU / GIoU / DIoU of det3d/core/utils/center_utils.py:124-228; the IoU-head target
(boxes_aligned_iou3d_gpu, ops/iou3d_nms/iou3d_nms_utils.py:76-113) uses the library's rotated-overlap kernel
(pn_boxes_aligned_overlap_bev) in place of iou3d_nms_cuda.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from . import ops


def gather_feat(feat, ind):
    """feat (B,H,W,C), ind (B,M) flat pixel indices -> (B,M,C)   (center_utils.py:65-78)"""
    B, H, W, C = feat.shape
    feat = feat.reshape(B, H * W, C)
    return feat.gather(1, ind.unsqueeze(2).expand(B, ind.shape[1], C))


class RegLoss(nn.Module):
    """per-channel masked L1 over the positive objects, normalised by their count (centernet_loss.py:9-32)"""

    def forward(self, output, mask, ind, target):
        # The reference returns zeros((1,)) when there is no positive (`if mask.sum() == 0`, a host sync per task and
        # step); without positives every masked term below is exactly 0, so the value is the same and the step
        # stays sync-free.
        pred = gather_feat(output, ind)
        m = mask.float().unsqueeze(2)
        loss = F.l1_loss(pred * m, target * m, reduction="none")
        loss = loss / (m.sum() + 1e-4)
        return loss.sum(dim=(0, 1))


class FastFocalLoss(nn.Module):
    """CornerNet focal loss with the positives gathered at the object centres (centernet_loss.py:35-67)"""

    def forward(self, out, target, ind, mask, cat):
        mask = mask.float()
        neg = (torch.log(1 - out) * out.pow(2) * (1 - target).pow(4)).sum()
        pos_pred = gather_feat(out, ind).gather(2, cat.unsqueeze(2))       # (B,M,1)
        num_pos = mask.sum()
        pos = (torch.log(pos_pred) * (1 - pos_pred).pow(2) * mask.unsqueeze(2)).sum()
        # `if num_pos == 0: return -neg` of the reference without the host sync (pos is exactly 0 there)
        return torch.where(num_pos == 0, -neg, -(pos + neg) / num_pos.clamp(min=1.0))


def to_pcdet(boxes):
    """iou3d_nms_utils.py:30-34"""
    boxes = boxes[:, [0, 1, 2, 4, 3, 5, -1]]
    boxes[:, -1] = -boxes[:, -1] - math.pi / 2
    return boxes


def boxes_aligned_iou3d(boxes_a, boxes_b):
    """pairwise-aligned rotated 3-D IoU, (n,1)  (iou3d_nms_utils.py:76-113)"""
    a, b = to_pcdet(boxes_a.float()), to_pcdet(boxes_b.float())
    a_max, a_min = (a[:, 2] + a[:, 5] / 2).view(-1, 1), (a[:, 2] - a[:, 5] / 2).view(-1, 1)
    b_max, b_min = (b[:, 2] + b[:, 5] / 2).view(-1, 1), (b[:, 2] - b[:, 5] / 2).view(-1, 1)
    bev = ops.boxes_aligned_overlap_bev(a.contiguous(), b.contiguous()).view(-1, 1)
    h = torch.clamp(torch.min(a_max, b_max) - torch.max(a_min, b_min), min=0)
    inter = bev * h
    va = (a[:, 3] * a[:, 4] * a[:, 5]).view(-1, 1)
    vb = (b[:, 3] * b[:, 4] * b[:, 5]).view(-1, 1)
    return inter / torch.clamp(va + vb - inter, min=1e-6)


def _positives(mask, pred_box, box_gt):
    """(weights (B*M,) f32, pred (B*M,7), gt (B*M,7)) with every slot kept (static shapes, no host sync): the slots
    without an object get a copy of their own prediction as target — finite IoU terms whose weight is 0."""
    w = mask.reshape(-1).float()
    p = pred_box.reshape(-1, pred_box.shape[-1])
    g = box_gt.reshape(-1, box_gt.shape[-1])
    g = torch.where(w.unsqueeze(1) > 0, g, p.detach())
    return w, p, g


class IouLoss(nn.Module):
    """L1 between the IoU head and 2*IoU3D(pred, gt)-1 on the positives (centernet_loss.py:70-96).  The reference
    returns zeros((1,)) when `mask.sum() == 0` (a host sync) and compacts the positives with a boolean index (another);
    here every slot is evaluated and weighted by the mask: the same value, shape-static."""

    def forward(self, iou_pred, mask, ind, box_pred, box_gt):
        w, p, g = _positives(mask, gather_feat(box_pred, ind), box_gt)
        pred = gather_feat(iou_pred, ind).reshape(-1, 1)
        target = 2 * boxes_aligned_iou3d(p, g) - 1
        return ((pred - target).abs() * w.unsqueeze(1)).sum() / (w.sum() + 1e-4)


_const_cache = {}


def device_const(values, device, dtype=torch.float32):
    """small constant tensors are uploaded once per device: a host->device copy of a Python list inside the step would
    break CUDA-graph capture of the training step"""
    key = (tuple(map(tuple, values)) if values and isinstance(values[0], (list, tuple)) else tuple(values), str(device), dtype)
    t = _const_cache.get(key)
    if t is None:
        t = _const_cache[key] = torch.tensor(values, dtype=dtype).to(device)
    return t


def _corners(center, dim):
    norm = device_const([[-0.5, -0.5], [-0.5, 0.5], [0.5, 0.5], [0.5, -0.5]], dim.device)
    return dim.view(-1, 1, 2) * norm.view(1, 4, 2) + center.view(-1, 1, 2)


def _aligned_terms(p, g):
    qc, gc = _corners(p[:, :2], p[:, 3:5]), _corners(g[:, :2], g[:, 3:5])
    i_max, i_min = torch.minimum(qc[:, 2], gc[:, 2]), torch.maximum(qc[:, 0], gc[:, 0])
    o_max, o_min = torch.maximum(qc[:, 2], gc[:, 2]), torch.minimum(qc[:, 0], gc[:, 0])
    vp, vg = p[:, 3] * p[:, 4] * p[:, 5], g[:, 3] * g[:, 4] * g[:, 5]
    top = torch.minimum(g[:, 2] + 0.5 * g[:, 5], p[:, 2] + 0.5 * p[:, 5])
    bot = torch.maximum(g[:, 2] - 0.5 * g[:, 5], p[:, 2] - 0.5 * p[:, 5])
    ih = torch.clamp(top - bot, min=0)
    inter = torch.clamp(i_max - i_min, min=0)
    v_inter = inter[:, 0] * inter[:, 1] * ih
    v_union = vg + vp - v_inter
    otop = torch.maximum(g[:, 2] + 0.5 * g[:, 5], p[:, 2] + 0.5 * p[:, 5])
    obot = torch.minimum(g[:, 2] - 0.5 * g[:, 5], p[:, 2] - 0.5 * p[:, 5])
    oh = torch.clamp(otop - obot, min=0)
    outer = torch.clamp(o_max - o_min, min=0)
    return v_inter, v_union, outer, oh


def bbox3d_overlaps_iou(p, g):
    """center_utils.py:124-154 (axis-aligned: the heading is ignored)"""
    v_inter, v_union, _, _ = _aligned_terms(p, g)
    return torch.clamp(v_inter / v_union, min=0, max=1.0)


def bbox3d_overlaps_giou(p, g):
    """center_utils.py:157-188"""
    v_inter, v_union, outer, oh = _aligned_terms(p, g)
    closure = outer[:, 0] * outer[:, 1] * oh
    return torch.clamp(v_inter / v_union - (closure - v_union) / closure, min=-1.0, max=1.0)


def bbox3d_overlaps_diou(p, g):
    """center_utils.py:191-228"""
    v_inter, v_union, outer, oh = _aligned_terms(p, g)
    inter_diag = (g[:, 0:3] - p[:, 0:3]).pow(2).sum(-1)
    outer_diag = outer[:, 0] ** 2 + outer[:, 1] ** 2 + oh ** 2
    return torch.clamp(v_inter / v_union - inter_diag / outer_diag, min=-1.0, max=1.0)


class IouRegLoss(nn.Module):
    """1 - (G/D)IoU of the decoded boxes on the positives (centernet_loss.py:99-125)"""

    def __init__(self, type="IoU"):
        super().__init__()
        funcs = {"IoU": bbox3d_overlaps_iou, "GIoU": bbox3d_overlaps_giou, "DIoU": bbox3d_overlaps_diou}
        if type not in funcs:
            raise NotImplementedError(type)
        self.bbox3d_iou_func = funcs[type]

    def forward(self, box_pred, mask, ind, box_gt):
        w, p, g = _positives(mask, gather_feat(box_pred, ind), box_gt)
        iou = self.bbox3d_iou_func(p, g)
        return ((1.0 - iou) * w).sum() / (w.sum() + 1e-4)
