"""Detection-level agreement between two runs of the same frames (e.g. the bf16 tensor-core mode against the
fp32 mode): what a user of the detector sees of a numeric difference.

For every detection of run A the best BEV IoU among run B's detections of the same class is taken
(pn_boxes_iou_bev, the reference's `boxes_iou_bev_gpu` arithmetic, iou3d_nms_kernel.cu:236-249, on boxes
converted with `to_pcdet`, iou3d_nms_utils.py:30-34).  A pair counts as matched at IoU >= iou_thr.
bench.py prints the summary as its `parity` block; tests/test_gpu_fullsize.py asserts on it.
"""
import math

import torch

from . import ops


def to_pcdet(boxes):
    """(n, 7|9) det3d boxes [x,y,z,w,l,h,(vx,vy),rot] -> (n,7) [x,y,z,l,w,h,-rot-pi/2]."""
    b = boxes[:, [0, 1, 2, 4, 3, 5, -1]].clone().float()
    b[:, -1] = -b[:, -1] - math.pi / 2
    return b.contiguous()


def frame_agreement(det_a, det_b, iou_thr=0.7, top=None):
    """det_*: {"box3d_lidar", "scores", "label_preds"} of one frame (CUDA tensors).
    Returns dict(n_a, n_b, matched, max_score_delta, mean_iou) — matched counts A's boxes with a same-class
    partner in B at BEV IoU >= iou_thr; `top` restricts A to its `top` highest scores."""
    ba, sa, la = det_a["box3d_lidar"], det_a["scores"], det_a["label_preds"]
    bb, sb, lb = det_b["box3d_lidar"], det_b["scores"], det_b["label_preds"]
    if top is not None and sa.numel() > top:
        idx = torch.topk(sa, top).indices
        ba, sa, la = ba[idx], sa[idx], la[idx]
    n_a, n_b = int(sa.numel()), int(sb.numel())
    if n_a == 0 or n_b == 0:
        return dict(n_a=n_a, n_b=n_b, matched=0, max_score_delta=0.0, mean_iou=0.0)
    iou = ops.boxes_iou_bev(to_pcdet(ba.cuda()), to_pcdet(bb.cuda()))
    same = la.cuda().view(-1, 1) == lb.cuda().view(1, -1)
    iou = torch.where(same, iou, torch.zeros_like(iou))
    best, arg = iou.max(1)
    ok = best >= iou_thr
    delta = (sa.cuda() - sb.cuda()[arg]).abs()
    return dict(n_a=n_a, n_b=n_b, matched=int(ok.sum()),
                max_score_delta=float(delta[ok].max()) if bool(ok.any()) else 0.0,
                mean_iou=float(best[ok].mean()) if bool(ok.any()) else 0.0)


def summarize(dets_a, dets_b, iou_thr=0.7, top=None):
    """Agreement over a list of frames: recall of A in B, recall of B in A, worst matched score delta."""
    tot = dict(n_a=0, n_b=0, a_in_b=0, b_in_a=0, max_score_delta=0.0)
    ious = []
    for da, db in zip(dets_a, dets_b):
        f = frame_agreement(da, db, iou_thr, top)
        g = frame_agreement(db, da, iou_thr, top)
        tot["n_a"] += f["n_a"]
        tot["n_b"] += g["n_a"]
        tot["a_in_b"] += f["matched"]
        tot["b_in_a"] += g["matched"]
        tot["max_score_delta"] = max(tot["max_score_delta"], f["max_score_delta"], g["max_score_delta"])
        if f["matched"]:
            ious.append(f["mean_iou"])
    tot["recall_a_in_b"] = tot["a_in_b"] / max(1, tot["n_a"])
    tot["recall_b_in_a"] = tot["b_in_a"] / max(1, tot["n_b"])
    tot["mean_matched_iou"] = float(sum(ious) / len(ious)) if ious else 0.0
    tot["iou_thr"] = iou_thr
    return tot


def rel_to_max(a, b):
    """max|a-b| / max(1, max|b|): the tolerance form the parity tests state (north_star: max-abs relative to fp32)."""
    return float((a.float() - b.float()).abs().max()) / max(1.0, float(b.float().abs().max()))
