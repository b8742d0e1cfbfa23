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


def candidate_agreement(head, plan_a, plan_b, score_thr=None, eps=1e-3):
    """Pre-NMS agreement of two CenterHead.predict_raw runs on the same frames (the score-sorted top-`pre_max`
    candidate boxes of every NMS segment, plan["sorted_boxes"] records [x y z w l h vx vy rot score rect label]):
    a candidate of run A is re-found in run B when B holds a candidate of the same class whose centre lies within a
    quarter of a head cell (i.e. the same heat-map pixel).  Reports the re-found fraction both ways, the BEV IoU of
    the matched pairs and their score difference.  This isolates the numeric difference of the two runs from the
    order sensitivity of greedy NMS (a near-tie in score reorders the sweep and changes which of two overlapping
    boxes survives — with random-init heads, whose scores sit within a hair of each other, that dominates the
    post-NMS comparison).

    A candidate that one run lists and the other does not is *explained* when its score lies within `eps` of the
    membership boundary of the other run's list — the score threshold, or that list's lowest score when the list is
    full (pre_max entries): the two runs then agree on the score (to within eps) and differ only in which side of the
    cut a near-tie falls.  `unexplained` counts the rest; the tests require it to be 0."""
    tot = dict(n_a=0, n_b=0, a_in_b=0, b_in_a=0, max_score_delta=0.0, min_iou=1.0, unexplained=0)
    iou_sum, iou_n = 0.0, 0
    ca, cb = plan_a["sorted_count"].tolist(), plan_b["sorted_count"].tolist()
    S = plan_a["S"]
    for seg in range(plan_a["sorted_boxes"].shape[0]):
        na, nb = min(ca[seg], plan_a["pre_cap"]), min(cb[seg], plan_b["pre_cap"])
        tot["n_a"] += na
        tot["n_b"] += nb
        if na == 0 or nb == 0:
            continue
        task = plan_a["segs"][seg % S]["task"]
        tol = 0.25 * head.task_strides[task] * float(head.pillar_size)
        A, B = plan_a["sorted_boxes"][seg, :na], plan_b["sorted_boxes"][seg, :nb]
        d = torch.cdist(A[:, :2].double(), B[:, :2].double())
        same = A[:, 11].view(-1, 1) == B[:, 11].view(1, -1)
        d = torch.where(same, d, torch.full_like(d, 1e9))
        da, ia = d.min(1)
        db, _ = d.min(0)
        ok = da <= tol
        tot["a_in_b"] += int(ok.sum())
        tot["b_in_a"] += int((db <= tol).sum())
        if score_thr is not None:
            pre = plan_a["segs"][seg % S]["pre"]
            cut_b = max(float(score_thr), float(B[:, 10].min())) if nb >= pre else float(score_thr)
            cut_a = max(float(score_thr), float(A[:, 10].min())) if na >= pre else float(score_thr)
            # sorted_boxes[:, 10] = the (rectified) score the list is ordered by
            tot["unexplained"] += int((A[~ok][:, 10] > cut_b + eps).sum()) + int((B[db > tol][:, 10] > cut_a + eps).sum())
        if bool(ok.any()):
            a7 = to_pcdet(A[ok][:, :9])
            b7 = to_pcdet(B[ia[ok]][:, :9])
            ov = ops.boxes_aligned_overlap_bev(a7, b7)
            area = a7[:, 3] * a7[:, 4] + b7[:, 3] * b7[:, 4] - ov
            # (nearly) coincident boxes: the fp32 polygon clipping can return slightly more than the true overlap when
            # edges coincide, which makes the union vanish — the ratio is capped at 1
            iou = (ov / area.clamp(min=1e-8)).clamp(max=1.0)
            iou_sum += float(iou.sum())
            iou_n += int(iou.numel())
            tot["min_iou"] = min(tot["min_iou"], float(iou.min()))
            tot["max_score_delta"] = max(tot["max_score_delta"], float((A[ok][:, 9] - B[ia[ok]][:, 9]).abs().max()))
    tot["recall_a_in_b"] = tot["a_in_b"] / max(1, tot["n_a"])
    tot["recall_b_in_a"] = tot["b_in_a"] / max(1, tot["n_b"])
    tot["mean_iou"] = iou_sum / max(1, iou_n)
    return tot


def rel_to_max(a, b):
    """max|a-b| / max(1, max|b|): the tolerance form the parity tests state (north_star: max-abs relative to fp32)."""
    return float((a.float() - b.float()).abs().max()) / max(1.0, float(b.float().abs().max()))
