"""TEST INFRASTRUCTURE ONLY — the whole point->detections path on the host CPU (oracle "port").

Used by bench.py's `cpu_baseline` leg / `--impl reference` arm and by tests as an end-to-end checker.
It restates the reference's CPU-runnable formulation of the path:
  * pillarization: pillarnet_oracle.pillarize (dynamic_pillar_encoder.py:29-47, pillar_utils.py:34-56)
  * PFN: torch-CPU Linear -> BatchNorm1d(eval) -> ReLU + amax scatter (pillar_modules.py:26-33,71-72)
  * sparse backbone: dense-equivalent masked F.conv2d (base.py:145-213,
    PillarResNet.py:134-149) — spconv itself has no CPU path and is not installable (parity unpinned)
  * dense conv5 / neck / head: the model's own nn.Conv2d / BatchNorm2d containers run by torch on CPU
    (necks/rpn.py:193-207,330-355; center_head.py:116-127)
  * decode + NMS: pillarnet_oracle.decode_task / post_process_frame (center_head.py:216-413)
The weights come from the model under test (a CPU copy), so outputs are comparable with the GPU path.
"""
import time

import numpy as np
import torch
import torch.nn.functional as F

from . import pillarnet_oracle as O


def _bn_eval(x, bn):
    return F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps)


def backbone_dense_equivalent(backbone, x, mask):
    def subm(seq, x, mask, relu, res=None):
        conv, bn = seq[0], seq[1]
        y = _bn_eval(F.conv2d(x, conv.weight.permute(0, 3, 1, 2), conv.bias, padding=1), bn)
        if res is not None:
            y = y + res
        if relu:
            y = F.relu(y)
        return y * mask

    def block(b, x, mask):
        if hasattr(b, "conv0"):
            x = subm(b.conv0, x, mask, False)
        out = subm(b.conv1, x, mask, True)
        return subm(b.conv2, out, mask, True, res=x)

    feats = {}
    for name in ("conv1", "conv2", "conv3", "conv4"):
        mods = list(getattr(backbone, name))
        i = 0
        if not hasattr(mods[0], "conv1"):
            conv, bn = mods[0], mods[1]
            mask = (F.max_pool2d(mask, 3, 2, 1) > 0).float()
            x = F.relu(_bn_eval(F.conv2d(x, conv.weight.permute(0, 3, 1, 2), None, stride=2, padding=1), bn)) * mask
            i = 3
        for b in mods[i:]:
            x = block(b, x, mask)
        feats[name] = x
    if hasattr(backbone, "conv5"):
        feats["conv5"] = backbone.conv5(feats["conv4"])
    return feats


def neck_forward(neck, feats):
    name = type(neck).__name__
    if name == "RPNV1":
        up = neck.deblock_5(neck.block_5(feats["conv5"]))
        return (neck.block_4(torch.cat([feats["conv4"], up], 1)),)
    if name == "RPNG":
        x5 = neck.block_5(feats["conv5"])
        x4 = neck.block_4(torch.cat([feats["conv4"], neck.top_down_54(x5)], 1))
        x3 = neck.block_3(torch.cat([feats["conv3"], neck.top_down_43(x4)], 1))
        return (x4, x3)
    if name == "RPNV2":       # necks/rpn.py:262-272 (reads the sparse conv3 / conv4 of the S backbones, densified)
        up = neck.deblock_4(neck.block_4(feats["conv4"]))
        return (neck.block_3(torch.cat([feats["conv3"], up], 1)),)
    if name == "RPNGV2":      # necks/rpn.py:428-450
        x5 = neck.block_5(feats["conv5"])
        x4 = neck.block_4(torch.cat([neck.reduce_4(feats["conv4"]), neck.top_down_54(x5)], 1))
        x3 = neck.block_3(torch.cat([neck.reduce_3(feats["conv3"]), neck.top_down_43(x4)], 1))
        return (x4, x3)
    raise NotImplementedError(name)


def head_forward(head, bev):
    """center_head.py:116-127 with the model's own torch containers"""
    share = [sc(bev[k]) for k, sc in enumerate(head.share_convs)]
    return [{name: getattr(th, name)(share[head.task_idx[t]]) for name in th.heads}
            for t, th in enumerate(head.task_heads)]


@torch.no_grad()
def dense_equivalent_from_reader(model, sp):
    """Dense-equivalent torch forward (fp32, TF32 off) from a reader output `sp` (features_f32, indices) of the model
    under test to its head maps, on sp's device: the per-stage checker of the -m gpu whole-model parity tests."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B, (H, W) = sp.batch_size, sp.spatial_shape
    n = sp.table.count()
    idx = sp.indices.long()
    f = sp.features_f32[:n] if getattr(sp, "features_f32", None) is not None else sp.feat[:n].float()
    x = torch.zeros(B, f.shape[1], H, W, device=f.device)
    x[idx[:, 0], :, idx[:, 1], idx[:, 2]] = f
    mask = torch.zeros(B, 1, H, W, device=f.device)
    mask[idx[:, 0], 0, idx[:, 1], idx[:, 2]] = 1
    feats = backbone_dense_equivalent(model.backbone, x, mask)
    del x
    bev = neck_forward(model.neck, feats)
    preds = head_forward(model.bbox_head, bev)
    return feats, bev, preds


def nms_cfg_for_task(head, test_cfg, t):
    nms = test_cfg["nms"]
    base = dict(score_threshold=test_cfg["score_threshold"],
                post_center_limit_range=test_cfg["post_center_limit_range"])
    if test_cfg.get("circular_nms", False):
        post = nms["nms_post_max_size"]
        base.update(mode="circle", min_radius=test_cfg["min_radius"][t],
                    post_max=post[t] if isinstance(post, (list, tuple)) else post)
    elif nms.get("use_rotate_nms", False):
        base.update(mode="rotate", rectifier=test_cfg.get("rectifier", 0), thr=nms["nms_iou_threshold"],
                    pre_max=nms["nms_pre_max_size"], post_max=nms["nms_post_max_size"])
    else:
        base.update(mode="multi_class", rectifiers=test_cfg["rectifier"][t], thrs=nms["nms_iou_threshold"][t],
                    pre_max=nms["nms_pre_max_size"][t], post_max=nms["nms_post_max_size"][t])
    return base


_p2v = None


@torch.no_grad()
def numba_reader(rd, frames):
    """BASELINE.json config 1 — the reference's CPU reader: numba `points_to_voxel` (hard voxelisation, z collapsed to
    one cell; det3d/ops/point_cloud/point_cloud_ops.py:112-184, executed from the copy oracle/build_ref.py stages;
    max_points 64, max_voxels 150000 as SURVEY §8d) per frame, then the PFN (pillar-centre offsets as
    pillar_utils.py:51-56, Linear+BN1d+ReLU, pillar_modules.py:26-33, max over the pillar's points) on torch-CPU.
    Returns ((pillar_features (M,C), pillar_indices (M,3) [b,y,x] int64, H, W), timings).  NOT the dynamic-pillar
    semantics of the GPU path (caps, first-come order): a timing baseline only."""
    global _p2v
    if _p2v is None:
        from . import build_ref
        _p2v = build_ref.load_points_to_voxel()
    pcr, ps = rd.point_cloud_range, rd.pillar_size
    t0 = time.perf_counter()
    vox, coors, nump, bidx = [], [], [], []
    for b, f in enumerate(frames):
        v, c, n = _p2v(np.ascontiguousarray(f), [ps, ps, pcr[5] - pcr[2]], list(map(float, pcr)), 64,
                       max_voxels=150000)
        vox.append(v), coors.append(c), nump.append(n), bidx.append(np.full(len(c), b, np.int64))
    t1 = time.perf_counter()
    v = torch.from_numpy(np.concatenate(vox))                      # (M, 64, D)
    c = torch.from_numpy(np.concatenate(coors).astype(np.int64))   # (M, 3) [z, y, x]
    n = torch.from_numpy(np.concatenate(nump).astype(np.int64))
    M = v.shape[0]
    valid = torch.arange(v.shape[1]).view(1, -1) < n.view(-1, 1)
    pid = torch.arange(M).view(-1, 1).expand_as(valid)[valid]      # point -> pillar
    pts = v[valid]                                                 # (L, D)
    cx = c[pid, 2].float() * ps + rd.x_offset
    cy = c[pid, 1].float() * ps + rd.y_offset
    feat = torch.cat([(pts[:, 0] - cx).unsqueeze(1), (pts[:, 1] - cy).unsqueeze(1), pts], 1)
    h = rd.shared_mlps(feat)
    pf = torch.zeros(M, h.shape[1])
    pf.index_reduce_(0, pid, h, "amax", include_self=True)
    t2 = time.perf_counter()
    idx = torch.stack([torch.from_numpy(np.concatenate(bidx)), c[:, 1], c[:, 2]], 1)
    return (pf, idx, rd.height, rd.width), dict(points_to_voxel_s=t1 - t0, pfn_s=t2 - t1)


@torch.no_grad()
def cpu_forward(model_cpu, frames, timings=None, pillarizer="numpy"):
    """frames: list of (Ni,5) float32 numpy. Returns det3d-style detections (numpy) per frame.
    pillarizer: "numpy" = the dynamic-pillar oracle (comparable with the GPU path); "numba" = the reference's own
    CPU voxeliser (bench.py's cpu_baseline: the CPU path BASELINE.json names)."""
    t0 = time.perf_counter()
    rd = model_cpu.reader.pfn_layers
    pcr, ps = rd.point_cloud_range, rd.pillar_size
    if pillarizer == "numba":
        (pf, pi, H, W), tr = numba_reader(rd, frames)
        t1 = t0 + tr["points_to_voxel_s"]
        t2 = time.perf_counter()
        B = len(frames)
        r = dict(H=H, W=W, pillar_indices=pi.numpy())
    else:
        r = O.pillarize(frames, pcr, ps, mode="cuda")
        feat = O.point_pillar_features(r["pts"], r["pts_xy"], pcr, ps)
        t1 = time.perf_counter()
        h = rd.shared_mlps(torch.from_numpy(feat))
        M = len(r["pillar_indices"])
        pf = torch.zeros(M, h.shape[1])
        idx = torch.from_numpy(r["point_pillar_indices"].astype(np.int64))
        pf.index_reduce_(0, idx, h, "amax", include_self=True)
        t2 = time.perf_counter()
        B, H, W = len(frames), r["H"], r["W"]
        pi = torch.from_numpy(r["pillar_indices"].astype(np.int64))
    x = torch.zeros(B, pf.shape[1], H, W)
    x[pi[:, 0], :, pi[:, 1], pi[:, 2]] = pf
    mask = torch.zeros(B, 1, H, W)
    mask[pi[:, 0], 0, pi[:, 1], pi[:, 2]] = 1
    feats = backbone_dense_equivalent(model_cpu.backbone, x, mask)
    t3 = time.perf_counter()
    bev = neck_forward(model_cpu.neck, feats)
    head = model_cpu.bbox_head
    share = [sc(bev[k]) for k, sc in enumerate(head.share_convs)]
    preds = []
    for t, th in enumerate(head.task_heads):
        preds.append({name: getattr(th, name)(share[head.task_idx[t]]) for name in th.heads})
    t4 = time.perf_counter()
    test_cfg = model_cpu.test_cfg
    per_frame = [[] for _ in range(B)]
    flag = 0
    for t, p in enumerate(preds):
        offs, parts, c = {}, [], 0
        for name, v in p.items():
            offs[name] = c
            c += v.shape[1]
            parts.append(v.permute(0, 2, 3, 1).numpy())
        maps = np.concatenate(parts, -1)
        boxes, hm, iou = O.decode_task(maps, offs, head.num_classes[t], head.task_strides[t], head.pillar_size,
                                       head.point_cloud_range)
        cfg = nms_cfg_for_task(head, test_cfg, t)
        for b in range(B):
            bx, sc, lb = O.post_process_frame(boxes[b], hm[b], iou[b], cfg)
            per_frame[b].append((bx, sc, lb + flag))
        flag += head.num_classes[t]
    t5 = time.perf_counter()
    out = []
    for b in range(B):
        out.append({"box3d_lidar": np.concatenate([p[0] for p in per_frame[b]]),
                    "scores": np.concatenate([p[1] for p in per_frame[b]]),
                    "label_preds": np.concatenate([p[2] for p in per_frame[b]])})
    if timings is not None:
        timings.update(pillarize_s=t1 - t0, pfn_s=t2 - t1, backbone_s=t3 - t2, neck_head_s=t4 - t3,
                       decode_nms_s=t5 - t4, total_s=t5 - t0)
    return out, dict(feats=feats, bev=bev, preds=preds, pillars=r)
