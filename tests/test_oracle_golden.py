"""CPU: the oracle against the committed golden vectors (outputs of the reference's own code, generated
by tests/golden/make_golden.py) — this is what pins the oracle before it is trusted as the checker."""
import json
import os

import numpy as np

from oracle import pillarnet_oracle as O


def test_iou_oracle_bit_exact_vs_reference_cpu_twin(golden_dir):
    g = np.load(os.path.join(golden_dir, "iou_pairs.npz"))
    mine = O.boxes_iou_bev(g["a"], g["b"])
    assert (g["iou"] > 0).sum() > 100
    assert np.array_equal(mine.view(np.int32), g["iou"].view(np.int32))


def test_circle_nms_oracle_vs_reference_numba(golden_dir):
    g = np.load(os.path.join(golden_dir, "circle_nms.npz"))
    for i in range(3):
        d, thr, keep = g[f"dets{i}"], float(g[f"thr{i}"]), g[f"keep{i}"]
        order = O.stable_order_desc(d[:, 2])
        mine = order[O.nms_circle_sorted(d[order][:, :2], thr)]
        assert np.array_equal(mine, keep)
        assert 0 < len(keep) < len(d)


def test_predict_oracle_vs_reference_centerhead(golden_dir):
    """decode + post_processing (circular_nms) vs CenterHead.predict executed from /root/reference."""
    g = np.load(os.path.join(golden_dir, "head_predict_circle.npz"))
    ps, pcr = 0.075, [-54, -54, -5.0, 54, 54, 3.0]
    names = ["reg", "height", "dim", "rot", "vel", "hm"]
    ncls = [1, 2]
    min_radius = [4.0, 0.85]
    B = g["t0_hm"].shape[0]
    per_frame = [[] for _ in range(B)]
    for t in range(2):
        offs, parts, c = {}, [], 0
        for n in names:
            v = g[f"t{t}_{n}"].transpose(0, 2, 3, 1)
            offs[n] = c
            c += v.shape[-1]
            parts.append(v)
        maps = np.concatenate(parts, -1)
        boxes, hm, iou = O.decode_task(maps, offs, ncls[t], 8, ps, pcr)
        for b in range(B):
            cfg = dict(mode="circle", min_radius=min_radius[t], post_max=83, score_threshold=0.1,
                       post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0])
            bx, sc, lb = O.post_process_frame(boxes[b], hm[b], iou[b], cfg)
            per_frame[b].append((bx, sc, lb + sum(ncls[:t])))
    for b in range(B):
        bx = np.concatenate([p[0] for p in per_frame[b]])
        sc = np.concatenate([p[1] for p in per_frame[b]])
        lb = np.concatenate([p[2] for p in per_frame[b]])
        assert len(bx) == len(g[f"out{b}_boxes"]) > 10
        assert np.array_equal(lb, g[f"out{b}_labels"])
        np.testing.assert_allclose(sc, g[f"out{b}_scores"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(bx, g[f"out{b}_boxes"], rtol=1e-5, atol=1e-5)


def test_predict_double_flip_oracle_vs_reference_centerhead(golden_dir):
    """double-flip branch of predict (center_head.py:233-304) vs the reference executed from /root/reference."""
    g = np.load(os.path.join(golden_dir, "head_predict_double_flip.npz"))
    ps, pcr = 0.075, [-54, -54, -5.0, 54, 54, 3.0]
    names = ["reg", "height", "dim", "rot", "vel", "iou", "hm"]
    ncls = [1, 2]
    min_radius = [4.0, 0.85]
    B = g["t0_hm"].shape[0] // 4
    per_frame = [[] for _ in range(B)]
    for t in range(2):
        offs, parts, c = {}, [], 0
        for n in names:
            v = g[f"t{t}_{n}"].transpose(0, 2, 3, 1)
            offs[n] = c
            c += v.shape[-1]
            parts.append(v)
        boxes, hm, iou = O.decode_task(np.concatenate(parts, -1), offs, ncls[t], 8, ps, pcr, double_flip=True)
        assert boxes.shape[0] == B
        for b in range(B):
            cfg = dict(mode="circle", min_radius=min_radius[t], post_max=83, score_threshold=0.1,
                       post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0])
            bx, sc, lb = O.post_process_frame(boxes[b], hm[b], iou[b], cfg)
            per_frame[b].append((bx, sc, lb + sum(ncls[:t])))
    for b in range(B):
        bx = np.concatenate([p[0] for p in per_frame[b]])
        sc = np.concatenate([p[1] for p in per_frame[b]])
        lb = np.concatenate([p[2] for p in per_frame[b]])
        assert len(bx) == len(g[f"out{b}_boxes"]) > 10
        assert np.array_equal(lb, g[f"out{b}_labels"])
        np.testing.assert_allclose(sc, g[f"out{b}_scores"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(bx, g[f"out{b}_boxes"], rtol=1e-5, atol=1e-5)


def test_set_by_task_cfg_vs_reference(golden_dir):
    import pillarnet_lts_b200  # noqa: F401
    from pillarnet_lts_b200.detector import set_by_task_cfg
    with open(os.path.join(golden_dir, "set_by_task_cfg.json")) as fh:
        want = json.load(fh)
    cfg = dict(nms=dict(use_multi_class_nms=True, nms_pre_max_size=[2048, 1024, 1024],
                        nms_post_max_size=[200, 150, 150], nms_iou_threshold=[0.8, 0.55, 0.55]),
               rectifier=[0., 0., 0.], score_threshold=0.1,
               post_center_limit_range=[-80, -80, -10.0, 80, 80, 10.0])
    got = set_by_task_cfg(cfg, [1, 2])
    assert json.loads(json.dumps(got)) == want


def test_centerhead_loss_vs_reference(golden_dir):
    """CenterHead.loss (PyTorch restatement, losses.py) vs the reference's loss executed from /root/reference."""
    import torch
    from pillarnet_lts_b200.head import CenterHead
    g = np.load(os.path.join(golden_dir, "head_loss.npz"))
    tasks = [dict(stride=8, class_names=["car"]), dict(stride=8, class_names=["ped", "cone"])]
    head = CenterHead(tasks=tasks, in_channels=[16], code_weights=[1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.2, 0.2, 1.0, 1.0],
                      common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "vel": (2, 2)},
                      share_channel=8, reg_iou="GIoU", pillar_size=0.075, point_cloud_range=[-54, -54, -5.0, 54, 54, 3.0])
    preds = [{n: torch.from_numpy(g[f"t{t}_{n}"]).requires_grad_(True) for n in ["reg", "height", "dim", "rot", "vel", "hm"]}
             for t in range(2)]
    example = {k: [torch.from_numpy(g[f"ex{t}_{k}"]) for t in range(2)]
               for k in ("hm", "ind", "mask", "cat", "anno_box", "gt_box")}
    out = head.loss(example, preds, dict(hm_weight=1, bbox_weight=0.25, iou_weight=1, reg_iou_weight=0.25))
    for t in range(2):
        for k in ("loss", "hm_loss", "loc_loss", "loc_loss_elem", "reg_iou_loss", "num_positive"):
            np.testing.assert_allclose(out[k][t].detach().numpy().reshape(-1), g[f"out{t}_{k}"], rtol=1e-5, atol=1e-6,
                                       err_msg=f"{t} {k}")
    sum(out["loss"]).sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for d in preds for p in d.values())


def _assign_cases(g):
    """(frame, task) -> per-task boxes / 1-based class ids in the reference's order (preprocess.py:203-237)"""
    tasks = [dict(stride=8, names=[1]), dict(stride=4, names=[2, 3])]
    for f in range(3):
        boxes, cls = g[f"f{f}_boxes"], g[f"f{f}_cls"]
        for t, task in enumerate(tasks):
            sel = np.concatenate([np.where(cls == c)[0] for c in task["names"]])
            tb = boxes[sel].copy()
            tb[:, -1] = tb[:, -1] - np.floor(tb[:, -1] / (np.pi * 2) + 0.5) * (np.pi * 2)   # limit_period
            tc = np.concatenate([np.full((cls == c).sum(), j + 1) for j, c in enumerate(task["names"])])
            yield f, t, task, tb.astype(np.float32), tc.astype(np.int32)


def test_assign_label_oracle_vs_reference(golden_dir):
    """numpy restatement of AssignLabel (one task, one frame) vs the reference pipeline stage run from /root/reference"""
    g = np.load(os.path.join(golden_dir, "assign_label.npz"))
    M, ps = 80, np.float32(0.075)
    positives = 0
    for f, t, task, tb, tc in _assign_cases(g):
        boxes = np.zeros((M, 9), np.float32)
        cls = np.zeros(M, np.int32)
        boxes[:len(tb)], cls[:len(tc)] = tb, tc
        grid = 640 // task["stride"]
        r = O.assign_labels_task(boxes, cls, len(task["names"]), grid, grid, -24.0, -24.0, np.float32(ps * task["stride"]),
                                 0.1, 2)
        for k in ("ind", "mask", "cat"):
            assert np.array_equal(r[k], g[f"f{f}_t{t}_{k}"]), (f, t, k)
        assert np.array_equal(r["hm"], g[f"f{f}_t{t}_hm"]), (f, t)                 # bit-exact heat-map
        assert np.array_equal(r["gt_box"], g[f"f{f}_t{t}_gt_box"])
        np.testing.assert_allclose(r["anno_box"], g[f"f{f}_t{t}_anno_box"], rtol=1e-6, atol=1e-7)
        positives += int(r["mask"].sum())
    assert positives > 50


def _sweep_case(g):
    key = g["raw0"]
    sweeps = []
    for i in g["order"]:                                # the order the reference drew (np.random.choice, seed 3)
        k = int(i) + 1
        T = g[f"T{k}"]
        sweeps.append((g[f"raw{k}"], None if np.isnan(T[0, 0]) else T, float(g[f"lag{k}"])))
    return key, sweeps


def test_merge_sweeps_oracle_vs_reference(golden_dir):
    """numpy restatement of the multi-sweep loader vs the reference's LoadPointCloudFromFile run on seeded files"""
    g = np.load(os.path.join(golden_dir, "sweeps.npz"))
    key, sweeps = _sweep_case(g)
    got = O.merge_sweeps(key, sweeps)
    assert got.dtype == np.float32 and np.array_equal(got, g["combined"])
    assert len(got) < sum(len(g[f"raw{k}"]) for k in range(5))        # remove_close dropped points
