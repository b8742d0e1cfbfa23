"""Generates the committed golden vectors by EXECUTING the reference's own code in the build container.

Run:  python tests/golden/make_golden.py      (needs /root/reference and oracle/_ref; see _ref_import.py)

Fixtures (all small .npz):
  iou_pairs.npz          det3d/ops/iou3d_nms/src/iou3d_cpu.cpp:232-250 boxes_iou_bev_cpu on seeded boxes
  circle_nms.npz         det3d/core/utils/circle_nms_jit.py:4-28 (numba) keep lists
  head_predict_circle.npz det3d/models/bbox_heads/center_head.py:216-413 CenterHead.predict (circular_nms)
  head_predict_double_flip.npz center_head.py:233-304 double-flip test-time augmentation branch of predict
  head_loss.npz          center_head.py:133-214 + losses/centernet_loss.py CenterHead.loss (focal + L1 + GIoU)
  assign_label.npz       datasets/pipelines/preprocess.py:177-350 AssignLabel (heat-maps, ind/mask/cat, anno_box, gt_box)
  sweeps.npz             datasets/pipelines/loading.py:102-141 key frame + sweeps (remove_close, transform, time lag)
  neck_head_forward.npz  det3d/models/necks/rpn.py:137-207 RPNV1 + center_head.py:116-127 forward (torch CPU)
  set_by_task_cfg.json   det3d/core/utils/center_utils.py:229-274 on the Waymo FPN test_cfg
  second_stage.npz       det3d/models/second_stage/bev_interpolation.py:162-308 BEVStrideFeature (two variants: the shipped
                         config's stride-1 laterals, and out_stride 2 with stride-2 transposed convs),
                         roi_heads/roi_mix_head.py:83-122 RoIMIXHead.forward(training=False),
                         point_heads/point_head_simple.py:68-96 PointHead.forward,
                         detectors/pillar_rcnn.py:141-170 PillarRCNN.post_process (torch CPU)
"""
import json
import logging
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import  # noqa: E402

_ref_import.setup()

from det3d.core.utils.circle_nms_jit import circle_nms  # noqa: E402
from det3d.core.utils.center_utils import set_by_task_cfg  # noqa: E402
from det3d.models.bbox_heads.center_head import CenterHead  # noqa: E402
from det3d.models.necks.rpn import RPNV1  # noqa: E402
from det3d.torchie import Config  # noqa: E402
import det3d.ops.iou3d_nms.iou3d_nms_cuda as iou3d  # noqa: E402


def rand_boxes(rng, n, spread=10.0):
    b = np.zeros((n, 7), np.float32)
    b[:, 0:2] = rng.uniform(-spread, spread, (n, 2))
    b[:, 2] = rng.uniform(-1, 1, n)
    b[:, 3] = rng.uniform(0.3, 5, n)
    b[:, 4] = rng.uniform(0.3, 2.5, n)
    b[:, 5] = rng.uniform(0.5, 2, n)
    b[:, 6] = rng.uniform(-4, 4, n)
    return b


def gen_iou():
    rng = np.random.default_rng(11)
    A = rand_boxes(rng, 96)
    B = rand_boxes(rng, 96)
    B[:32] = A[:32] + rng.normal(0, 0.05, (32, 7)).astype(np.float32)       # near duplicates
    B[32:40] = A[32:40]                                                      # identical
    A[40:56, 3:5] = rng.uniform(0.4, 0.9, (16, 2))                           # pedestrian-sized
    B[40:56] = A[40:56] + rng.normal(0, 0.02, (16, 7)).astype(np.float32)
    B[56:60, 6] = A[56:60, 6] + np.float32(np.pi / 2)                        # axis swaps
    B[56:60, :2] = A[56:60, :2]
    out = torch.zeros(len(A), len(B))
    iou3d.boxes_iou_bev_cpu(torch.from_numpy(A), torch.from_numpy(B), out)
    np.savez_compressed(os.path.join(HERE, "iou_pairs.npz"), a=A, b=B, iou=out.numpy())


def gen_circle():
    rng = np.random.default_rng(12)
    cases = {}
    for i, (n, thr) in enumerate([(200, 4.0), (500, 0.85), (64, 0.175)]):
        d = np.zeros((n, 3), np.float32)
        d[:, :2] = rng.uniform(-20, 20, (n, 2))
        d[:, 2] = rng.permutation(n).astype(np.float32) / n * 0.9 + 0.1   # distinct scores
        keep = np.array(circle_nms(d, thresh=thr), np.int64)
        cases[f"dets{i}"], cases[f"thr{i}"], cases[f"keep{i}"] = d, np.float64(thr), keep
    np.savez_compressed(os.path.join(HERE, "circle_nms.npz"), **cases)


def gen_predict():
    rng = np.random.default_rng(13)
    tasks = [dict(stride=8, class_names=["car"]), dict(stride=8, class_names=["ped", "cone"])]
    ps, pcr = 0.075, [-54, -54, -5.0, 54, 54, 3.0]
    head = CenterHead(tasks=[Config(t) for t in tasks], in_channels=[16], code_weights=[1.0] * 10,
                      common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "vel": (2, 2)},
                      share_channel=8, pillar_size=ps, point_cloud_range=pcr, logger=logging.getLogger("g"))
    B, H, W = 2, 40, 40
    test_cfg = Config(dict(circular_nms=True, min_radius=[4.0, 0.85],
                           nms=dict(nms_pre_max_size=[1000, 1000], nms_post_max_size=[83, 83], nms_iou_threshold=0.2),
                           score_threshold=0.1, post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0]))
    preds, save = [], {}
    for t, task in enumerate(tasks):
        K = len(task["class_names"])
        d = {}
        for name, c in (("reg", 2), ("height", 1), ("dim", 3), ("rot", 2), ("vel", 2), ("hm", K)):
            v = rng.normal(0, 0.7, (B, c, H, W)).astype(np.float32)
            if name == "hm":
                v = rng.normal(-3.5, 1.5, (B, c, H, W)).astype(np.float32)
            if name == "reg":
                v = rng.uniform(0, 1, (B, c, H, W)).astype(np.float32)
            d[name] = torch.from_numpy(v)
            save[f"t{t}_{name}"] = v
        preds.append(d)
    rets = head.predict({"metadata": [None] * B}, [dict((k, v.clone()) for k, v in p.items()) for p in preds], test_cfg)
    for b, r in enumerate(rets):
        save[f"out{b}_boxes"] = r["box3d_lidar"].numpy()
        save[f"out{b}_scores"] = r["scores"].numpy()
        save[f"out{b}_labels"] = r["label_preds"].numpy()
    np.savez_compressed(os.path.join(HERE, "head_predict_circle.npz"), **save)


def gen_predict_double_flip():
    """center_head.py:233-304,319-323: 4 flipped views per frame (2 output frames), circular NMS.
    The maps carry an 'iou' head: without one the reference's own double-flip branch raises IndexError
    (its stand-in torch.ones((B,4,H)) loses the W axis at :265-266, then `ious[mask]` fails at :376)."""
    rng = np.random.default_rng(15)
    tasks = [dict(stride=8, class_names=["car"]), dict(stride=8, class_names=["ped", "cone"])]
    ps, pcr = 0.075, [-54, -54, -5.0, 54, 54, 3.0]
    head = CenterHead(tasks=[Config(t) for t in tasks], in_channels=[16], code_weights=[1.0] * 10,
                      common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "vel": (2, 2)},
                      share_channel=8, pillar_size=ps, point_cloud_range=pcr, logger=logging.getLogger("g"))
    Bo, H, W = 2, 36, 44
    test_cfg = Config(dict(circular_nms=True, min_radius=[4.0, 0.85], double_flip=True,
                           nms=dict(nms_pre_max_size=[1000, 1000], nms_post_max_size=[83, 83], nms_iou_threshold=0.2),
                           score_threshold=0.1, post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0]))
    preds, save = [], {}
    for t, task in enumerate(tasks):
        K = len(task["class_names"])
        d = {}
        for name, c in (("reg", 2), ("height", 1), ("dim", 3), ("rot", 2), ("vel", 2), ("iou", 1), ("hm", K)):
            v = rng.normal(0, 0.7, (Bo, 4, c, H, W)).astype(np.float32)
            if name == "hm":
                base = rng.normal(-3.5, 1.8, (Bo, 1, c, H, W)).astype(np.float32)
                v = np.repeat(base, 4, 1) + rng.normal(0, 0.3, (Bo, 4, c, H, W)).astype(np.float32)
                v[:, 1] = v[:, 1, :, ::-1]          # the network sees the flipped scene
                v[:, 2] = v[:, 2, :, :, ::-1]
                v[:, 3] = v[:, 3, :, ::-1, ::-1]
            if name == "reg":
                v = rng.uniform(0, 1, (Bo, 4, c, H, W)).astype(np.float32)
            v = np.ascontiguousarray(v.reshape(Bo * 4, c, H, W))
            d[name] = torch.from_numpy(v)
            save[f"t{t}_{name}"] = v
        preds.append(d)
    rets = head.predict({"metadata": [None] * (4 * Bo)}, [dict((k, v.clone()) for k, v in p.items()) for p in preds],
                        test_cfg)
    assert len(rets) == Bo
    for b, r in enumerate(rets):
        save[f"out{b}_boxes"] = r["box3d_lidar"].numpy()
        save[f"out{b}_scores"] = r["scores"].numpy()
        save[f"out{b}_labels"] = r["label_preds"].numpy()
        assert r["scores"].numel() > 5
    np.savez_compressed(os.path.join(HERE, "head_predict_double_flip.npz"), **save)


def gen_loss():
    """center_head.py:133-214 CenterHead.loss (FastFocalLoss + RegLoss + GIoU IouRegLoss, nuScenes head set) on CPU."""
    rng = np.random.default_rng(16)
    tasks = [dict(stride=8, class_names=["car"]), dict(stride=8, class_names=["ped", "cone"])]
    ps, pcr = 0.075, [-54, -54, -5.0, 54, 54, 3.0]
    cw = [1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.2, 0.2, 1.0, 1.0]
    head = CenterHead(tasks=[Config(t) for t in tasks], in_channels=[16], code_weights=cw,
                      common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "vel": (2, 2)},
                      share_channel=8, reg_iou="GIoU", pillar_size=ps, point_cloud_range=pcr,
                      logger=logging.getLogger("g"))
    B, H, W, M = 2, 16, 20, 12
    train_cfg = Config(dict(hm_weight=1, bbox_weight=0.25, iou_weight=1, reg_iou_weight=0.25))
    preds, example, save = [], {k: [] for k in ("hm", "ind", "mask", "cat", "anno_box", "gt_box")}, {}
    for t, task in enumerate(tasks):
        K = len(task["class_names"])
        d = {}
        for name, c in (("reg", 2), ("height", 1), ("dim", 3), ("rot", 2), ("vel", 2), ("hm", K)):
            v = rng.normal(0, 0.7, (B, c, H, W)).astype(np.float32)
            if name == "hm":
                v = rng.normal(-2.0, 1.5, (B, c, H, W)).astype(np.float32)
            d[name] = torch.from_numpy(v)
            save[f"t{t}_{name}"] = v
        preds.append(d)
        hm = rng.uniform(0, 1, (B, H, W, K)).astype(np.float32) ** 4   # (B,H,W,C): preprocess.py:317
        ind = np.stack([rng.choice(H * W, M, replace=False) for _ in range(B)]).astype(np.int64)
        mask = (rng.uniform(0, 1, (B, M)) < 0.7).astype(np.uint8)
        mask[1, :] = mask[1, :] if t == 0 else 0              # one frame without objects in task 1
        cat = rng.integers(0, K, (B, M)).astype(np.int64)
        for b in range(B):
            for m in range(M):
                if mask[b, m]:
                    hm[b, ind[b, m] // W, ind[b, m] % W, cat[b, m]] = 1.0
        anno = rng.normal(0, 0.5, (B, M, 10)).astype(np.float32)
        gt = np.zeros((B, M, 7), np.float32)
        gt[..., 0:2] = rng.uniform(-50, 50, (B, M, 2))
        gt[..., 2] = rng.uniform(-2, 1, (B, M))
        gt[..., 3:6] = rng.uniform(0.5, 5, (B, M, 3))
        gt[..., 6] = rng.uniform(-3, 3, (B, M))
        for k, v in (("hm", hm), ("ind", ind), ("mask", mask), ("cat", cat), ("anno_box", anno), ("gt_box", gt)):
            example[k].append(torch.from_numpy(v))
            save[f"ex{t}_{k}"] = v
    out = head.loss(example, [dict((k, v.clone()) for k, v in p.items()) for p in preds], train_cfg)
    for t in range(2):
        for k in ("loss", "hm_loss", "loc_loss", "loc_loss_elem", "reg_iou_loss", "num_positive"):
            save[f"out{t}_{k}"] = np.asarray(out[k][t].detach().numpy(), np.float32).reshape(-1)
    np.savez_compressed(os.path.join(HERE, "head_loss.npz"), **save)


def _load_reference_preprocess():
    """det3d/datasets/pipelines/preprocess.py loaded as a stand-alone module: the package __init__ chain pulls in
    dataset code that needs pyquaternion / Python < 3.10, so its few imports are stubbed."""
    import importlib.util
    import types

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Reg:
        def register_module(self, cls):
            return cls

    stub("det3d.builder", build_dbsampler=lambda *a, **k: None)
    stub("det3d.datasets").__path__ = []
    stub("det3d.datasets.registry", PIPELINES=_Reg())
    stub("det3d.datasets.pipelines").__path__ = []
    import det3d.core.bbox.box_np_ops  # noqa: F401
    stub("det3d.core.sampler").__path__ = []
    stub("det3d.core.sampler.preprocess")
    spec = importlib.util.spec_from_file_location("det3d.datasets.pipelines.preprocess",
                                                  "/root/reference/det3d/datasets/pipelines/preprocess.py")
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "det3d.datasets.pipelines"
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


def gen_assign_label():
    """datasets/pipelines/preprocess.py:177-350 AssignLabel on seeded nuScenes-style annotations (3 frames, 2 tasks
    with strides 8 and 4, objects on the border and outside the range; the reference asserts total objects <= max_objs)."""
    mod = _load_reference_preprocess()
    rng = np.random.default_rng(17)
    tasks = [dict(stride=8, class_names=["car"]), dict(stride=4, class_names=["ped", "cone"])]
    cfg = Config(dict(target_assigner=dict(tasks=tasks), gaussian_overlap=0.1, max_objs=80, min_radius=2,
                      pc_range=[-24.0, -24.0, -5.0, 24.0, 24.0, 3.0], pillar_size=0.075))
    al = mod.AssignLabel(cfg=cfg)
    save = {}
    for f, n in enumerate([25, 70, 3]):
        boxes = np.zeros((n, 9), np.float32)
        boxes[:, 0:2] = rng.uniform(-26, 26, (n, 2))              # some outside the range
        boxes[:3, 0] = [-24.0, 23.99, -24.04][:min(3, n)]         # on / just outside the border
        boxes[:, 2] = rng.uniform(-2, 1, n)
        boxes[:, 3:6] = rng.uniform(0.3, 6.0, (n, 3))
        boxes[:, 6:8] = rng.normal(0, 2, (n, 2))
        boxes[:, 8] = rng.uniform(-7, 7, n)                       # headings beyond +-pi: limit_period
        cls = rng.integers(1, 4, n).astype(np.int32)
        names = np.array([["car", "ped", "cone"][c - 1] for c in cls])
        res = {"type": "NuScenesDataset",
               "lidar": {"annotations": {"gt_boxes": boxes.copy(), "gt_classes": cls.copy(), "gt_names": names}}}
        res, _ = al(res, {})
        tg = res["lidar"]["targets"]
        save[f"f{f}_boxes"], save[f"f{f}_cls"] = boxes, cls
        for t in range(2):
            for k in ("hm", "anno_box", "ind", "mask", "cat", "gt_box"):
                save[f"f{f}_t{t}_{k}"] = np.ascontiguousarray(tg[k][t])
    np.savez_compressed(os.path.join(HERE, "assign_label.npz"), **save)


def gen_sweeps():
    """datasets/pipelines/loading.py:102-141 LoadPointCloudFromFile (nuScenes branch) on seeded .bin files: a key
    frame + 4 sweeps with rigid transforms and time lags, points close to the sensor in every sweep."""
    import importlib.util
    import tempfile
    import types

    class _Reg:
        def register_module(self, cls):
            return cls

    if "det3d.datasets" not in sys.modules:
        m = types.ModuleType("det3d.datasets")
        m.__path__ = []
        sys.modules["det3d.datasets"] = m
        r = types.ModuleType("det3d.datasets.registry")
        r.PIPELINES = _Reg()
        sys.modules["det3d.datasets.registry"] = r
        pm = types.ModuleType("det3d.datasets.pipelines")
        pm.__path__ = []
        sys.modules["det3d.datasets.pipelines"] = pm
    spec = importlib.util.spec_from_file_location("det3d.datasets.pipelines.loading",
                                                  "/root/reference/det3d/datasets/pipelines/loading.py")
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "det3d.datasets.pipelines"
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(18)
    tmp = tempfile.mkdtemp(prefix="pn_sweeps_")
    save, files = {}, []
    for k, n in enumerate([3000, 2500, 2800, 100, 2600]):
        p = np.zeros((n, 5), np.float32)
        p[:, :2] = rng.normal(0, 12, (n, 2))
        p[: n // 10, :2] = rng.uniform(-1.5, 1.5, (n // 10, 2))      # many inside / on the 1 m box
        p[:, 2] = rng.uniform(-3, 1, n)
        p[:, 3] = rng.uniform(0, 255, n)
        p[:, 4] = rng.integers(0, 32, n)
        path = os.path.join(tmp, f"s{k}.bin")
        p.tofile(path)
        files.append(path)
        save[f"raw{k}"] = p
    sweeps = []
    for k in range(1, 5):
        a = rng.uniform(-0.2, 0.2)
        T = np.eye(4)
        T[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
        T[:3, 3] = rng.normal(0, 1.5, 3)
        if k == 3:
            T = None                                                   # a sweep without a transform
        sweeps.append(dict(lidar_path=files[k], transform_matrix=T, time_lag=0.05 * k))
        save[f"T{k}"] = np.full((4, 4), np.nan) if T is None else T
        save[f"lag{k}"] = np.float64(0.05 * k)
    loader = mod.LoadPointCloudFromFile(dataset="NuScenesDataset")
    np.random.seed(3)                                                  # the stage draws the sweep order at random
    res = {"lidar": {"nsweeps": 5}, "virtual": False}
    res, _ = loader(res, {"lidar_path": files[0], "sweeps": sweeps})
    np.random.seed(3)
    save["order"] = np.random.choice(4, 4, replace=False)
    save["combined"] = res["lidar"]["combined"].astype(np.float32)
    assert res["lidar"]["combined"].dtype == np.float32
    np.savez_compressed(os.path.join(HERE, "sweeps.npz"), **save)


def gen_neck_head():
    torch.manual_seed(14)
    tasks = [dict(stride=8, class_names=["car"]), dict(stride=8, class_names=["ped", "cone"])]
    neck = RPNV1(layer_nums=[1, 2], num_filters=32, in_channels=[32, 32], logger=logging.getLogger("g"))
    head = CenterHead(tasks=[Config(t) for t in tasks], in_channels=[32], code_weights=[1.0] * 10,
                      common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "vel": (2, 2)},
                      share_channel=16, pillar_size=0.075, point_cloud_range=[-54, -54, -5.0, 54, 54, 3.0],
                      logger=logging.getLogger("g"))
    for m in list(neck.modules()) + list(head.modules()):
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.1)
    neck.eval()
    head.eval()
    x4 = torch.randn(2, 32, 12, 12) * (torch.rand(2, 1, 12, 12) > 0.5)
    x5 = torch.randn(2, 32, 6, 6)
    with torch.no_grad():
        bev = neck({"conv4": x4, "conv5": x5})
        preds = head(bev)
    save = {"x4": x4.numpy(), "x5": x5.numpy(), "bev": bev[0].numpy()}
    for k, v in neck.state_dict().items():
        save["neck." + k] = v.numpy()
    for k, v in head.state_dict().items():
        save["head." + k] = v.numpy()
    for t, p in enumerate(preds):
        for k, v in p.items():
            save[f"pred{t}_{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "neck_head_forward.npz"), **save)


def _randomise_bn(module):
    for m in module.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.1)


def gen_second_stage():
    from det3d.models.second_stage.bev_interpolation import BEVStrideFeature
    from det3d.models.roi_heads.roi_mix_head import RoIMIXHead
    from det3d.models.point_heads.point_head_simple import PointHead
    from det3d.models.detectors.pillar_rcnn import PillarRCNN
    torch.manual_seed(21)
    rng = np.random.default_rng(21)
    ps, pcr = 0.4, [-12.8, -12.8, -2.0, 12.8, 12.8, 4.0]          # 64 x 64 pillars; stride-4 map 16 x 16
    chans = {"conv1": 32, "conv2": 64, "conv3": 128, "conv4": 256}
    strides = {"conv1": 1, "conv2": 2, "conv3": 4, "conv4": 8}
    B, N = 2, 14
    save = {"pillar_size": np.float32(ps), "pc_range": np.asarray(pcr, np.float32)}
    rois = np.zeros((B, N, 7), np.float32)
    for b in range(B):
        r = rand_boxes(rng, N, spread=11.0)
        r[-3:] = 0.0                                                # padded slots of reorder_first_stage_prediction
        r[0, :2] = [12.5, -12.7]                                    # a RoI hanging over the map border (clamped corners)
        rois[b] = r
    roi_scores = rng.uniform(0.05, 0.95, (B, N)).astype(np.float32)
    roi_labels = rng.integers(1, 4, (B, N)).astype(np.int64)
    roi_scores[:, -3:], roi_labels[:, -3:] = 0.0, 0
    save.update(rois=rois, roi_scores=roi_scores, roi_labels=roi_labels)
    conv2 = torch.randn(B, 64, 32, 32) * (torch.rand(B, 1, 32, 32) > 0.6)
    conv3 = torch.randn(B, 128, 16, 16) * (torch.rand(B, 1, 16, 16) > 0.4)
    bev = torch.randn(B, 128, 16, 16)
    save.update(conv2=conv2.numpy(), conv3=conv3.numpy(), bev=bev.numpy())
    mcfg = Config(dict(CLASS_AGNOSTIC=True, SHARED_FC=[32, 32], CLS_FC=[32, 32], REG_FC=[32, 32], DP_RATIO=0.3,
                       TARGET_CONFIG=dict(ROI_PER_IMAGE=128, FG_RATIO=0.5, SAMPLE_ROI_BY_EACH_CLASS=True,
                                          CLS_SCORE_TYPE="roi_iou", CLS_FG_THRESH=0.7, CLS_BG_THRESH=0.25,
                                          CLS_BG_THRESH_LO=0.1, HARD_BG_RATIO=0.8, REG_FG_THRESH=0.5),
                       LOSS_CONFIG=dict(CLS_LOSS="BinaryCrossEntropy", REG_LOSS="L1",
                                        LOSS_WEIGHTS={"rcnn_cls_weight": 1.0, "rcnn_reg_weight": 1.0,
                                                      "code_weights": [1.0] * 7})))
    pcfg = Config(dict(CLASS_AGNOSTIC=True, CLS_FC=[32, 32], TARGET_CONFIG=dict(GT_EXTRA_WIDTH=[0.2, 0.2, 0.2]),
                       LOSS_CONFIG=dict(LOSS_REG="smooth-l1", LOSS_WEIGHTS={"point_cls_weight": 1.0})))
    variants = {
        # the shipped config: out_stride 4, lateral conv3 with stride 1 (ConvTranspose2d k = 1)
        "a": dict(feature_sources=["conv3"], out_stride=4),
        # out_stride 2: top-down and conv3 lateral are ConvTranspose2d(k = 2, stride = 2), conv2 lateral k = 1
        "b": dict(feature_sources=["conv2", "conv3"], out_stride=2),
    }
    for tag, v in variants.items():
        torch.manual_seed(100 + ord(tag))
        mod = BEVStrideFeature(grid_size=7, in_channels=128, share_channels=64, pillar_size=ps, pc_range=pcr,
                               backbone_channels=chans, backbone_strides=strides, **v)
        head = RoIMIXHead(in_channels=64, model_cfg=mcfg, num_class=1, code_size=7, mixer_type="", num_patches=49)
        phead = PointHead(in_channels=64, num_class=1, model_cfg=pcfg)
        for m in (mod, head, phead):
            _randomise_bn(m)
            m.eval()
        torch.nn.init.normal_(head.reg_layers[-1].weight, mean=0, std=0.05)     # visible residuals
        example = {"rois": torch.from_numpy(rois.copy()), "roi_scores": torch.from_numpy(roi_scores.copy()),
                   "roi_labels": torch.from_numpy(roi_labels.copy()), "bev_feature": bev,
                   "backbone_features": {"conv2": conv2, "conv3": conv3}, "batch_size": B,
                   "metadata": [None] * B}
        with torch.no_grad():
            example = mod.forward(example)
            save[f"{tag}.roi_features"] = example["roi_features"].numpy().copy()
            save[f"{tag}.point_coords"] = example["point_coords"].numpy().copy()
            example = phead(example)
            save[f"{tag}.point_cls_scores"] = example["point_cls_scores"].numpy().copy()
            out = head(example, training=False)
            save[f"{tag}.batch_cls_preds"] = out["batch_cls_preds"].numpy().copy()
            save[f"{tag}.batch_box_preds"] = out["batch_box_preds"].numpy().copy()
            dets = PillarRCNN.post_process(None, out)
        for i, d in enumerate(dets):
            for k in ("box3d_lidar", "scores", "label_preds"):
                save[f"{tag}.det{i}.{k}"] = d[k].numpy().copy()
        for name, m in (("second_stage", mod), ("roi_head", head), ("point_head", phead)):
            for k, t in m.state_dict().items():
                save[f"{tag}.{name}.{k}"] = t.numpy()
    np.savez_compressed(os.path.join(HERE, "second_stage.npz"), **save)


def gen_cfg():
    cfg = Config.fromfile("/root/reference/configs/pillarnet/pillarnet34_fpn_centerhead_waymo.py")
    out = set_by_task_cfg(cfg.test_cfg, [1, 2])
    with open(os.path.join(HERE, "set_by_task_cfg.json"), "w") as fh:
        json.dump(json.loads(json.dumps(out)), fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    gen_iou()
    gen_circle()
    gen_predict()
    gen_predict_double_flip()
    gen_loss()
    gen_assign_label()
    gen_sweeps()
    gen_neck_head()
    gen_second_stage()
    gen_cfg()
    print("golden fixtures written to", HERE)
