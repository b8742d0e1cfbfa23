"""PillarNet single-stage detector behind det3d's detector interface.

Mirrors det3d/models/detectors/pillarnet.py:6-49 and single_stage.py:10-59: builds reader / backbone /
neck / bbox_head from their cfg dicts through the registries, regroups test_cfg per task
(core/utils/center_utils.py:229-274) and runs reader -> backbone -> neck -> head -> predict.
`forward_device` is the sync-free variant used under CUDA graphs / by bench.py.
"""
import torch
from torch import nn

from .registry import (DETECTORS, ConfigDict, build_backbone, build_detector, build_head, build_neck,
                       build_reader)


def set_by_task_cfg(test_cfg, task_num_classes):
    """core/utils/center_utils.py:229-274: scalars are kept, per-class lists are regrouped per task."""

    def _param_org(param):
        if isinstance(param, (float, int)):
            return param
        assert isinstance(param, (list, tuple))
        assert len(param) == sum(task_num_classes)
        ret, flag = [], 0
        for num in task_num_classes:
            ret.append(list(param[flag:flag + num]))
            flag += num
        return ret

    if test_cfg.get("rectifier", False):
        test_cfg["rectifier"] = _param_org(test_cfg["rectifier"])
    if test_cfg.get("use_rectify", False):
        test_cfg["use_rectify"] = _param_org(test_cfg["use_rectify"])
    for k in ("nms_pre_max_size", "nms_post_max_size", "nms_iou_threshold"):
        test_cfg["nms"][k] = _param_org(test_cfg["nms"][k])
    return test_cfg


@DETECTORS.register_module
class SingleStageDetector(nn.Module):
    def __init__(self, reader, backbone, neck=None, bbox_head=None, train_cfg=None, test_cfg=None,
                 pretrained=None):
        super().__init__()
        self.reader = build_reader(reader)
        self.backbone = build_backbone(backbone)
        if neck is not None:
            self.neck = build_neck(neck)
        self.bbox_head = build_head(bbox_head)
        self.train_cfg = train_cfg
        self.test_cfg = ConfigDict.wrap(test_cfg) if test_cfg is not None else None
        self.init_weights(pretrained)

    def init_weights(self, pretrained=None):
        """detectors/single_stage.py:27-35: goes through load_checkpoint (module. prefix stripped, spconv-1.x sparse
        weights re-laid out, `meta` objects tolerated, mismatches reported) — not a bare torch.load"""
        if pretrained is None:
            return
        from .checkpoint import load_checkpoint
        load_checkpoint(self, pretrained, map_location="cpu", strict=False)

    @property
    def with_neck(self):
        return hasattr(self, "neck") and self.neck is not None

    def extract_feat(self, data):
        x = self.backbone(self.reader(data))
        if self.with_neck:
            x = self.neck(x)
        return x


@DETECTORS.register_module
class PillarNet(SingleStageDetector):
    def __init__(self, reader, backbone, neck, bbox_head, train_cfg=None, test_cfg=None, pretrained=None):
        super().__init__(reader, backbone, neck, bbox_head, train_cfg, test_cfg, pretrained)
        post = self.test_cfg.nms.nms_post_max_size
        self.NMS_POST_MAXSIZE = sum(post) if isinstance(post, list) else post
        self.num_classes = self.bbox_head.num_classes
        self.test_cfg = set_by_task_cfg(self.test_cfg, self.bbox_head.num_classes)

    def extract_feat(self, data):
        sp_tensor = self.reader(data)
        pillar_features = self.backbone(sp_tensor)
        bev_features = self.neck(pillar_features) if self.with_neck else pillar_features
        return bev_features, pillar_features

    def forward(self, example, return_loss=True, **kwargs):
        batch_size = len(example["metadata"]) if "metadata" in example else len(example["points"])
        data = dict(points=example["points"], batch_size=batch_size)
        if "points_batched" in example:
            data["points_batched"] = example["points_batched"]
        bev_features, _ = self.extract_feat(data)
        preds = self.bbox_head(bev_features)
        if return_loss:
            return self.bbox_head.loss(example, preds, self.train_cfg)
        return self.bbox_head.predict(example, preds, self.test_cfg)

    def forward_two_stage(self, example, return_loss=True, **kwargs):
        """detectors/pillarnet.py:51-82: first-stage detections plus the maps the second stage pools from"""
        batch_size = len(example["metadata"]) if "metadata" in example else len(example["points"])
        data = dict(points=example["points"], batch_size=batch_size)
        if "points_batched" in example:
            data["points_batched"] = example["points_batched"]
        bev_features, backbone_features = self.extract_feat(data)
        preds = self.bbox_head(bev_features)
        if return_loss:
            raise NotImplementedError("two-stage training is outside the inference path")
        boxes = self.bbox_head.predict(example, preds, self.test_cfg)
        return boxes, bev_features, backbone_features, None

    @torch.no_grad()
    def forward_device(self, points, frame_offsets):
        """Sync-free inference: (det_out, keep_count, plan) — see CenterHead.predict_raw."""
        bev_features, _ = self.extract_feat(dict(points_batched=(points, frame_offsets)))
        preds = self.bbox_head(bev_features)
        return self.bbox_head.predict_raw(preds, self.test_cfg)


@DETECTORS.register_module
class PillarRCNN(nn.Module):
    """detectors/pillar_rcnn.py:9-170, inference path: first stage -> RoIs (padded to NMS_POST_MAXSIZE) -> second-stage
    modules -> point head -> RoI head -> post-processing.  Same constructor kwargs and module tree (`single_det`,
    `second_stage`, `point_head`, `roi_head`) as the reference, so its checkpoints load."""

    def __init__(self, first_stage_cfg, second_stage_modules, roi_head, freeze=False, point_head=None,
                 train_cfg=None, test_cfg=None, pretrained=None, **kwargs):
        super().__init__()
        from .registry import build_point_head, build_roi_head, build_second_stage_module
        self.single_det = build_detector(first_stage_cfg, train_cfg=train_cfg, test_cfg=test_cfg)
        if freeze:
            for p in self.single_det.parameters():
                p.requires_grad = False
            self.single_det.eval()
        self.bbox_head = self.single_det.bbox_head
        self.test_cfg = self.single_det.test_cfg
        self.num_classes = sum(self.single_det.num_classes)
        first = dict(backbone_channels=self.single_det.backbone.backbone_channels,
                     backbone_strides=self.single_det.backbone.backbone_strides)
        self.second_stage = nn.ModuleList()
        for m in second_stage_modules:
            m = dict(m)
            m.update(first)
            self.second_stage.append(build_second_stage_module(m))
        self.point_head = build_point_head(point_head) if point_head is not None else None
        self.roi_head = build_roi_head(roi_head)
        if pretrained is not None:
            from .checkpoint import load_checkpoint
            load_checkpoint(self, pretrained, map_location="cpu", strict=False)

    def reorder_first_stage_prediction(self, first_pred, example):
        """pillar_rcnn.py:50-83: per-frame detections -> fixed-size (B, NMS_POST_MAXSIZE, .) RoI tensors, labels + 1"""
        B = len(first_pred)
        D = first_pred[0]["box3d_lidar"].shape[1]
        n_max = self.single_det.NMS_POST_MAXSIZE
        dev = first_pred[0]["box3d_lidar"].device
        rois = torch.zeros(B, n_max, D, dtype=torch.float32, device=dev)
        roi_scores = torch.zeros(B, n_max, dtype=torch.float32, device=dev)
        roi_labels = torch.zeros(B, n_max, dtype=torch.long, device=dev)
        for i in range(B):
            boxes = first_pred[i]["box3d_lidar"]
            n = boxes.shape[0]
            if self.roi_head.code_size == 9:
                boxes = boxes[:, [0, 1, 2, 3, 4, 5, 8, 6, 7]]
            rois[i, :n] = boxes
            roi_labels[i, :n] = first_pred[i]["label_preds"] + 1
            roi_scores[i, :n] = first_pred[i]["scores"]
        example["rois"], example["roi_labels"], example["roi_scores"] = rois, roi_labels, roi_scores
        example["has_class_labels"] = True
        return example

    def second_stage_forward(self, example):
        """everything after the first stage: needs rois / roi_scores / roi_labels, bev_feature, backbone_features"""
        for module in self.second_stage:
            example = module(example)
        if self.point_head is not None:
            example = self.point_head(example)
        return self.roi_head(example, training=False)

    def forward(self, example, return_loss=True, **kwargs):
        if return_loss:
            raise NotImplementedError("PillarRCNN: only the inference path is implemented")
        batch_size = len(example["metadata"]) if "metadata" in example else len(example["points"])
        example["batch_size"] = batch_size
        preds, bev_features, backbone_features, _ = self.single_det.forward_two_stage(example, False, **kwargs)
        example["bev_feature"] = bev_features[-1]
        example["backbone_features"] = backbone_features
        example = self.reorder_first_stage_prediction(preds, example)
        return self.post_process(self.second_stage_forward(example))

    @torch.no_grad()
    def forward_device(self, points, frame_offsets):
        """Sync-free inference (CUDA-graph capturable): every NMS slot of the first stage is a RoI — slot k of segment
        s of frame b, valid iff k < keep_count[b, s] — instead of the reference's compacted-then-padded list, so nothing
        is read back on the host; invalid slots carry label 0 and come out with valid = False, exactly as the padded
        slots of `reorder_first_stage_prediction` do.  Returns (boxes (B, N, code), scores (B, N), labels (B, N) int64
        zero-based, valid (B, N) bool) with N = segments * post_cap; the valid rows are the detections `forward`
        returns (same values, segment-major order)."""
        det = self.single_det
        bev_features, backbone_features = det.extract_feat(dict(points_batched=(points, frame_offsets)))
        preds = det.bbox_head(bev_features)
        det_out, keep_count, plan = det.bbox_head.predict_raw(preds, det.test_cfg)
        B, S, post_cap = plan["B"], plan["S"], plan["post_cap"]
        head = det.bbox_head
        d = det_out.view(B, S, post_cap, 11)
        live = torch.arange(post_cap, device=d.device).view(1, 1, post_cap) < keep_count.view(B, S, 1)
        cls_off, flag = [], 0
        for n in head.num_classes:
            cls_off.append(flag)
            flag += n
        from .losses import device_const           # cached device constant: no host-to-device copy inside a capture
        off = device_const(tuple(cls_off[seg["task"]] for seg in plan["segs"]), d.device, torch.int64).view(1, S, 1)
        if self.roi_head.code_size == 9:
            boxes = d[..., [0, 1, 2, 3, 4, 5, 8, 6, 7]]
        else:
            boxes = torch.cat([d[..., :6], d[..., 8:9]], -1)
        zero = torch.zeros((), dtype=d.dtype, device=d.device)
        example = {
            "batch_size": B,
            "rois": torch.where(live.unsqueeze(-1), boxes, zero).reshape(B, S * post_cap, -1).contiguous(),
            "roi_scores": torch.where(live, d[..., 9], zero).reshape(B, S * post_cap),
            "roi_labels": torch.where(live, d[..., 10].to(torch.int64) + off + 1,
                                      torch.zeros((), dtype=torch.int64, device=d.device)).reshape(B, S * post_cap),
            "bev_feature": bev_features[-1], "backbone_features": backbone_features, "has_class_labels": True,
        }
        out = self.second_stage_forward(example)
        return out["batch_box_preds"], out["refined_scores"], example["roi_labels"] - 1, out["refined_valid"]

    def post_process(self, batch_dict):
        """pillar_rcnn.py:141-170 (score fusion + validity mask computed by pn_roi_refine)"""
        out = []
        for i in range(batch_dict["batch_size"]):
            boxes = batch_dict["batch_box_preds"][i]
            if boxes.shape[-1] == 9:
                boxes = boxes[:, [0, 1, 2, 3, 4, 5, 7, 8, 6]]
            mask = batch_dict["refined_valid"][i]
            out.append({"box3d_lidar": boxes[mask], "scores": batch_dict["refined_scores"][i][mask],
                        "label_preds": batch_dict["roi_labels"][i][mask] - 1,
                        "metadata": batch_dict["metadata"][i] if "metadata" in batch_dict else None})
        return out
