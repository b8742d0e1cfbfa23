"""PillarNet single-stage detector behind det3d's detector interface.

Mirrors det3d/models/detectors/pillarnet.py:6-49 and single_stage.py:10-59: builds reader / backbone /
neck / bbox_head from their cfg dicts through the registries, regroups test_cfg per task
(core/utils/center_utils.py:229-274) and runs reader -> backbone -> neck -> head -> predict.
`forward_device` is the sync-free variant used under CUDA graphs / by bench.py.
"""
import torch
from torch import nn

from .registry import DETECTORS, ConfigDict, build_backbone, build_head, build_neck, build_reader


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

    @torch.no_grad()
    def forward_device(self, points, frame_offsets):
        """Sync-free inference: (det_out, keep_count, plan) — see CenterHead.predict_raw."""
        bev_features, _ = self.extract_feat(dict(points_batched=(points, frame_offsets)))
        preds = self.bbox_head(bev_features)
        return self.bbox_head.predict_raw(preds, self.test_cfg)
