"""Model / test configurations of the benchmark workloads, restated as plain dicts.

Values follow the reference's config modules (which are not available on the GPU box):
  nusc18  : configs/pillarnet/pillarnet_centerhead_nusc.py:6-50,68-81   (PillarNet-18, nuScenes, 10 sweeps)
  nusc34  : the same with backbone.type = "PillarResNet34" (BASELINE config 4; no such file ships)
  waymo34 : configs/pillarnet/pillarnet34_fpn_centerhead_waymo.py:4-45,64-79 (PillarNet-34 + RPNG FPN)
  pillarrcnn_waymo : configs/pillarrcnn/pillarrcnn_fpn_centerhead_waymo.py:4-112,133-142 (Pillar R-CNN: PillarNet-18 + RPNG
            first stage, BEVStrideFeature + PointHead + RoIMIXHead second stage); pillarrcnn_toy: the same on a 128 x 128 grid
On a machine that has the reference tree, `Config.fromfile(<reference config>)` builds the same models.
"""
import copy


def _nusc(backbone):
    tasks = [
        dict(stride=8, class_names=["car"]),
        dict(stride=8, class_names=["truck", "construction_vehicle"]),
        dict(stride=8, class_names=["bus", "trailer"]),
        dict(stride=8, class_names=["barrier"]),
        dict(stride=8, class_names=["motorcycle", "bicycle"]),
        dict(stride=8, class_names=["pedestrian", "traffic_cone"]),
    ]
    ps, pcr = 0.075, [-54, -54, -5.0, 54, 54, 3.0]
    model = dict(
        type="PillarNet",
        pretrained=None,
        reader=dict(type="DynamicPFE", in_channels=5, num_filters=(32,), pillar_size=ps, pc_range=pcr),
        backbone=dict(type=backbone, in_channels=32),
        neck=dict(type="RPNV1", layer_nums=[5, 5], num_filters=256, in_channels=[256, 256]),
        bbox_head=dict(
            type="CenterHead", tasks=tasks, in_channels=[256],
            code_weights=[1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.2, 0.2, 1.0, 1.0],
            common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "vel": (2, 2)},
            reg_iou="GIoU", pillar_size=ps, point_cloud_range=pcr),
    )
    test_cfg = dict(
        nms=dict(use_rotate_nms=True, nms_pre_max_size=1000, nms_post_max_size=83, nms_iou_threshold=0.2),
        rectifier=0, score_threshold=0.1, double_flip=False,
        post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0],
    )
    # pillarnet_centerhead_nusc.py:61-66
    train_cfg = dict(hm_weight=1, bbox_weight=0.25, iou_weight=1, reg_iou_weight=0.25)
    return dict(model=model, test_cfg=test_cfg, train_cfg=train_cfg, synth="nuscenes", pillar_size=ps, pc_range=pcr)


def _waymo34():
    tasks = [dict(stride=8, class_names=["VEHICLE"]), dict(stride=4, class_names=["PEDESTRIAN", "CYCLIST"])]
    ps, pcr = 0.1, [-75.2, -75.2, -2, 75.2, 75.2, 4]
    model = dict(
        type="PillarNet",
        reader=dict(type="DynamicPFE", in_channels=5, num_filters=(32,), pillar_size=ps, pc_range=pcr),
        backbone=dict(type="PillarResNet34", in_channels=32),
        neck=dict(type="RPNG", layer_nums=[5, 5], num_filters=[256, 128], in_channels=[256, 256, 128]),
        bbox_head=dict(
            type="CenterHead", tasks=tasks, in_channels=[256, 128],
            code_weights=[1.0] * 8,
            common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "iou": (1, 2)},
            reg_iou="GIoU", pillar_size=ps, point_cloud_range=pcr),
    )
    test_cfg = dict(
        nms=dict(use_multi_class_nms=True, nms_pre_max_size=[2048, 1024, 1024],
                 nms_post_max_size=[200, 150, 150], nms_iou_threshold=[0.8, 0.55, 0.55]),
        rectifier=[0., 0., 0.], score_threshold=0.1,
        post_center_limit_range=[-80, -80, -10.0, 80, 80, 10.0],
    )
    # pillarnet34_fpn_centerhead_waymo.py:56-61
    train_cfg = dict(hm_weight=1, bbox_weight=2, iou_weight=1, reg_iou_weight=2)
    return dict(model=model, test_cfg=test_cfg, train_cfg=train_cfg, synth="waymo", pillar_size=ps, pc_range=pcr)


def _pillarrcnn(ps, pcr, synth, post_range):
    tasks = [dict(stride=8, class_names=["VEHICLE"]), dict(stride=4, class_names=["PEDESTRIAN", "CYCLIST"])]
    first = dict(
        type="PillarNet",
        reader=dict(type="DynamicPFE", in_channels=5, num_filters=(32,), pillar_size=ps, pc_range=pcr),
        backbone=dict(type="PillarResNet18", in_channels=32),
        neck=dict(type="RPNG", layer_nums=[5, 5], num_filters=[256, 128], in_channels=[256, 256, 128]),
        bbox_head=dict(type="CenterHead", tasks=tasks, in_channels=[256, 128], code_weights=[1.0] * 8,
                       common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2)},
                       reg_iou="GIoU", pillar_size=ps, point_cloud_range=pcr),
    )
    model = dict(
        type="PillarRCNN", freeze=False, first_stage_cfg=first,
        second_stage_modules=[dict(type="BEVStrideFeature", feature_sources=["conv3"], grid_size=7, out_stride=4,
                                   in_channels=128, share_channels=64, pillar_size=ps, pc_range=pcr)],
        point_head=dict(type="PointHead", in_channels=64, num_class=1,
                        model_cfg=dict(CLASS_AGNOSTIC=True, CLS_FC=[256, 256],
                                       TARGET_CONFIG=dict(GT_EXTRA_WIDTH=[0.2, 0.2, 0.2]),
                                       LOSS_CONFIG=dict(LOSS_REG="smooth-l1", LOSS_WEIGHTS={"point_cls_weight": 1.0}))),
        roi_head=dict(type="RoIMIXHead", in_channels=64, mixer_type="", num_patches=7 * 7, code_size=7,
                      model_cfg=dict(CLASS_AGNOSTIC=True, SHARED_FC=[256, 256], CLS_FC=[256, 256], REG_FC=[256, 256],
                                     DP_RATIO=0.3,
                                     TARGET_CONFIG=dict(ROI_PER_IMAGE=128, FG_RATIO=0.5, SAMPLE_ROI_BY_EACH_CLASS=True,
                                                        CLS_SCORE_TYPE="roi_iou", CLS_FG_THRESH=0.7, CLS_BG_THRESH=0.25,
                                                        CLS_BG_THRESH_LO=0.1, HARD_BG_RATIO=0.8, REG_FG_THRESH=0.5),
                                     LOSS_CONFIG=dict(CLS_LOSS="BinaryCrossEntropy", REG_LOSS="L1",
                                                      LOSS_WEIGHTS={"rcnn_cls_weight": 1.0, "rcnn_reg_weight": 1.0,
                                                                    "code_weights": [1.0] * 7}))),
    )
    test_cfg = dict(
        nms=dict(use_multi_class_nms=True, nms_pre_max_size=[2048, 1024, 1024],
                 nms_post_max_size=[200, 150, 150], nms_iou_threshold=[0.8, 0.55, 0.55]),
        rectifier=[0., 0., 0.], score_threshold=0.1, post_center_limit_range=post_range,
    )
    train_cfg = dict(hm_weight=1, bbox_weight=2, iou_weight=1, reg_iou_weight=2)
    return dict(model=model, test_cfg=test_cfg, train_cfg=train_cfg, synth=synth, pillar_size=ps, pc_range=pcr)


WORKLOADS = {
    "pillarrcnn_waymo": lambda: _pillarrcnn(0.1, [-75.2, -75.2, -2, 75.2, 75.2, 4], "waymo", [-80, -80, -10.0, 80, 80, 10.0]),
    "pillarrcnn_toy": lambda: _pillarrcnn(0.3, [-19.2, -19.2, -2, 19.2, 19.2, 4], "waymo", [-20, -20, -10.0, 20, 20, 10.0]),
    "nusc18": lambda: _nusc("PillarResNet18"),
    "nusc34": lambda: _nusc("PillarResNet34"),
    "waymo34": _waymo34,
}


def get(name):
    return copy.deepcopy(WORKLOADS[name]())
