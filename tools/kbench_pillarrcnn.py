"""Whole Pillar R-CNN (configs/pillarrcnn/pillarrcnn_fpn_centerhead_waymo.py restated: PillarNet-18 + RPNG first stage,
BEVStrideFeature + PointHead + RoIMIXHead second stage) on one synthetic Waymo-shaped frame: eager inference through
the detector's public forward (the RoI reordering reads the detection counts on the host, as the reference does), CUDA
events; random-init weights with the heat-map bias calibrated so the first stage emits a few hundred RoIs."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pillarnet_lts_b200 as P  # noqa: E402
from pillarnet_lts_b200 import configs, synth  # noqa: E402
from pillarnet_lts_b200.engine import calibrate_heatmap_bias  # noqa: E402

dev = torch.device("cuda")
P.set_precision("bf16")
cfg = configs.get("pillarrcnn_waymo")
torch.manual_seed(0)
model = P.build_detector(cfg["model"], train_cfg=None, test_cfg=cfg["test_cfg"]).to(dev).eval()
g = torch.Generator().manual_seed(1)
for m in model.modules():
    if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
        m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
        m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
frames = synth.make_batch(cfg["synth"], 1, 0)
calibrate_heatmap_bias(model.single_det, frames, target_cells=600)
pts = [torch.from_numpy(f).to(dev) for f in frames]


def run():
    with torch.no_grad():
        return model(dict(points=pts, metadata=[{}]), return_loss=False)


def first_stage():
    with torch.no_grad():
        return model.single_det(dict(points=pts, metadata=[{}]), return_loss=False)


def timeit(fn, reps=10):
    ts = []
    for _ in range(reps + 3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    ts = sorted(ts[3:])
    return ts[len(ts) // 2], out


# the sync-free path as one CUDA graph
offs = torch.tensor([0, pts[0].shape[0]], dtype=torch.int32, device=dev)
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for _ in range(2):
        model.forward_device(pts[0], offs)
    st.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=st):
        gout = model.forward_device(pts[0], offs)
torch.cuda.synchronize()
ms_graph, _ = timeit(lambda: graph.replay(), reps=30)
ms_all, dets = timeit(run)
ms_first, d1 = timeit(first_stage)
res = {"points": int(pts[0].shape[0]), "rois_first_stage": int(d1[0]["scores"].shape[0]),
       "detections_second_stage": int(dets[0]["scores"].shape[0]), "pillar_rcnn_ms": ms_all, "first_stage_only_ms": ms_first,
       "second_stage_ms": ms_all - ms_first,
       "pillar_rcnn_graph_ms": ms_graph, "graph_frames_per_s": 1e3 / ms_graph,
       "graph_valid_detections": int(gout[3].sum()), "mode": "eager (host-driven RoI reordering, as the reference), bf16"}
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "kbench_pillarrcnn.json"), "w"), indent=1)
