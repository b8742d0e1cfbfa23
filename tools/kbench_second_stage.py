"""Pillar R-CNN second stage at Waymo size (stride-4 map 376 x 376, 500 RoIs per frame): CUDA-event timing of the
whole second stage (BEV fusion -> RoI grid pooling -> point head -> RoI head -> refinement) and of its parts, eager and
as one CUDA graph; random-init weights, synthetic first-stage outputs."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pillarnet_lts_b200 as P  # noqa: E402
from pillarnet_lts_b200 import configs, ops  # noqa: E402
from pillarnet_lts_b200.registry import build_point_head, build_roi_head, build_second_stage_module  # noqa: E402

dev = torch.device("cuda")
P.set_precision("bf16")
cfg = configs.get("pillarrcnn_waymo")["model"]
chans = {"conv1": 32, "conv2": 64, "conv3": 128, "conv4": 256}
strides = {"conv1": 1, "conv2": 2, "conv3": 4, "conv4": 8}
torch.manual_seed(0)
ss_cfg = dict(cfg["second_stage_modules"][0], backbone_channels=chans, backbone_strides=strides)
mod = build_second_stage_module(ss_cfg).to(dev).eval()
phead = build_point_head(cfg["point_head"]).to(dev).eval()
head = build_roi_head(cfg["roi_head"]).to(dev).eval()
B, N, H, W = 1, 500, 376, 376
g = torch.Generator(device="cuda").manual_seed(1)
bev = torch.randn(B, 128, H, W, device=dev, generator=g).to(torch.bfloat16)
conv3 = (torch.randn(B, 128, H, W, device=dev, generator=g) * (torch.rand(B, 1, H, W, device=dev, generator=g) > 0.8)).to(torch.bfloat16)
rois = torch.zeros(B, N, 7, device=dev)
rois[..., :2] = torch.rand(B, N, 2, device=dev, generator=g) * 140 - 70
rois[..., 2] = 0.5
rois[..., 3:6] = torch.tensor([4.5, 2.0, 1.6], device=dev)
rois[..., 6] = torch.rand(B, N, device=dev, generator=g) * 6.28 - 3.14
scores = torch.rand(B, N, device=dev, generator=g)
labels = torch.randint(1, 4, (B, N), device=dev, generator=g)


def run():
    ex = {"rois": rois, "roi_scores": scores, "roi_labels": labels, "bev_feature": bev,
          "backbone_features": {"conv3": conv3}, "batch_size": B, "metadata": [None] * B}
    ex = mod(ex)
    ex = phead(ex)
    return head(ex, training=False)


flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=20):
    ts = []
    for _ in range(reps + 3):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts = sorted(ts[3:])
    return ts[len(ts) // 2]


res = {}
with torch.no_grad():
    for _ in range(3):
        out = run()
    torch.cuda.synchronize()
    res["eager_us"] = timeit(run)
    graph = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        run()
        st.synchronize()
        with torch.cuda.graph(graph, stream=st):
            out = run()
    torch.cuda.synchronize()
    res["graph_us"] = timeit(graph.replay)
    fused = mod.fused_map(bev, {"conv3": conv3})
    res["bev_fusion_us"] = timeit(lambda: mod.fused_map(bev, {"conv3": conv3}))
    res["roi_grid_bilinear_us"] = timeit(lambda: ops.roi_grid_bilinear(rois, 7, fused.rows, B, fused.H, fused.W, fused.C,
                                                                       -75.2, -75.2, 0.4, feat_coff=fused.coff,
                                                                       padded=bool(fused.pad)))
    ex = {"rois": rois, "roi_scores": scores, "roi_labels": labels, "bev_feature": bev,
          "backbone_features": {"conv3": conv3}, "batch_size": B}
    ex = mod(ex)
    res["point_head_us"] = timeit(lambda: phead(dict(ex)))
    res["roi_head_us"] = timeit(lambda: head(dict(ex), training=False))
# algorithmic work: two 1x1 GEMMs 128 -> 128 and the 3x3 fusion conv 256 -> 64 over H*W pixels; FC stacks per RoI
flop = 2.0 * B * H * W * (2 * 128 * 128 + 9 * 256 * 64) + 2.0 * B * N * (3136 * 256 + 256 * 256 + 2 * (2 * 256 * 256) + 256 * 8) \
       + 2.0 * B * N * 49 * (64 * 256 + 256 * 256 + 256)
res["flop"] = flop
res["graph_tflops"] = flop / res["graph_us"] / 1e6
res["config"] = dict(B=B, rois=N, map=[H, W], grid=7, precision="bf16")
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "kbench_second_stage.json"), "w"), indent=1)
