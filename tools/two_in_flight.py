"""Experiment: two InferenceEngines (own graphs, buffers, streams) on one GPU, alternate frames in flight at once —
does the tail of one frame (top-k, NMS: a handful of CTAs) hide under the other frame's convs?"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pillarnet_lts_b200 as P  # noqa: E402
from pillarnet_lts_b200.engine import InferenceEngine, calibrate_heatmap_bias  # noqa: E402

dev = torch.device("cuda:0")
P.set_precision("bf16")
model, cfg = bench.build_model("nusc18", dev)
frames = bench.make_frames(cfg["synth"], 8, 0)
calibrate_heatmap_bias(model, frames[:1], target_cells=1500)
cap = max(len(f) for f in frames) + 4096
dev_frames = [(torch.from_numpy(f).to(dev), torch.tensor([0, len(f)], dtype=torch.int32, device=dev)) for f in frames]


def make():
    e = InferenceEngine(model, 1, cap)
    e.stage_host([torch.from_numpy(frames[0])])
    e.upload(len(frames[0]))
    return e.prepare(warmup=2)


def run(engs, steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        e = engs[i % len(engs)]
        p, o = dev_frames[i % 8]
        with torch.cuda.stream(e.stream):
            e.points[:p.shape[0]].copy_(p, non_blocking=True)
            e.offsets.copy_(o, non_blocking=True)
        e.launch()
    torch.cuda.synchronize()
    return steps / (time.perf_counter() - t0)


for n in (1, 2, 3):
    engs = [make() for _ in range(n)]
    run(engs, 50)
    r = [run(engs, 400) for _ in range(3)]
    print(f"{n} engine(s) in flight: {max(r):.0f} frames/s (runs {[round(x) for x in r]})", flush=True)
    del engs
