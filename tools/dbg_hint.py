import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
import pillarnet_lts_b200 as P
from pillarnet_lts_b200 import backbone
from pillarnet_lts_b200.engine import InferenceEngine, calibrate_heatmap_bias
dev = torch.device("cuda")
P.set_precision("bf16")
model, cfg = bench.build_model("nusc18", dev)
frames = bench.make_frames(cfg["synth"], 1, seed0=1000)
calibrate_heatmap_bias(model, frames, target_cells=1500)
eng = InferenceEngine(model, 1, 300000, device=dev)
eng.upload(eng.stage_host(frames))
eng.prepare(warmup=2)
print("observed", backbone._observed_rows, getattr(eng, "rows_seen", None))
