"""per-CTA pipeline timelines of every tensor-core conv launch of one eager inference pass"""
import os, sys
os.environ["PN_DENSE_TIMELINE"] = "1"
os.environ["PN_CONV_TIMELINE"] = "1"
os.environ["PN_PDL"] = "0"
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import pillarnet_lts_b200 as P
from pillarnet_lts_b200.engine import calibrate_heatmap_bias
dev = torch.device("cuda")
P.set_precision("bf16")
model, cfg = bench.build_model("nusc18", dev)
frames = [bench.make_frames(cfg["synth"], 1, seed0=1000)[0]]
calibrate_heatmap_bias(model, frames, target_cells=1500)
pts = torch.from_numpy(frames[0]).to(dev)
off = torch.tensor([0, len(frames[0])], dtype=torch.int32, device=dev)
with torch.no_grad():
    for i in range(2):
        if i == 1:
            print("==== pass 2 (warm) ====", file=sys.stderr)
        model.forward_device(pts, off)
        torch.cuda.synchronize()
