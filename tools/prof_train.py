"""CUPTI breakdown of one training step as the TrainEngine replays it (two CUDA graphs): per-kernel device time."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pillarnet_lts_b200 as P  # noqa: E402
from pillarnet_lts_b200 import configs, synth, train  # noqa: E402
from pillarnet_lts_b200.registry import ConfigDict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="nusc34")
ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--out", default="gpurun_out/prof_train.json")
args = ap.parse_args()
dev = torch.device("cuda")
cfg = configs.get(args.workload)
torch.manual_seed(0)
model = P.build_detector(ConfigDict.wrap(cfg["model"]), cfg["train_cfg"], ConfigDict.wrap(cfg["test_cfg"])).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, capturable=True, fused=True)
rng = np.random.default_rng(0)
B = args.frames
fs = synth.make_batch(cfg["synth"], B, 100)
offs = np.cumsum([0] + [len(f) for f in fs]).astype(np.int32)
ex = {"points_batched": (torch.from_numpy(np.concatenate(fs)).to(dev), torch.from_numpy(offs).to(dev)), "points": None,
      "metadata": [None] * B}
ex.update(train.synthetic_targets(model.bbox_head, B, model.reader.height, model.reader.width, rng, device=dev))
eng = train.TrainEngine(model, opt, B, ex["points_batched"][0].shape[0] + 1024, ex).prepare(warmup=3)
for _ in range(3):
    eng.step(ex)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(args.steps):
        eng.step(ex)
    torch.cuda.synchronize()
agg = {}
for e in prof.events():
    if e.device_type.name != "CUDA":
        continue
    a = agg.setdefault(e.name, [0.0, 0])
    a[0] += e.device_time
    a[1] += 1
rows = sorted(((k, v[0] / args.steps, v[1] / args.steps) for k, v in agg.items()), key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"kernel time per step {tot / 1e3:.2f} ms over {sum(r[2] for r in rows):.0f} launches")
for name, us, n in rows[:45]:
    print(f"{us:10.1f} us {100 * us / tot:5.1f}%  n={n:6.1f}  {name[:120]}")
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump([dict(kernel=k[:160], us_per_step=u, launches=n) for k, u, n in rows], open(args.out, "w"), indent=0)
