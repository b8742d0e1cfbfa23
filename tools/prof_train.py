"""torch.profiler breakdown of one training step (where the time goes: our kernels vs cuDNN vs elementwise vs host)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pillarnet_lts_b200 as P  # noqa: E402
from pillarnet_lts_b200 import configs, synth, train  # noqa: E402
from pillarnet_lts_b200.registry import ConfigDict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="nusc18")
ap.add_argument("--frames", type=int, default=2)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
dev = torch.device("cuda")
cfg = configs.get(args.workload)
torch.manual_seed(0)
model = P.build_detector(ConfigDict.wrap(cfg["model"]), cfg["train_cfg"], ConfigDict.wrap(cfg["test_cfg"])).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
rng = np.random.default_rng(0)
B = args.frames
fs = synth.make_batch(cfg["synth"], B, 100)
offs = np.cumsum([0] + [len(f) for f in fs]).astype(np.int32)
ex = {"points_batched": (torch.from_numpy(np.concatenate(fs)).to(dev), torch.from_numpy(offs).to(dev)), "points": None,
      "metadata": [None] * B}
ex.update(train.synthetic_targets(model.bbox_head, B, model.reader.height, model.reader.width, rng, device=dev))
for _ in range(3):
    train.train_step(model, ex, opt)
torch.cuda.synchronize()


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return r, (time.perf_counter() - t0) * 1e3


stages = {}
opt.zero_grad(set_to_none=True)
sp, stages["reader_fwd"] = timed(lambda: model.reader(dict(points_batched=ex["points_batched"])))
feats, stages["backbone_fwd"] = timed(lambda: model.backbone(sp))
bev, stages["neck_fwd"] = timed(lambda: model.neck(feats))
preds, stages["head_fwd"] = timed(lambda: model.bbox_head(bev))
losses, stages["loss"] = timed(lambda: model.bbox_head.loss(ex, preds, model.train_cfg))
loss = sum(l.sum() for l in losses["loss"])
_, stages["backward"] = timed(lambda: loss.backward())
_, stages["optimizer"] = timed(lambda: opt.step())
print("stage wall ms (sync on both sides):", json.dumps({k: round(v, 2) for k, v in stages.items()}))

from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(args.steps):
        train.train_step(model, ex, opt)
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = sorted([(k.key, getattr(k, "device_time_total", 0.0) / args.steps, k.count / args.steps) for k in ka
               if getattr(k, "device_time_total", 0.0) > 0 and k.device_type.name == "CUDA"], key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"CUDA kernel time per step: {tot / 1e3:.2f} ms over {sum(r[2] for r in rows):.0f} launches")
for name, us, n in rows[:40]:
    print(f"{us:10.1f} us {100 * us / tot:5.1f}%  n={n:6.1f}  {name[:110]}")
