"""Which aten / autograd ops own the small-kernel time of a training step: the eager TrainEngine step under the PyTorch
profiler (CPU + CUDA), grouped by op name and input shapes."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pillarnet_lts_b200 as P  # noqa: E402
from pillarnet_lts_b200 import configs, synth, train  # noqa: E402
from pillarnet_lts_b200.registry import ConfigDict  # noqa: E402

dev = torch.device("cuda")
cfg = configs.get("nusc34")
torch.manual_seed(0)
model = P.build_detector(ConfigDict.wrap(cfg["model"]), cfg["train_cfg"], ConfigDict.wrap(cfg["test_cfg"])).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, capturable=True, fused=True)
rng = np.random.default_rng(0)
B = 4
fs = synth.make_batch(cfg["synth"], B, 100)
offs = np.cumsum([0] + [len(f) for f in fs]).astype(np.int32)
ex = {"points_batched": (torch.from_numpy(np.concatenate(fs)).to(dev), torch.from_numpy(offs).to(dev)), "points": None,
      "metadata": [None] * B}
ex.update(train.synthetic_targets(model.bbox_head, B, model.reader.height, model.reader.width, rng, device=dev))
eng = train.TrainEngine(model, opt, B, ex["points_batched"][0].shape[0] + 1024, ex, use_graph=False).prepare(warmup=2)
for _ in range(2):
    eng.step(ex)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    eng.step(ex)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=48,
                                                         max_shapes_column_width=60))
