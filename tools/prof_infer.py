"""Warm per-kernel durations of one inference step (CUDA-graph replay) from torch.profiler / CUPTI activity records.
Unlike ncu (which flushes caches and serialises), this is the step exactly as bench.py times it."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pillarnet_lts_b200 as P  # noqa: E402
from pillarnet_lts_b200.engine import InferenceEngine, calibrate_heatmap_bias  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="nusc18")
ap.add_argument("--frames", type=int, default=1)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--out", default="gpurun_out/prof_infer.json")
args = ap.parse_args()
dev = torch.device("cuda")
P.set_precision("bf16")
model, cfg = bench.build_model(args.workload, dev)
B = args.frames
frames = [bench.make_frames(cfg["synth"], 1, seed0=1000 + j)[0] for j in range(B)]
calibrate_heatmap_bias(model, frames, target_cells=1500)
cap = int(sum(len(f) for f in frames) * 1.05) + 1024
eng = InferenceEngine(model, B, cap, device=dev)
eng.upload(eng.stage_host(frames))
eng.prepare(warmup=2)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5):
    eng.launch()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(args.steps):
        with torch.cuda.stream(eng.stream):
            flush.zero_()
        eng.launch()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type.name == "CUDA" and "Memset" not in e.name and "fill" not in e.name.lower()]
# order of first appearance within a step
agg = {}
for e in ev:
    a = agg.setdefault(e.name, [0.0, 0])
    a[0] += e.device_time
    a[1] += 1
rows = sorted(((k, v[0] / args.steps, v[1] / args.steps) for k, v in agg.items()), key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
# timeline gaps: sort events by start, sum idle time between consecutive kernels inside a step
evs = sorted(ev, key=lambda e: e.time_range.start)
busy = sum(e.device_time for e in evs)
span = (evs[-1].time_range.end - evs[0].time_range.start)
print(f"kernel time per step {tot:.1f} us over {sum(r[2] for r in rows):.0f} launches; profiled span per step "
      f"{span / args.steps:.1f} us (includes the L2-flush fill)")
out = []
for name, us, n in rows:
    short = name.replace("(anonymous namespace)::", "").replace("void ", "")[:90]
    print(f"{us:9.1f} us {100 * us / tot:5.1f}%  n={n:5.1f} avg={us / n:7.1f}  {short}")
    out.append(dict(kernel=short, us_per_step=us, launches=n))
json.dump(out, open(os.path.join(ROOT, args.out), "w"), indent=1)

# --- timeline of the last profiled step: start offset, duration and stream of every kernel (what overlaps what) ---
if os.environ.get("PN_PROF_TIMELINE", "0") == "1":
    per = len(evs) // args.steps
    last = evs[-per:]
    t0 = last[0].time_range.start
    end_prev = {}
    print(f"--- timeline of one replay ({per} records) ---")
    crit = t0
    for e in last:
        st, en = e.time_range.start - t0, e.time_range.end - t0
        sid = getattr(e, "device_resource_id", None)
        short = e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:48]
        print(f"{st:8.1f} -> {en:8.1f}  ({en - st:6.1f} us) stream {sid}  {short}")
