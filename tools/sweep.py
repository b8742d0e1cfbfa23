"""BASELINE config 5: synthetic point-cloud sweep (points/frame x frames/step) on one GPU.

For every grid point: whole-path throughput (CUDA-graph replays of pillarize -> PFN -> sparse backbone -> neck/head ->
decode -> NMS, L2 flushed between steps) and the reader kernels' achieved bandwidth against the measured HBM peak
(algorithmic bytes of DESIGN.md §3: pillarize 20N+4L+12M, bf16 PFN+scatter-max 24L+64M)."""
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
from pillarnet_lts_b200 import ops, synth  # noqa: E402
from pillarnet_lts_b200.engine import InferenceEngine, calibrate_heatmap_bias  # noqa: E402
from tools.kbench_reader import graph_time  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--points", default="50000,260000,1000000,2000000")
ap.add_argument("--frames", default="1,8,32")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--out", default="gpurun_out/sweep.json")
args = ap.parse_args()
dev = torch.device("cuda")
P.set_precision("bf16")
peaks = bench._peaks()
model, cfg = bench.build_model("nusc18", dev)
calibrate_heatmap_bias(model, [synth.make_frame("nuscenes", 999)], target_cells=1500)
pcr, ps, H = cfg["pc_range"], cfg["pillar_size"], model.reader.height
w, sc, sh = torch.randn(32, 7), torch.ones(32), torch.zeros(32)
rows = []
for n_pts in [int(v) for v in args.points.split(",")]:
    base = [synth.make_frame("nuscenes", 50 + i, n_pts) for i in range(2)]
    for B in [int(v) for v in args.frames.split(",")]:
        if n_pts * B > 70_000_000:
            continue
        frames = [base[i % 2] for i in range(B)]
        N = sum(len(f) for f in frames)
        r = dict(points_per_frame=n_pts, frames=B, N=N)
        try:
            eng = InferenceEngine(model, B, N + 1024, device=dev)
            eng.upload(eng.stage_host(frames))
            eng.prepare(warmup=1)
            flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
            ts = []
            for _ in range(args.steps):
                with torch.cuda.stream(eng.stream):
                    flush.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(eng.stream)
                eng.launch()
                with torch.cuda.stream(eng.stream):
                    b.record(eng.stream)
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ts.sort()
            ms = ts[len(ts) // 2]
            r.update(ms_per_step=round(ms, 3), frames_per_s=round(B / ms * 1e3, 1), mpoints_per_s=round(N / ms / 1e3, 1))
            pts, off = eng.points[:N], eng.offsets
            table, pp = ops.pillarize(pts, off, B, H, H, pcr[0], pcr[1], ps)
            M = table.count()
            t1 = graph_time(lambda: ops.pillarize(pts, off, B, H, H, pcr[0], pcr[1], ps), iters=8)
            t2 = graph_time(lambda: ops.pfn_scatter_max(pts, pp, table, pcr[0], pcr[1], ps, ps / 2 + pcr[0], ps / 2 + pcr[1],
                                                        w, sc, sh, want_bf16=True, want_f32=False), iters=8)
            r.update(pillars=M, pillarize_us=round(t1, 1), pfn_us=round(t2, 1),
                     pillarize_GBs=round((24 * N + 12 * M) / t1 / 1e3, 1), pfn_GBs=round((24 * N + 64 * M) / t2 / 1e3, 1),
                     pillarize_frac_hbm=round((24 * N + 12 * M) / t1 / 1e3 / peaks["hbm_gbs"], 3),
                     pfn_frac_hbm=round((24 * N + 64 * M) / t2 / 1e3 / peaks["hbm_gbs"], 3))
            del eng
        except RuntimeError as e:   # out of memory at the largest corner: record and go on
            r["error"] = str(e)[:120]
        torch.cuda.empty_cache()
        rows.append(r)
        print(json.dumps(r), flush=True)
json.dump(dict(peaks=peaks, rows=rows), open(os.path.join(ROOT, args.out), "w"), indent=1)
