"""Per-stage CUDA-event timings of the hot-path kernels (development aid; numbers for DESIGN.md/profiles)."""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pillarnet_lts_b200 as P  # noqa: E402
from pillarnet_lts_b200 import ops, synth  # noqa: E402


def timeit(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "nuscenes"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    cfg = dict(nuscenes=dict(pcr=[-54, -54, -5.0, 54, 54, 3.0], ps=0.075, H=1440),
               waymo=dict(pcr=[-75.2, -75.2, -2, 75.2, 75.2, 4], ps=0.1, H=1504))[kind]
    frames = synth.make_batch(kind, B, 0)
    counts = np.cumsum([0] + [len(f) for f in frames]).astype(np.int32)
    pts = torch.from_numpy(np.concatenate(frames)).cuda()
    off = torch.from_numpy(counts).cuda()
    H = W = cfg["H"]
    pcr, ps = cfg["pcr"], cfg["ps"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {"kind": kind, "B": B, "N": int(pts.shape[0])}
    table, pp = ops.pillarize(pts, off, B, H, W, pcr[0], pcr[1], ps)
    M = table.count()
    res["M"] = M
    res["pillarize_us"] = timeit(lambda: ops.pillarize(pts, off, B, H, W, pcr[0], pcr[1], ps), flush=flush)
    w = torch.randn(32, 7, device="cuda")
    sc, sh = torch.ones(32, device="cuda"), torch.zeros(32, device="cuda")
    res["pfn_scatter_us"] = timeit(lambda: ops.pfn_scatter_max(pts, pp, table, pcr[0], pcr[1], ps, ps / 2 + pcr[0],
                                                              ps / 2 + pcr[1], w, sc, sh, want_bf16=True), flush=flush)
    N = pts.shape[0]
    res["pillarize_GBs"] = (20 * N + 4 * N + 12 * M) / res["pillarize_us"] / 1e3
    res["pfn_GBs"] = (24 * N + 128 * M) / res["pfn_scatter_us"] / 1e3
    res["rulebook_subm_us"] = timeit(lambda: ops.rulebook_subm3x3(table), flush=flush)
    res["rulebook_down_us"] = timeit(lambda: ops.rulebook_down3x3s2(table), flush=flush)
    t2, nbr2 = ops.rulebook_down3x3s2(table)
    res["M2"] = t2.count()
    # one SubM conv at stage 1, both impls
    for impl, name in ((0, "simt"), (1, "tc")):
        try:
            dt = torch.float32 if impl == 0 else torch.bfloat16
            x = torch.randn(table.cap, 32, device="cuda").to(dt)
            wt = torch.randn(32, 288, device="cuda")
            wt = wt if impl == 0 else ops.pack_weight_bf16(wt)
            out = torch.empty(table.cap, 32, device="cuda", dtype=dt)
            nb = table.subm_nbr()
            res[f"subm32_{name}_us"] = timeit(lambda: ops.conv_gather(x, wt, nb, 9, 32, 32, out, num=table.num, impl=impl),
                                              iters=5, flush=flush)
        except RuntimeError as e:
            res[f"subm32_{name}_us"] = str(e)[:60]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
