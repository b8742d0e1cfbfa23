"""The CenterHead pair at nuScenes size (180 x 180, 36 branches) in isolation: first-level conv 64 -> 36*64 (planar
output) and the grouped last conv in both forms; CUDA events, L2 flushed before every launch."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pillarnet_lts_b200 import ops  # noqa: E402

dev = torch.device("cuda")
B, H, W, G, hc = 1, 180, 180, 36, 64
n_pos = B * (H + 2) * (W + 2)
g = torch.Generator(device="cuda").manual_seed(0)
feat = torch.zeros(B, H + 2, W + 2, 64, device=dev, dtype=torch.bfloat16)
feat[:, 1:-1, 1:-1] = torch.randn(B, H, W, 64, device=dev, generator=g).to(torch.bfloat16)
rows = feat.view(-1, 64)
w1 = ops.pack_weight_bf16(torch.randn(G * hc, 9 * 64, device=dev, generator=g) * 0.05)
sc, sh = torch.rand(G * hc, device=dev) + 0.5, torch.randn(G * hc, device=dev) * 0.1
planar = torch.empty(G * n_pos, hc, device=dev, dtype=torch.bfloat16)
couts = [(1, 2, 3)[i % 3] for i in range(G)]
tab, col = [], 0
for c in couts:
    tab.append([col, c])
    col += c
tabd = torch.tensor(tab, dtype=torch.int32).to(dev)
w32 = (torch.randn(G * 32, hc, device=dev, generator=g) * 0.1).to(torch.bfloat16)
b4 = torch.randn(G * 4, device=dev)
wg = ops.pack_weight_bf16(torch.randn(G * 16, 9 * hc, device=dev, generator=g) * 0.1)
bg = torch.randn(G * 16, device=dev)
out = torch.empty(B * H * W, col, device=dev)
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
    return ts[len(ts) // 2], ts[0]


res = {}
res["conv1_64_to_2304_planar"] = timeit(lambda: ops.conv_dense3x3(rows, 0, 64, B, H, W, w1, G * hc, planar, scale=sc, shift=sh,
                                                                  relu=True, out_group_cols=hc))
res["conv1 without scale/shift"] = timeit(lambda: ops.conv_dense3x3(rows, 0, 64, B, H, W, w1, G * hc, planar, relu=True,
                                                                    out_group_cols=hc))
res["last_conv_shift"] = timeit(lambda: ops.conv_dense3x3_grouped_shift(planar, G, B, H, W, w32, b4, tabd, out))
res["last_conv_grouped_implicit_gemm"] = timeit(lambda: ops.conv_dense3x3_grouped(planar, 0, hc, G, B, H, W, wg, bg, tabd, out,
                                                                                  out_compact=True, in_planar=True))


def pair(n_chunks):
    """the pair in n_chunks branch chunks through ONE chunk-sized intermediate buffer (stays in L2 when small enough)"""
    gc = G // n_chunks
    buf = planar[:gc * n_pos]

    def run():
        for k in range(n_chunks):
            ops.conv_dense3x3(rows, 0, 64, B, H, W, w1[k * gc * hc:(k + 1) * gc * hc], gc * hc, buf,
                              scale=sc[k * gc * hc:(k + 1) * gc * hc], shift=sh[k * gc * hc:(k + 1) * gc * hc], relu=True,
                              out_group_cols=hc)
            ops.conv_dense3x3_grouped_shift(buf, gc, B, H, W, w32[k * gc * 32:(k + 1) * gc * 32], b4[k * gc * 4:(k + 1) * gc * 4],
                                            tabd[k * gc:(k + 1) * gc], out)
    return run


for nc in (1,):
    res[f"pair_in_{nc}_chunks"] = timeit(pair(nc))
bytes_in = planar.numel() * 2
for k, (med, best) in res.items():
    print(f"{k:36s} median {med:7.1f} us  best {best:7.1f} us   ({bytes_in / med / 1e3:7.0f} GB/s of the intermediate)")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({k: dict(median_us=v[0], best_us=v[1]) for k, v in res.items()}, open(os.path.join(ROOT, "gpurun_out", "kbench_head.json"), "w"), indent=1)
