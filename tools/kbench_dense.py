"""CUDA-event timing of pn_conv_dense3x3 per tile shape (development aid)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pillarnet_lts_b200 as P  # noqa
from pillarnet_lts_b200 import ops


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    res = []
    for (H, cin, cout) in [(180, 256, 256), (90, 256, 256), (180, 64, 2304), (180, 512, 256), (180, 256, 64)]:
        rows = torch.randn((H + 2) * (H + 2), cin, device="cuda").to(torch.bfloat16)
        w = ops.pack_weight_bf16(torch.randn(cout, 9 * cin, device="cuda") * 0.02)
        out = torch.empty((H + 2) * (H + 2), cout, device="cuda", dtype=torch.bfloat16)
        for hint in (0, 1, 2, 3, 4, 0x1001, 0x1002, 0x1003, 0x1004, 0x203):   # 0x1xxx: cta_group::2 pair, 0x2xx: weight multicast over 2
            if (hint & 0xF) in (1, 2) and cout <= 128:
                continue
            us = timeit(lambda: ops.conv_dense3x3(rows, 0, cin, 1, H, H, w, cout, out, relu=True, tile_hint=hint))
            res.append(dict(H=H, cin=cin, cout=cout, hint=hex(hint), us=round(us, 1),
                            tflops=round(2.0 * H * H * 9 * cin * cout / us / 1e6, 1)))
            print(res[-1], flush=True)
    json.dump(res, open("gpurun_out/kbench_dense.json", "w"))


if __name__ == "__main__":
    main()
