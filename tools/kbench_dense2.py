"""per-tile time of the dense conv vs number of concurrently active CTAs (is the feed limit per-SM or shared?)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pillarnet_lts_b200 as P  # noqa
from pillarnet_lts_b200 import ops
from tools.kbench_dense import timeit  # noqa
for cin, cout in [(256, 256), (128, 128)]:
    for H in (14, 30, 46, 62, 90, 126, 180, 254, 360):
        for hint in (0x802, 0x202, 0x804):
            rows = torch.randn((H + 2) * (H + 2), cin, device="cuda").to(torch.bfloat16)
            w = ops.pack_weight_bf16(torch.randn(cout, 9 * cin, device="cuda") * 0.02)
            out = torch.empty((H + 2) * (H + 2), cout, device="cuda", dtype=torch.bfloat16)
            us = timeit(lambda: ops.conv_dense3x3(rows, 0, cin, 1, H, H, w, cout, out, relu=True, tile_hint=hint))
            mt = 2 if (hint & 0xF) in (1, 3) else 1
            bn = 256 if (hint & 0xF) in (1, 2) else 128
            tiles = -(-((H + 2) ** 2) // (128 * mt)) * -(-cout // bn)
            print(dict(cin=cin, cout=cout, H=H, hint=hex(hint), tiles=tiles, waves=round(tiles / 148, 2), us=round(us, 1),
                       tflops=round(2.0 * H * H * 9 * cin * cout / us / 1e6, 1)), flush=True)
