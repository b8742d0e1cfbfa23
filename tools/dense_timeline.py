import os, sys
os.environ["PN_DENSE_TIMELINE"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pillarnet_lts_b200 as P  # noqa
from pillarnet_lts_b200 import ops
for (H, cin, cout) in [(180, 256, 256), (90, 256, 256), (180, 512, 256), (180, 64, 2304), (376, 128, 128)]:
    rows = torch.randn((H + 2) * (H + 2), cin, device="cuda").to(torch.bfloat16)
    w = ops.pack_weight_bf16(torch.randn(cout, 9 * cin, device="cuda") * 0.02)
    out = torch.empty((H + 2) * (H + 2), cout, device="cuda", dtype=torch.bfloat16)
    print(f"---- H={H} cin={cin} cout={cout}", file=sys.stderr)
    for hint in (0x801, 0x1001, 0x802, 0x1002, 0x803, 0x1003, 0x804, 0x1004):
        if cout <= 128 and (hint & 0xF) in (1, 2):
            continue
        for _ in range(2):
            ops.conv_dense3x3(rows, 0, cin, 1, H, H, w, cout, out, relu=True, tile_hint=hint)
        torch.cuda.synchronize()
