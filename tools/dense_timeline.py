import os, sys
os.environ["PN_DENSE_TIMELINE"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pillarnet_lts_b200 as P  # noqa
from pillarnet_lts_b200 import ops
for (H, cin, cout, hint) in [(180, 256, 256, 0x801), (180, 256, 256, 0x802), (180, 128, 128, 0x804), (90, 256, 256, 0x804),
                             (180, 64, 2304, 0x803), (360, 256, 256, 0x802)]:
    rows = torch.randn((H + 2) * (H + 2), cin, device="cuda").to(torch.bfloat16)
    w = ops.pack_weight_bf16(torch.randn(cout, 9 * cin, device="cuda") * 0.02)
    out = torch.empty((H + 2) * (H + 2), cout, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.conv_dense3x3(rows, 0, cin, 1, H, H, w, cout, out, relu=True, tile_hint=hint)
    torch.cuda.synchronize()
    print("----", H, cin, cout, hex(hint), file=sys.stderr)
