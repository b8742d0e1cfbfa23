import os, sys
os.environ["PN_DENSE_TIMELINE"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pillarnet_lts_b200 as P  # noqa
from pillarnet_lts_b200 import ops
import sys as _s
H, cin, cout, hint = 180, int(_s.argv[2]) if len(_s.argv) > 2 else 256, 256, int(_s.argv[1], 16)
rows = torch.randn((H + 2) * (H + 2), cin, device="cuda").to(torch.bfloat16)
w = ops.pack_weight_bf16(torch.randn(cout, 9 * cin, device="cuda") * 0.02)
out = torch.empty((H + 2) * (H + 2), cout, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.conv_dense3x3(rows, 0, cin, 1, H, H, w, cout, out, relu=True, tile_hint=hint)
torch.cuda.synchronize()
