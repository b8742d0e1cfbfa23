#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_reader.py tests/test_gpu_model.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 600 python tools/kbench_reader.py 2>&1 | tail -5
cat > /tmp/rd.py <<'PY'
import sys, os, numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import pillarnet_lts_b200
from pillarnet_lts_b200 import ops, synth
base = synth.make_batch("nuscenes", 2, 0, 2_000_000)
frames = [base[i % 2] for i in range(8)]
counts = np.cumsum([0] + [len(f) for f in frames]).astype(np.int32)
pts = torch.from_numpy(np.concatenate(frames)).cuda(); off = torch.from_numpy(counts).cuda()
w = torch.randn(32, 7); sc = torch.ones(32); sh = torch.zeros(32)
for _ in range(2):
    table, pp = ops.pillarize(pts, off, 8, 1440, 1440, -54.0, -54.0, 0.075)
    ops.pfn_scatter_max(pts, pp, table, -54.0, -54.0, 0.075, 0.0375 - 54, 0.0375 - 54, w, sc, sh, want_bf16=True)
    ops.rulebook_subm3x3(table); ops.rulebook_down3x3s2(table)
torch.cuda.synchronize()
PY
timeout 300 python /tmp/rd.py && timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/reader_launches.csv python /tmp/rd.py > /dev/null 2>&1
echo rc=$?
