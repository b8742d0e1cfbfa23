#!/bin/bash
# dense conv parity + A/B of an env switch given as $1 (e.g. PN_DENSE_DBGMODE=8) on the warm per-kernel profile and the bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense_conv.py tests/test_gpu_model.py tests/test_gpu_head.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for mode in "" "$1"; do
  echo "=== env: '$mode'"
  env $mode PN_PDL=0 timeout 400 python tools/prof_infer.py 2>&1 | grep -E "kernel time|k_conv_dense|k_conv_tc" | head -12
  env $mode timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done
