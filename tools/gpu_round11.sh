#!/bin/bash
# conv parity tests, stall timeline, headline bench, warm per-kernel profile
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rulebook_conv.py tests/test_gpu_model.py tests/test_gpu_dense_conv.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/pytest_conv.log
tail -4 gpurun_out/pytest_conv.log
timeout 300 python tools/infer_timeline.py > gpurun_out/timeline.log 2>&1; grep -A200 "pass 2" gpurun_out/timeline.log | grep -A1 "conv_tc" | cut -c1-330 | head -60
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_nusc18.json 2> gpurun_out/bench_err.log
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_nusc18.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','launches_per_step','stages_us','detections_last_step')}, d['e2e']['value'], d['roofline'])
"; tail -3 gpurun_out/bench_err.log
PN_PDL=0 timeout 400 python tools/prof_infer.py 2>&1 | tail -27 | head -14
