#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/pytest_all.log
tail -4 gpurun_out/pytest_all.log
PN_PROF_TIMELINE=1 timeout 400 python tools/prof_infer.py > gpurun_out/prof_timeline.log 2>&1; grep "kernel time" gpurun_out/prof_timeline.log; grep -A200 "timeline of one replay" gpurun_out/prof_timeline.log | head -${1:-58}
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_nusc18.json 2> gpurun_out/bench_err.log
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_nusc18.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','launches_per_step','stages_us','detections_last_step')}, d['e2e']['value'])
"; tail -3 gpurun_out/bench_err.log
