#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_nusc18.json 2> gpurun_out/bench_err.log; echo "rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_nusc18.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','launches_per_step','detections_last_step','clocks')}, d['e2e'], d['roofline'], d['cpu_baseline']['value'])"
timeout 600 python bench.py --workload waymo34 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_waymo34_b1.json 2>> gpurun_out/bench_err.log; echo "rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_waymo34_b1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','launches_per_step','detections_last_step')}, d['e2e'], d['roofline'])"
timeout 900 python bench.py --workload waymo34 --frames-per-step 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_waymo34_b8.json 2>> gpurun_out/bench_err.log; echo "rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_waymo34_b8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','ms_per_frame','launches_per_step','detections_last_step')}, d['e2e'], d['roofline'])"
tail -5 gpurun_out/bench_err.log
