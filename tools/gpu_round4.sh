#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_all.log
tail -8 gpurun_out/pytest_all.log
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --profile-pass"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nusc18.json 2> gpurun_out/bench_err.log
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_nusc18.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','launches_per_step','stages_us','detections_last_step')}, d['e2e'], d['roofline'])
b=json.load(open('gpurun_out/breakdown_nusc18_n1.json'))
for r in b['convs']: print({k:(round(v,1) if isinstance(v,float) else v) for k,v in r.items()})
"; tail -5 gpurun_out/bench_err.log
timeout 600 $CMD > /dev/null 2>&1 && timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
