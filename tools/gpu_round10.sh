#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/pytest_all.log
tail -4 gpurun_out/pytest_all.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_nusc18.json 2> gpurun_out/bench_err.log
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_nusc18.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','launches_per_step','stages_us','detections_last_step')}, d['e2e']['value'], d['roofline'])
b=json.load(open('gpurun_out/breakdown_nusc18_n1.json'))
for r in b['convs'][:14]: print({k:(round(v,1) if isinstance(v,float) else v) for k,v in r.items()})
"; tail -3 gpurun_out/bench_err.log
PN_PDL=0 timeout 400 python tools/prof_infer.py 2>&1 | tail -27 | head -12
