#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/pytest_all.log
tail -4 gpurun_out/pytest_all.log
for pdl in 0 1; do
  PN_PDL=$pdl timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_err.log
  echo "pdl=$pdl rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_pdl$pdl.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','launches_per_step','detections_last_step')}, d['e2e']['value'])
"; tail -3 gpurun_out/bench_err.log
done
