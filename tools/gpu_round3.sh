#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_all.log
tail -15 gpurun_out/pytest_all.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_nusc18.json 2> gpurun_out/bench_err.log
echo "bench rc=$?"; cat gpurun_out/bench_nusc18.json; tail -20 gpurun_out/bench_err.log
python -c "
import json; d=json.load(open('gpurun_out/breakdown_nusc18_n1.json'))
print(d['stages_us'])
for r in d['convs'][:14]: print(r)
"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_ref.json
