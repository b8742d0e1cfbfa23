#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile-pass"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "rc=$?"; tail -c 300 gpurun_out/ncu.log; wc -l gpurun_out/launches.csv
