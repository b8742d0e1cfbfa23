#!/bin/bash
# round-end validation: tests, headline bench (+ reference arm), other workloads, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/pytest_all.log; tail -2 gpurun_out/pytest_all.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_nusc18.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_nusc18.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_err.log; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 600 python bench.py --workload waymo34 --frames-per-step 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_waymo34_b8.json 2>> gpurun_out/bench_err.log; echo "waymo rc=$?"; cut -c1-200 gpurun_out/bench_waymo34_b8.json
timeout 600 python bench.py --workload waymo34 --frames-per-step 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_waymo34_b1.json 2>> gpurun_out/bench_err.log; echo "waymo1 rc=$?"; cut -c1-200 gpurun_out/bench_waymo34_b1.json
timeout 600 python bench.py --mode train --workload nusc34 --frames-per-step 4 --steps 5 --warmup 3 > gpurun_out/bench_train34.json 2>> gpurun_out/bench_err.log; echo "train rc=$?"; cut -c1-300 gpurun_out/bench_train34.json
PN_PDL=0 timeout 400 python tools/prof_infer.py > gpurun_out/prof_infer.log 2>&1; tail -28 gpurun_out/prof_infer.log | head -12
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-pass"
timeout 600 $CMD > /dev/null 2>&1 && timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"
tail -4 gpurun_out/bench_err.log
