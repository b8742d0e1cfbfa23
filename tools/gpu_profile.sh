#!/bin/bash
# round profile: launch list of one step + full counters of the top kernel (one ncu family per call)
mkdir -p gpurun_out
CMD="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-pass"
timeout 600 $CMD > gpurun_out/plain.json 2> gpurun_out/plain_err.log &&
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1 &&
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_conv_tc -c 30 -o gpurun_out/prof_conv_tc $CMD > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -c 400 gpurun_out/plain.json; ls -la gpurun_out/*.ncu-rep
