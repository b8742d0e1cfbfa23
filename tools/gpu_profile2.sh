#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-pass"
timeout 600 $CMD > gpurun_out/plain.json 2> gpurun_out/plain_err.log &&
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_conv_tc -s 6 -c 8 -o gpurun_out/prof_tc -f $CMD > gpurun_out/ncu_b.log 2>&1 &&
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_conv_dense -s 1 -c 4 -o gpurun_out/prof_dense -f $CMD > gpurun_out/ncu_c.log 2>&1
echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
