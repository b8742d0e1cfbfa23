#!/bin/bash
# final-state evidence: ncu --set full of every conv launch of one profiled pass, then the config-5 sweep
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-pass"
PN_PDL=0 timeout 600 $CMD > /dev/null 2>&1 && PN_PDL=0 timeout 1200 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_conv_dense|k_conv_tc" -o gpurun_out/prof_convs_r1 -f $CMD > gpurun_out/ncu_convs.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_convs.log
python tools/ncu_convs_summary.py gpurun_out/prof_convs_r1.ncu-rep > gpurun_out/ncu_full_convs.json; echo "summary rc=$?"
timeout 900 python tools/sweep.py > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"; tail -3 gpurun_out/sweep.log | cut -c1-250
