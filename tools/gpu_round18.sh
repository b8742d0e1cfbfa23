#!/bin/bash
# round-end validation + final-state ncu --set full capture of every conv launch
bash tools/gpu_final.sh 2>&1 | cut -c1-220
PN_PROF_TIMELINE=1 timeout 300 python tools/prof_infer.py > gpurun_out/prof_timeline.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -1
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-pass"
PN_PDL=0 timeout 1200 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_conv_dense|k_conv_tc" -o gpurun_out/prof_convs_r1 -f $CMD > gpurun_out/ncu_convs.log 2>&1
echo "ncu rc=$?"; python tools/ncu_convs_summary.py gpurun_out/prof_convs_r1.ncu-rep > gpurun_out/ncu_full_convs.json; echo "summary rc=$?"
