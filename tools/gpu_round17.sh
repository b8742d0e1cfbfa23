#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-pass"
PN_PDL=0 timeout 600 $CMD > /dev/null 2>&1 && PN_PDL=0 timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"k_select_topk|k_nms_mask|k_mark|k_scan_emit" -c 6 -o gpurun_out/prof_decode_r1 -f $CMD > gpurun_out/ncu_decode.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_decode.log
