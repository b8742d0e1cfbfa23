#!/bin/bash
mkdir -p gpurun_out
PN_PROF_TIMELINE=1 timeout 400 python tools/prof_infer.py > gpurun_out/prof_timeline.log 2>&1; grep -A200 "timeline of one replay" gpurun_out/prof_timeline.log | head -120
