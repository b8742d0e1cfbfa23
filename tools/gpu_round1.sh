#!/bin/bash
# first GPU validation: parity tests of the SIMT paths + per-kernel timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "not tcgen05 and not bf16" -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_simt.log
cat gpurun_out/pytest_simt.log | tail -40
timeout 300 python tools/kbench.py nuscenes 1 > gpurun_out/kbench_nusc1.json 2> gpurun_out/kbench_err.log; cat gpurun_out/kbench_nusc1.json; tail -5 gpurun_out/kbench_err.log
timeout 300 python tools/kbench.py waymo 8 > gpurun_out/kbench_waymo8.json 2>> gpurun_out/kbench_err.log; cat gpurun_out/kbench_waymo8.json
