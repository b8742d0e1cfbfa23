#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_rulebook_conv.py -m gpu -q -x -k "tcgen05" -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_tc.log
cat gpurun_out/pytest_tc.log | tail -40
timeout 600 python -m pytest tests -m gpu -q -k "not tcgen05" -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_rest.log
tail -30 gpurun_out/pytest_rest.log
timeout 300 python tools/kbench.py nuscenes 1 > gpurun_out/kbench_nusc1.json 2> gpurun_out/kbench_err.log; cat gpurun_out/kbench_nusc1.json; tail -5 gpurun_out/kbench_err.log
