#!/bin/bash
# training-path validation: new GPU tests first (fail fast), then the full suite, then a short training bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_head.py -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_train.log
tail -40 gpurun_out/pytest_train.log
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_gpu_train.py 2>&1 | tail -8 > gpurun_out/pytest_all.log
tail -4 gpurun_out/pytest_all.log
timeout 600 python bench.py --mode train --workload nusc18 --frames-per-step 2 --steps 5 --warmup 3 > gpurun_out/bench_train18.json 2> gpurun_out/bench_train_err.log
echo "train bench rc=$?"; tail -3 gpurun_out/bench_train_err.log; cut -c1-600 gpurun_out/bench_train18.json
