#!/bin/bash
# config-5 sweep + ncu counters for the reader kernels at scale and the sparse conv at waymo B=8
mkdir -p gpurun_out
timeout 900 python tools/sweep.py > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"; tail -14 gpurun_out/sweep.log | cut -c1-330
timeout 600 python bench.py --workload waymo34 --frames-per-step 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_waymo34_b8.json 2> gpurun_out/bench_err.log; echo "waymo rc=$?"; cut -c1-400 gpurun_out/bench_waymo34_b8.json
# ncu: reader kernels at 8 x 2M points, sparse + dense conv at waymo34 B=8 (one profiled pass)
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_mark|k_pfn_scatter_max_bf16|k_emit|k_rank" -c 8 -o gpurun_out/prof_reader_r1 -f python tools/kbench_reader.py > gpurun_out/ncu_reader.log 2>&1; echo "ncu reader rc=$?"
