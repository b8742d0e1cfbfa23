#!/bin/bash
# One parameterised GPU-box job (replaces the per-round scripts of round 1).  Usage, under gpurun:
#   bash tools/gpu_job.sh <step> [<step> ...]
# steps:
#   tests            python -m pytest tests -m gpu                      -> gpurun_out/pytest_gpu.log
#   tests:<expr>     ... -k <expr>
#   smoke            __graft_entry__.smoke()
#   bench            headline bench line                                 -> gpurun_out/bench_nusc18.json
#   bench_ref        the CPU reference arm                               -> gpurun_out/bench_reference.json
#   bench_waymo      waymo34 batch 8 / batch 1                           -> gpurun_out/bench_waymo34_b{8,1}.json
#   bench_train      nusc34 training step, 4 frames                      -> gpurun_out/bench_train_nusc34_b4.json
#   launches         ncu launch list of one graph replay                 -> gpurun_out/launches.csv (+ .txt summary)
#   ncu_mem          ncu --set full of the memory-bound kernels          -> gpurun_out/prof_mem.ncu-rep
#   ncu_k:<regex>    ncu --set full of the kernels matching <regex>           -> gpurun_out/prof_k.ncu-rep
#   ncu_convs        ncu --set full of the conv kernels                  -> gpurun_out/prof_convs.ncu-rep
#   timeline         per-CTA pipeline timelines + stall accounting of every conv launch -> gpurun_out/timeline.txt
#   kbench_head      the CenterHead conv pair in isolation (events)      -> gpurun_out/kbench_head.json
#   kbench_reader    reader kernels at scale (events)                    -> gpurun_out/kbench_reader.json
# Every ncu step first runs the same command plain (the recipe's rule) and only profiles if that exited 0.
mkdir -p gpurun_out
PROF_CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-extra --sustain-s 0 --profile-pass"
plain_ok=0
plain() {
  if [ $plain_ok -eq 0 ]; then
    timeout 600 $PROF_CMD > gpurun_out/plain.log 2>&1 && plain_ok=1
    echo "plain profile command ok=$plain_ok"
  fi
}
for step in "$@"; do
  echo "=== $step"
  case "$step" in
    tests)
      timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
      echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log ;;
    tests:*)
      timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider -k "${step#tests:}" > gpurun_out/pytest_gpu_k.log 2>&1
      echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_k.log ;;
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log ;;
    bench)
      timeout 1200 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_nusc18.json 2> gpurun_out/bench_err.log
      echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_nusc18.json; tail -3 gpurun_out/bench_err.log ;;
    bench_quick)
      timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-parity --no-extra --sustain-s 0 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick_err.log
      echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick_err.log ;;
    bench_ref)
      timeout 1200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_ref_err.log
      echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_reference.json ;;
    bench_waymo)
      timeout 900 python bench.py --workload waymo34 --frames-per-step 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_waymo34_b8.json 2> gpurun_out/bench_waymo_err.log
      echo "waymo b8 rc=$?"; cut -c1-300 gpurun_out/bench_waymo34_b8.json
      timeout 900 python bench.py --workload waymo34 --frames-per-step 1 --steps 30 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_waymo34_b1.json 2>> gpurun_out/bench_waymo_err.log
      echo "waymo b1 rc=$?"; cut -c1-300 gpurun_out/bench_waymo34_b1.json ;;
    bench_train)
      timeout 900 python bench.py --mode train --workload nusc34 --frames-per-step 4 --steps 20 --warmup 3 > gpurun_out/bench_train_nusc34_b4.json 2> gpurun_out/bench_train_err.log
      echo "train rc=$?"; cut -c1-400 gpurun_out/bench_train_nusc34_b4.json; tail -3 gpurun_out/bench_train_err.log ;;
    launches)
      plain
      [ $plain_ok -eq 1 ] && timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
        --csv --log-file gpurun_out/launches.csv $PROF_CMD > gpurun_out/ncu_launches.log 2>&1
      python tools/ncu_agg.py gpurun_out/launches.csv > gpurun_out/launches.txt 2>&1; head -40 gpurun_out/launches.txt ;;
    ncu_mem)
      plain
      [ $plain_ok -eq 1 ] && timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on \
        -k 'regex:k_mark|k_scan_emit|k_rank|k_pfn|k_zero_rows|k_decode_candidates|k_select_topk|k_nms_|k_sparse_to_dense|k_subm_nbr|k_down_mask|k_pyramid_nbr' \
        -o gpurun_out/prof_mem -f $PROF_CMD > gpurun_out/ncu_mem.log 2>&1
      echo "ncu_mem rc=$?"; tail -3 gpurun_out/ncu_mem.log ;;
    ncu_k:*)
      plain
      [ $plain_ok -eq 1 ] && timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on \
        -k "regex:${step#ncu_k:}" -o gpurun_out/prof_k -f $PROF_CMD > gpurun_out/ncu_k.log 2>&1
      echo "ncu_k rc=$?"; tail -3 gpurun_out/ncu_k.log | cut -c1-200 ;;
    ncu_convs)
      plain
      [ $plain_ok -eq 1 ] && timeout 2400 ncu --profile-from-start off --set full --clock-control none --import-source on \
        -k 'regex:k_conv' -o gpurun_out/prof_convs -f $PROF_CMD > gpurun_out/ncu_convs.log 2>&1
      echo "ncu_convs rc=$?"; tail -3 gpurun_out/ncu_convs.log ;;
    timeline)
      timeout 600 python tools/infer_timeline.py > gpurun_out/timeline.log 2> gpurun_out/timeline.txt; echo "timeline rc=$?"
      grep -A200 "pass 2" gpurun_out/timeline.txt | grep -A1 "conv_win\|conv_tc" | head -120 ;;
    kbench_head)
      timeout 300 python tools/kbench_head.py 2>&1 | tail -5 ;;
    kbench_reader)
      timeout 900 python tools/kbench_reader.py > gpurun_out/kbench_reader.log 2>&1; echo "kbench rc=$?"; tail -5 gpurun_out/kbench_reader.log ;;
    *)
      echo "unknown step $step" ;;
  esac
done
