#!/bin/bash
# memcheck on the smallest end-to-end case (one tool per call; see B200_PROFILING.md)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_plain.log 2>&1 && \
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/smoke_plain.log; tail -15 gpurun_out/sanitizer_memcheck.log
