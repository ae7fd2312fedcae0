#!/bin/bash
# compute-sanitizer --tool $1 (memcheck | racecheck | synccheck | initcheck) over __graft_entry__.smoke():
# one small forward + backward through the per-view API and a 3-view fit step through both per-step paths
# (TMA-staged blends with their mbarrier, onesweep look-back, partition, moment-sum backward).
# One tool per gpurun call (B200_PROFILING.md).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain smoke failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $1 --log-file gpurun_out/sanitizer_$1.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$1.out 2>&1
echo "rc=$?"; tail -3 gpurun_out/sanitizer_$1.out; tail -15 gpurun_out/sanitizer_$1.log
