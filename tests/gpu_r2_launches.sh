#!/bin/bash
# Round-2 launch list of OUR arm: the bench command plain first, then the same command under
# ncu --metrics gpu__time_duration.sum (profiles/launches_r2.md is profiles/summarize.py over the csv).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/launches_plain.json 2> gpurun_out/launches_plain.err || { echo "plain run failed"; tail -5 gpurun_out/launches_plain.err; exit 1; }
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/launches_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_r2.csv
