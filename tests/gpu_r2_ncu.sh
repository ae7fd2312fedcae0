#!/bin/bash
# ncu --set full of one kernel (regex $2, skip $3 launches) inside a short bench run (plain run of the same command first).
mkdir -p gpurun_out
CMD="python bench.py --no-extras --no-cpu-baseline --steps 2 --warmup 3"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c ${4:-1} -f -o gpurun_out/prof_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$1.log | cut -c1-300
