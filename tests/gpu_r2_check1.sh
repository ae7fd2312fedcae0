#!/bin/bash
# Round-2 first GPU check: the GPU suite, both bench arms with the new harness, then the launch list of the
# reference arm (same command, plain run first).
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_ref.json 2> gpurun_out/r2_ref.err; echo rc=$?; cut -c1-600 gpurun_out/r2_ref.json; tail -5 gpurun_out/r2_ref.err
echo "== bench ours"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_ours.json 2> gpurun_out/r2_ours.err; echo rc=$?; cut -c1-600 gpurun_out/r2_ours.json; tail -5 gpurun_out/r2_ours.err
echo "== ncu launch list of the reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r2_ref_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2_reference.csv python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r2_ref_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_ref_ncu.log
