#!/bin/bash
# Round check on one B200: the whole GPU suite, the drop-in (import swap only) number, and the launch list +
# one full ncu capture of the two-level partition at 1080p. Run under gpurun from the repo root.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== bench config2 --dropin"; timeout 300 python bench.py --dropin --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config2_dropin.json 2> gpurun_out/bench_config2_dropin.err; cut -c1-300 gpurun_out/bench_config2_dropin.json
echo "== bench config2 --dropin --impl reference"; timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_config2_reference.json 2> gpurun_out/bench_config2_reference.err; cut -c1-300 gpurun_out/bench_config2_reference.json
echo "== ncu launch list (config5)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1h_config5.csv \
  python bench.py --config config5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches5.log 2>&1; tail -1 gpurun_out/ncu_launches5.log | cut -c1-200
echo "== ncu full: partition kernels (config5)"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"part_scatter|part_count|expand_kernel|scan_reduce" --launch-skip 12 -c 6 \
  -o gpurun_out/part2_r1h -f python bench.py --config config5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full5.log 2>&1; tail -2 gpurun_out/ncu_full5.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
