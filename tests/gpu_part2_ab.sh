#!/bin/bash
# A/B of the two-level tile partition (binning.cu, DGE_PART2) on one B200: bit-identity of the lists
# against the per-view API with the partition forced on every tile count, then the step time of the
# configurations with more than 2048 tiles with and without it. Run under gpurun from the repo root.
set -u
mkdir -p gpurun_out
T="tests/test_fit_gpu.py"
K="bit_identical or empty_views or backprojection or semantic"
echo "== forced two-level, 16-tile groups"; DGE_PART2=2 DGE_PART2_SHIFT=4 timeout 300 python -m pytest $T -x -q -m gpu -k "$K" 2>&1 | tail -5
echo "== forced two-level, 256-tile groups"; DGE_PART2=2 timeout 240 python -m pytest $T -x -q -m gpu -k "$K" 2>&1 | tail -5
echo "== default path, whole file"; timeout 420 python -m pytest $T -x -q -m gpu 2>&1 | tail -5
echo "== bench config2"; timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_config2.json 2> gpurun_out/bench_config2.err; tail -c 600 gpurun_out/bench_config2.json
for c in config4 config5; do
  for m in 0 1; do
    echo "== bench $c DGE_PART2=$m"
    DGE_PART2=$m timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${c}_part2_$m.json 2> gpurun_out/bench_${c}_part2_$m.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${c}_part2_$m.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d.get("stages_ms_per_launch"))
except Exception as ex:
    print("failed", ex)
PY
  done
done
echo "== ncu launch list (config2)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1h.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log | cut -c1-300
for m in 0 1; do
  echo "== bench config5 6M DGE_PART2=$m"
  DGE_PART2=$m timeout 400 python bench.py --config config5 --gaussians 6000000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config5_6M_part2_$m.json 2> gpurun_out/bench_config5_6M_part2_$m.err
  tail -c 400 gpurun_out/bench_config5_6M_part2_$m.json
done
