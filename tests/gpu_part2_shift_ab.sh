#!/bin/bash
# Two-level partition: tile-group width sweep (DGE_PART2_SHIFT) at 1264x832 / 3M and 1080p / 2M, after the
# bit-identity tests with the partition forced on every tile count. Run under gpurun from the repo root.
set -u
mkdir -p gpurun_out
T="tests/test_fit_gpu.py"
K="bit_identical or empty_views or backprojection or semantic"
echo "== forced two-level, 16-tile groups"; DGE_PART2=2 DGE_PART2_SHIFT=4 timeout 300 python -m pytest $T -x -q -m gpu -k "$K" 2>&1 | tail -3
echo "== forced two-level, 256-tile groups"; DGE_PART2=2 timeout 240 python -m pytest $T -x -q -m gpu -k "$K" 2>&1 | tail -3
echo "== default path, whole file"; timeout 420 python -m pytest $T -x -q -m gpu 2>&1 | tail -3
for c in config5 config4; do
  for sh in 8 7 9 6; do
    echo "== bench $c DGE_PART2_SHIFT=$sh"
    DGE_PART2_SHIFT=$sh timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${c}_shift$sh.json 2> gpurun_out/bench_${c}_shift$sh.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${c}_shift$sh.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d.get("stages_ms_per_launch"))
except Exception as ex:
    print("failed", ex)
PY
  done
done
