#!/bin/bash
# A/B of blend-backward variants built with `python -m dge_b200.build --variant NAME -D...` (selected with
# DGE_B200_LIB): step time at config 2 and the gradient parity tests per variant. Run under gpurun.
set -u
mkdir -p gpurun_out
for v in "" r12 fexp both; do
  lib=dge_b200/_build/${v:+var_$v/}libdge_b200.so
  echo "== variant '${v:-default}' ($lib)"
  DGE_B200_LIB=$PWD/$lib timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bwd_${v:-default}.json 2> gpurun_out/bench_bwd_${v:-default}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_bwd_${v:-default}.json").read().strip().splitlines()[-1])
    print(round(d["value"], 1), round(d["ms_per_step"], 3), "bwd", round(d["stages_ms_per_launch"]["render_bwd"], 4), "fwd", round(d["stages_ms_per_launch"]["render_fwd"], 4))
except Exception as ex:
    print("failed", ex)
PY
done
for v in both; do
  echo "== parity tests, variant $v"
  DGE_B200_LIB=$PWD/dge_b200/_build/var_$v/libdge_b200.so timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
done
