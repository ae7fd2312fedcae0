#!/bin/bash
# Register-resident scatter for <= 32 buckets (part_scatter_small_kernel): parity with the partition forced
# on every tile count, then binning time at configs 5 / 4 against the shared-memory kernel (DGE_PART_NO_SMALL)
# and the 80-register build; finally the default bench (new JSON keys). Run under gpurun.
set -u
mkdir -p gpurun_out
T="tests/test_fit_gpu.py"
K="bit_identical or empty_views or backprojection or semantic"
echo "== forced two-level, 16-tile groups"; DGE_PART2=2 DGE_PART2_SHIFT=4 timeout 300 python -m pytest $T -x -q -m gpu -k "$K" 2>&1 | tail -2
echo "== forced two-level, 256-tile groups"; DGE_PART2=2 timeout 240 python -m pytest $T -x -q -m gpu -k "$K" 2>&1 | tail -2
echo "== default path, whole file"; timeout 420 python -m pytest $T -x -q -m gpu 2>&1 | tail -2
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"], 1), round(d["ms_per_step"], 3), "binning", round(d["stages_ms_per_launch"]["binning"], 3), d.get("rates"), (d.get("roofline") or {}).get("issue"))
except Exception as ex:
    print("failed", ex)
PY
}
for c in config5 config4; do
  echo "== $c small (64 regs)"; timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ss_$c.json 2> gpurun_out/ss_$c.err; show gpurun_out/ss_$c.json
  echo "== $c small (80 regs)"; DGE_B200_LIB=$PWD/dge_b200/_build/var_small80/libdge_b200.so timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ss80_$c.json 2> gpurun_out/ss80_$c.err; show gpurun_out/ss80_$c.json
  echo "== $c shared-memory kernel"; DGE_PART_NO_SMALL=1 timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ssno_$c.json 2> gpurun_out/ssno_$c.err; show gpurun_out/ssno_$c.json
done
echo "== bench default"; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default3.json 2> gpurun_out/bench_default3.err; show gpurun_out/bench_default3.json; tail -2 gpurun_out/bench_default3.err
