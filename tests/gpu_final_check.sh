#!/bin/bash
# End-of-round check on one B200: smoke(), the whole GPU suite, both bench arms, wider tile groups.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d.get("stages_ms_per_launch"), d.get("cpu_baseline"))
except Exception as ex:
    print("failed", ex)
PY
}
echo "== bench (default flags)"; timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; show gpurun_out/bench_final.json
echo "== bench --impl reference"; timeout 400 python bench.py --impl reference > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err; show gpurun_out/bench_final_reference.json
for c in config5 config4; do
  for sh in 10 11 8; do
    echo "== bench $c DGE_PART2_SHIFT=$sh"
    DGE_PART2_SHIFT=$sh timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${c}_shift$sh.json 2> gpurun_out/bench_${c}_shift$sh.err; show gpurun_out/bench_${c}_shift$sh.json
  done
done
