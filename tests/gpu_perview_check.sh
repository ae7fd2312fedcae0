#!/bin/bash
# Per-view family after a change: whole GPU suite, the drop-in number, the per-view mask back-projection loop,
# the default bench. Run under gpurun from the repo root.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d.get("stages_ms_per_launch"))
except Exception as ex:
    print("failed", ex)
PY
}
echo "== bench --dropin"; timeout 300 python bench.py --dropin --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dropin2.json 2> gpurun_out/bench_dropin2.err; show gpurun_out/bench_dropin2.json
echo "== bench --dropin, DGE_NO_PARTITION=1"; DGE_NO_PARTITION=1 timeout 300 python bench.py --dropin --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dropin2_nopart.json 2> gpurun_out/bench_dropin2_nopart.err; show gpurun_out/bench_dropin2_nopart.json
echo "== bench --streams 4 (per-view C-ABI, 4 streams)"; timeout 300 python bench.py --streams 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_streams4.json 2> gpurun_out/bench_streams4.err; show gpurun_out/bench_streams4.json
echo "== bench config3 per-view loop"; timeout 300 python bench.py --config config3 --streams 1 --steps 5 --warmup 3 > gpurun_out/bench_config3_perview.json 2> gpurun_out/bench_config3_perview.err; show gpurun_out/bench_config3_perview.json
echo "== bench default"; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default2.json 2> gpurun_out/bench_default2.err; show gpurun_out/bench_default2.json
