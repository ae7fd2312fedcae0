#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    s = d.get("stages_ms_per_launch", {})
    print(f"  value {d['value']:.0f} views/s ({d['ms_per_step']:.3f} ms/step, median {d['timing']['resident']['ms_per_step_median']:.3f}), e2e {d['e2e']['value']:.0f}; " + ", ".join(f"{k} {v:.3f}" for k, v in s.items()))
except Exception as ex:
    print("  failed:", ex, open(sys.argv[1]).read()[-300:], open(sys.argv[1].replace('.json','.err')).read()[-600:])
PY
}
for p in 0 1; do echo "== 1 GPU --pipeline $p"; timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 10 --warmup 3 --pipeline $p > gpurun_out/pipe$p.json 2> gpurun_out/pipe$p.err; show gpurun_out/pipe$p.json; done
