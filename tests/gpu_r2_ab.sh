#!/bin/bash
# A/B of kernel build variants (python -m dge_b200.build --variant NAME -D...): the GPU suite on the default
# build, then bench.py's stage times for the default and every variant under dge_b200/_build/var_*.
mkdir -p gpurun_out
if [ "$1" != "notest" ]; then echo "== pytest -m gpu (default build)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8; fi
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    s = d.get("stages_ms_per_launch", {})
    print(f"  value {d['value']:.0f} views/s ({d['ms_per_step']:.3f} ms/step, median {d['timing']['resident']['ms_per_step_median']:.3f}), e2e {d['e2e']['value']:.0f}; " + ", ".join(f"{k} {v:.3f}" for k, v in s.items()))
except Exception as ex:
    print("  failed:", ex, open(sys.argv[1]).read()[-300:])
PY
}
echo "== default"; timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/ab_default.json 2> gpurun_out/ab_default.err; show gpurun_out/ab_default.json
for d in dge_b200/_build/var_*/; do
  n=$(basename $d); n=${n#var_}
  echo "== $n"; DGE_B200_LIB=$d/libdge_b200.so timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/ab_$n.json 2> gpurun_out/ab_$n.err; show gpurun_out/ab_$n.json
done
