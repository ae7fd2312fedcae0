#!/bin/bash
# End-of-round check on one B200, the way the driver runs things: smoke(), the whole GPU suite, then both bench arms
# with the driver's flags (the reference first). Prints the headline numbers of each line.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    t = d["timing"]
    print(f"  value {d['value']:.1f} (median {t['resident']['value_median']:.1f}) e2e {d['e2e']['value']:.1f} (median {t['e2e']['value_median']:.1f}) "
          f"ms/step {d['ms_per_step']:.3f}; clocks {d['clocks'].get('sm_mhz')} MHz {d['clocks'].get('reasons')}; launches {d.get('gpu_launches')}")
    if d.get("roofline"): print("  roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "avg_launch_ms")})
    if d.get("stages_ms_per_launch"): print("  stages", {k: round(v, 3) for k, v in d["stages_ms_per_launch"].items()})
    for k, v in (d.get("extras") or {}).items():
        print(f"  extras.{k}:", {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "ms_per_step", "ms_per_view", "value_median")} if isinstance(v, dict) else v,
              "e2e", round(v["e2e"]["value"], 1) if isinstance(v, dict) and "e2e" in v else "")
    if d.get("cpu_baseline"): print("  cpu_baseline", d["cpu_baseline"].get("value"), d["cpu_baseline"].get("kind"))
except Exception as ex:
    print("  failed:", ex, open(sys.argv[1].replace(".json", ".err")).read()[-800:])
PY
}
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; show gpurun_out/final_reference.json
echo "== bench"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_ours.json 2> gpurun_out/final_ours.err; show gpurun_out/final_ours.json
