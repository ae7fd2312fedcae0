#!/bin/bash
# two GPUs: the NCCL test of the suite, then bench.py under torchrun with and without the prefetched front half
mkdir -p gpurun_out
echo "== pytest (2 GPUs visible)"; timeout 900 python -m pytest tests/test_parity_full_gpu.py tests/test_fit_gpu.py -m gpu -x -q -k "two_ranks or prefetched" -s 2>&1 | tail -15
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    s = d.get("stages_ms_per_launch", {})
    print(f"  value {d['value']:.0f} views/s ({d['ms_per_step']:.3f} ms/step, median {d['timing']['resident']['ms_per_step_median']:.3f}), e2e {d['e2e']['value']:.0f}; " + ", ".join(f"{k} {v:.3f}" for k, v in s.items()))
    if "extras" in d: print("  extras:", {k: (round(v.get("value", 0)), round(v.get("ms_per_step", 0), 2)) if isinstance(v, dict) else v for k, v in d["extras"].items()})
except Exception as ex:
    print("  failed:", ex, open(sys.argv[1]).read()[-300:], open(sys.argv[1].replace('.json','.err')).read()[-900:])
PY
}
N=${1:-2}
for p in 0 1; do
  echo "== $N GPUs --pipeline $p"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --pipeline $p > gpurun_out/g${N}_pipe$p.json 2> gpurun_out/g${N}_pipe$p.err; show gpurun_out/g${N}_pipe$p.json
done
echo "== $N GPUs default with extras"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/g${N}_full.json 2> gpurun_out/g${N}_full.err; show gpurun_out/g${N}_full.json
