#!/bin/bash
# One point of the driver's scaling run on N GPUs: both arms launched the way the driver launches them.
mkdir -p gpurun_out
N=${1:-4}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5"
timeout 600 $RUN --impl reference > gpurun_out/scale${N}_reference.json 2> gpurun_out/scale${N}_reference.err; echo "reference rc=$?"
timeout 900 $RUN > gpurun_out/scale${N}_ours.json 2> gpurun_out/scale${N}_ours.err; echo "ours rc=$?"
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for arm in ("reference", "ours"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/scale{n}_{arm}.json").read().splitlines() if l.startswith("{")][-1])
        print(arm, f"value {d['value']:.0f} e2e {d['e2e']['value']:.0f} ms/step {d['ms_per_step']:.3f} n_gpus {d['n_gpus']}", d.get("nccl"))
        if arm == "ours":
            print("  extras", {k: round(v["value"]) for k, v in d.get("extras", {}).items()})
    except Exception as ex:
        print(arm, "failed:", ex, open(f"gpurun_out/scale{n}_{arm}.err").read()[-600:])
PY
