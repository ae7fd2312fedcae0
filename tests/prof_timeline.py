"""Profiling driver (not a pytest file): kernel timeline of fit steps of BASELINE.json config 2 with the
multi-stream pipeline ON, captured with torch.profiler (CUPTI activity records cover the kernels this
library launches through ctypes as well). Prints, per kernel name, launches / total / mean duration,
and the step's concurrency figures: wall span, union of busy time, sum of kernel time."""
import argparse, collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from dge_b200 import fit, scene

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=20)
ap.add_argument("--P", type=int, default=1_000_000)
ap.add_argument("--res", type=int, default=512)
ap.add_argument("--streams", type=int, default=0, help="0 = batched path")
ap.add_argument("--out", default="")
ap.add_argument("--pipeline", action="store_true", help="fit_step(next_cameras=cams): prefetch the next step's front half")
ap.add_argument("--timeline", action="store_true", help="print every device activity of the profiled step")
ap.add_argument("--e2e", action="store_true", help="pinned host inputs + loss read back each step; prints the idle gaps")
args = ap.parse_args()
rank = int(os.environ.get("RANK", "0"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if "RANK" in os.environ:  # under torchrun: every rank runs the same 20 views, gradients all-reduced
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
g = scene.make_gaussians(args.P, seed=1236)
cams = [scene.camera_to(c, dev) for c in scene.ring_cameras(20, args.res, args.res)[:args.views]]
gen = torch.Generator().manual_seed(3)
targets = [torch.rand(3, args.res, args.res, generator=gen).to(dev) for _ in range(args.views)]
if args.e2e:
    cams = [scene.Camera(*[t.pin_memory() if isinstance(t, torch.Tensor) else t for t in c]) for c in scene.ring_cameras(20, args.res, args.res)[:args.views]]
    targets = [t.cpu().pin_memory() for t in targets]
elif args.streams == 0:
    targets = torch.stack(targets)
model = fit.FitModel(g, dev, fused_adam=True)
bg = torch.zeros(3, device=dev)
kw = dict(global_batch=args.views, num_streams=max(args.streams, 1), batched=args.streams == 0, host_inputs=args.e2e,
          next_cameras=cams if args.pipeline else None)
for _ in range(3):
    fit.fit_step(model, cams, targets, bg, **kw).item()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3 if args.e2e else 1):
        l = fit.fit_step(model, cams, targets, bg, **kw)
        if args.e2e:
            l.item()
    torch.cuda.synchronize()
if rank != 0:
    dist.barrier(); dist.destroy_process_group(); sys.exit(0)
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in ev)
agg = collections.OrderedDict()
for s, e, n in iv:
    a = agg.setdefault(n.split("(")[0][:60], [0, 0.0])
    a[0] += 1
    a[1] += e - s
span = iv[-1][1] - iv[0][0]
busy, cur_s, cur_e = 0.0, None, None
for s, e, _ in iv:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
total = sum(a[1] for a in agg.values())
print(f"span {span / 1e3:.3f} ms, busy(union) {busy / 1e3:.3f} ms, sum of kernel time {total / 1e3:.3f} ms, "
      f"{len(iv)} device activities, streams={args.streams}")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:60s} {c:5d} {t / 1e3:9.3f} ms {t / c:9.1f} us")
if args.e2e or "RANK" in os.environ:
    kern = [x for x in iv if not x[2].startswith("Memcpy HtoD")]
    end = kern[0][1]
    for s_, e_, n_ in kern[1:]:
        if s_ - end > 15:
            print(f"gap {s_ - end:8.1f} us before {n_[:50]} at {(s_ - iv[0][0]) / 1e3:.3f} ms")
        end = max(end, e_)
if args.timeline:
    t0 = iv[0][0]
    for s_, e_, n_ in iv:
        print(f"{(s_ - t0) / 1e3:8.3f} .. {(e_ - t0) / 1e3:8.3f} ms  {(e_ - s_):8.1f} us  {n_.split('(')[0][:70]}")
if args.out:
    json.dump([(s - iv[0][0], e - iv[0][0], n.split("(")[0][:40]) for s, e, n in iv], open(args.out, "w"))
if "RANK" in os.environ:
    t0 = iv[0][0]
    for s_, e_, n_ in iv:
        if "nccl" in n_.lower() or "adam" in n_ or "geom_backward" in n_:
            print(f"{(s_ - t0) / 1e3:8.3f} .. {(e_ - t0) / 1e3:8.3f} ms  {n_[:70]}")
    dist.barrier(); dist.destroy_process_group()
