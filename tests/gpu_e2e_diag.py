"""Diagnostic (GPU): where the end-to-end step spends its extra time over the resident one."""
import os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dge_b200 import fit, scene
import bench

dev = torch.device("cuda:0")
cfg = bench.CONFIGS["config2"]
wl = bench.Workload(cfg, dev, 20, range(20))
model = fit.FitModel(wl.g, dev, lrs=bench.LRS)
block = torch.stack(wl.targets_host).pin_memory()
restore = bench.fit_restorer(model)  # the fit moves the scene: every variant starts from the same state

def run(name, fn, n=20):
    restore()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    t0 = time.perf_counter()
    ev[0].record()
    for k in range(n):
        fn()
        ev[k + 1].record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    per = [ev[k].elapsed_time(ev[k + 1]) for k in range(n)]
    print(f"{name:60s} median {statistics.median(per):.3f} ms  min {min(per):.3f} max {max(per):.3f}  wall/step {wall:.3f}")

kw = dict(global_batch=20, image_size=(512, 512))
run("resident, no readback", lambda: fit.fit_step(model, wl.cams_dev, wl.targets_stacked, wl.bg, **kw))
run("resident, loss.item()", lambda: fit.fit_step(model, wl.cams_dev, wl.targets_stacked, wl.bg, **kw).item())
run("host cameras, resident targets, item", lambda: fit.fit_step(model, wl.cams_host, wl.targets_stacked, wl.bg, **kw).item())
run("host inputs (20 pinned targets), item", lambda: fit.fit_step(model, wl.cams_host, wl.targets_host, wl.bg, host_inputs=True, **kw).item())
run("host inputs (20 pinned targets), no readback", lambda: fit.fit_step(model, wl.cams_host, wl.targets_host, wl.bg, host_inputs=True, **kw))
run("host inputs (views of ONE pinned block), item", lambda: fit.fit_step(model, wl.cams_host, list(block.unbind(0)), wl.bg, host_inputs=True, **kw).item())
