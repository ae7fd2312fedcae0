"""Profiling driver (not a pytest file): fit steps of BASELINE.json config 2 through dge_b200.fit.fit_step
(direct C-ABI path, single stream so that ncu's serialised launch list reads like the real step)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dge_b200 import fit, scene

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=20)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--P", type=int, default=1_000_000)
ap.add_argument("--res", type=int, default=512)
ap.add_argument("--streams", type=int, default=0)
args = ap.parse_args()
dev = torch.device("cuda:0")
g = scene.make_gaussians(args.P, seed=1236)
cams = [scene.camera_to(c, dev) for c in scene.ring_cameras(20, args.res, args.res)[:args.views]]
gen = torch.Generator().manual_seed(3)
targets = [torch.rand(3, args.res, args.res, generator=gen).to(dev) for _ in range(args.views)]
model = fit.FitModel(g, dev, fused_adam=True)
bg = torch.zeros(3, device=dev)
for _ in range(args.steps):
    loss = fit.fit_step(model, cams, targets, bg, global_batch=args.views, num_streams=max(args.streams, 1), batched=args.streams == 0)
torch.cuda.synchronize()
print("ok", float(loss))
