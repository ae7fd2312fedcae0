"""Timing driver (not a pytest file): fit.render_views of config 2's scene (20 views of 512^2, 1 M Gaussians) with and
without the fused semantic channel (DGE.forward's second, mask-colour render); CUDA events over 10 calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dge_b200 import fit, scene
dev = torch.device("cuda:0")
model = fit.FitModel(scene.make_gaussians(1_000_000, seed=1236), dev)
a = model.activations_fused()
cams = [scene.camera_to(c, dev) for c in scene.ring_cameras(20, 512, 512)]
bg = torch.zeros(3, device=dev)
mask = (torch.rand(model.P, device=dev) < 0.4).float()
for name, extra in (("colour only", None), ("colour + semantic channel", mask)):
    for _ in range(3):
        fit.render_views(a["means3D"], a["shs"], a["opacities"], a["scales"], a["rotations"], cams, bg, extra=extra)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        fit.render_views(a["means3D"], a["shs"], a["opacities"], a["scales"], a["rotations"], cams, bg, extra=extra)
    e1.record()
    torch.cuda.synchronize()
    print(f"render_views, {name}: {e0.elapsed_time(e1) / 10:.3f} ms per 20 views")
