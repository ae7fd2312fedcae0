"""Timing driver (not a pytest file): FitModel.activations_fused() of config 2's model, CUDA events over 50 calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dge_b200 import fit, scene
dev = torch.device("cuda:0")
model = fit.FitModel(scene.make_gaussians(1_000_000, seed=1236), dev)
for which in ("all", "geometry", "features"):
    for _ in range(5):
        model.activations_fused(which)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(50):
        model.activations_fused(which)
    b.record()
    torch.cuda.synchronize()
    print(f"activations_fused({which!r}): {a.elapsed_time(b) / 50 * 1e3:.1f} us per call")
acts = model.activations_fused()
ref = torch.cat([model.params["f_dc"], model.params["f_rest"]], dim=1)
print("shs == cat(f_dc, f_rest):", torch.equal(acts["shs"], ref.detach()))
