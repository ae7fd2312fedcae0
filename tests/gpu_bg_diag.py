"""Diagnostic (GPU): second fit step with a non-black background — autograd path vs direct paths vs the reference."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dge_b200 import fit, scene
from tests import util
import bench

cuda = torch.device("cuda:0")
P, W, H, V = 30000, 160, 128, 5
def setup(lrs=None):
    g = scene.make_gaussians(P, seed=61, scale_median=0.03)
    cams = [scene.camera_to(c, cuda) for c in scene.ring_cameras(V, W, H)]
    gen = torch.Generator().manual_seed(5)
    targets = [torch.rand(3, H, W, generator=gen).to(cuda) for _ in range(V)]
    return fit.FitModel(g, cuda, fused_adam=True, lrs=lrs), cams, targets
refr = bench.make_reference_rasterize()
for lrs in (None, dict(fit.DEFAULT_LRS, f_dc=0.0025, f_rest=0.0025 / 20)):
    for bgv in (0.0, 0.4):
        a, cams, targets = setup(lrs)
        b, _, _ = setup(lrs); c, _, _ = setup(lrs); d, _, _ = setup(lrs)
        bg = torch.zeros(3, device=cuda) + bgv
        for step in range(2):
            fit.fit_step(a, cams, targets, bg, global_batch=V, direct=False)
            ga = a.flat_grad.clone()
            fit.fit_step(b, cams, targets, bg, global_batch=V, direct=True, batched=True)
            fit.fit_step(c, cams, targets, bg, global_batch=V, direct=True, batched=False)
            fit.fit_step(d, cams, targets, bg, global_batch=V, direct=False, rasterize=refr)
            torch.cuda.synchronize()
            for name, sl in list(a.slices.items()) + [("means2D", a.means2D_slice)]:
                r = d.flat_grad[sl].cpu().numpy()
                msg = [f"{n}: {util.l2_err(m.flat_grad[sl].cpu().numpy(), r):.1e}/{util.rel_err(m.flat_grad[sl].cpu().numpy(), r):.1e}" for n, m in (("autograd", a), ("batched", b), ("direct", c))]
                print(f"lrs={'new' if lrs is None else 'old'} bg={bgv} step={step} {name:9s} vs reference (L2/max): " + "  ".join(msg))
            for m in (b, c, d):
                m.flat.copy_(a.flat); m.exp_avg.copy_(a.exp_avg); m.exp_avg_sq.copy_(a.exp_avg_sq)
