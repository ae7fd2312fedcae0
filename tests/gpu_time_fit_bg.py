"""Timing driver (not a pytest file): fit.fit_step of config 2 with a NON-black background (the HAS_BG variant of the
backward blend); CUDA events over 10 steps, the scene restored before the timed region."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dge_b200 import fit, scene
from dge_b200 import _lib as L
dev = torch.device("cuda:0")
model = fit.FitModel(scene.make_gaussians(1_000_000, seed=1236), dev)
cams = [scene.camera_to(c, dev) for c in scene.ring_cameras(20, 512, 512)]
gen = torch.Generator().manual_seed(3)
targets = torch.rand(20, 3, 512, 512, generator=gen).to(dev)
lib = L.load()
for bgv in (0.0, 0.4):
    bg = torch.full((3,), bgv, device=dev)
    snap = model.flat.clone()
    for _ in range(3):
        fit.fit_step(model, cams, targets, bg, global_batch=20)
    model.flat.copy_(snap); model.exp_avg.zero_(); model.exp_avg_sq.zero_(); model.step_count = 0; model.parameters_changed()
    lib.dge_profile_enable(0xFF)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        fit.fit_step(model, cams, targets, bg, global_batch=20)
    e1.record()
    torch.cuda.synchronize()
    ms, cnt = (L.C.c_float * 8)(), (L.C.c_int * 8)()
    lib.dge_profile_read(ms, cnt)
    lib.dge_profile_enable(0)
    print(f"bg {bgv}: {e0.elapsed_time(e1) / 10:.3f} ms per step; render_bwd {ms[4] / max(cnt[4], 1):.3f} ms per launch")
    model.flat.copy_(snap); model.exp_avg.zero_(); model.exp_avg_sq.zero_(); model.step_count = 0; model.parameters_changed()
