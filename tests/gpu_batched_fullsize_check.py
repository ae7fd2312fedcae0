"""Diagnostic (not a pytest file): batched forward at full size (1M Gaussians, 20 views, 512^2) vs the per-view API:
images, instance lists and ranges bit-identical? Deterministic between two runs?"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dge_b200 import fit, scene, _lib as L
from dge_b200 import diff_gaussian_rasterization as dgr
dev = torch.device("cuda:0")
P, res, V = 1_000_000, 512, 20
W = H = res
lib = L.load()
g = scene.make_gaussians(P, seed=1236)
cams = [scene.camera_to(c, dev) for c in scene.ring_cameras(V, res, res)]
gen = torch.Generator().manual_seed(3)
targets = torch.stack([torch.rand(3, res, res, generator=gen).to(dev) for _ in range(V)])
bg = torch.zeros(3, device=dev)
runs = []
for rep in range(2):
    m = fit.FitModel(g, dev)
    fit.fit_step(m, cams, targets, bg, global_batch=V, batched=True, update_stats=False)
    torch.cuda.synchronize()
    vb = m._batches[0]
    runs.append((vb.color.clone(), vb.binning.clone(), vb.img.clone(), list(vb.num_rendered), m.flat_grad.clone()))
print("two batched runs: color equal", torch.equal(runs[0][0], runs[1][0]), "lists equal",
      torch.equal(runs[0][1][:4 * sum(runs[0][3])], runs[1][1][:4 * sum(runs[1][3])]),
      "grad max rel diff", float((runs[0][4] - runs[1][4]).abs().max() / runs[0][4].abs().max()))
m2 = fit.FitModel(g, dev)
a2 = m2.activations_fused()
total = 0
color_b, binning_b, img_b, nr, _ = runs[0]
istride = (lib.dge_image_bytes(W, H) + 255) // 256 * 256
T = ((W + 15) // 16) * ((H + 15) // 16)
for v, cam in enumerate(cams):
    rs = scene.raster_settings(cam, bg, 3, module=dgr)
    e = torch.empty(0, device=dev)
    R, color, depth, radii, geom, binning, img = dgr._forward_call(rs, a2["means3D"], e, a2["opacities"], a2["scales"],
                                                                   a2["rotations"], e, a2["shs"])
    bp, ip = (C.c_void_p * 2)(), (C.c_void_p * 3)()
    lib.dge_binning_pointers(binning.data_ptr(), R, W, H, bp)
    lib.dge_image_pointers(img.data_ptr(), W, H, ip)
    pl = binning[bp[0] - binning.data_ptr():][:4 * R].view(torch.int32)
    rg = img[ip[2] - img.data_ptr():][:8 * T].view(torch.int32)
    plb = binning_b[:4 * (total + R)].view(torch.int32)[total:]
    rgb_ = img_b[v * istride + (ip[2] - img.data_ptr()):][:8 * T].view(torch.int32)
    print(v, "R", R, nr[v], "color", torch.equal(color, color_b[v]), "list", torch.equal(pl, plb), "mismatches", int((pl != plb).sum()),
          "ranges", torch.equal(rg, rgb_))
    total += R
