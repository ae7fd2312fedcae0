"""Profiling driver (not a pytest file): N views of BASELINE.json config 2 (1M Gaussians, 512x512),
forward + backward through the public API, nothing else. Used under ncu; see profiles/."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dge_b200 import scene
from dge_b200 import diff_gaussian_rasterization as dgr

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=3)
ap.add_argument("--P", type=int, default=1_000_000)
ap.add_argument("--res", type=int, default=512)
args = ap.parse_args()
dev = torch.device("cuda:0")
g = scene.Gaussians(*[t.to(dev) for t in scene.make_gaussians(args.P, seed=1236)])
bg = torch.zeros(3, device=dev)
cams = [scene.camera_to(c, dev) for c in scene.ring_cameras(20, args.res, args.res)[:args.views]]
leaves = [t.clone().requires_grad_(True) for t in g]
m2d = torch.zeros(args.P, 3, device=dev, requires_grad=True)
dL = scene.upstream_grad(args.res, args.res, 3).to(dev)
for cam in cams:
    rs = scene.raster_settings(cam, bg, 3, module=dgr)
    color, radii, depth = dgr.GaussianRasterizer(rs)(means3D=leaves[0], means2D=m2d, opacities=leaves[3], shs=leaves[4],
                                                     scales=leaves[1], rotations=leaves[2])
    color.backward(dL)
torch.cuda.synchronize()
print("ok", float(color.sum()), int((radii > 0).sum()))
