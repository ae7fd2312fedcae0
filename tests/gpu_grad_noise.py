"""Diagnostic (not a pytest file): gradient differences reference-vs-reference (two runs: float atomics
are summed in a different order each time), ours-vs-ours, ours-vs-reference and both vs the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dge_b200 import scene
from tests import util

dev = torch.device("cuda:0")
PAIRS = [("means3D", "dL_dmeans3D"), ("means2D", "dL_dmeans2D"), ("shs", "dL_dsh"), ("opacities", "dL_dopacity"),
         ("scales", "dL_dscales"), ("rotations", "dL_drotations")]
CASES = [(200000, 1264, 832, 77, 0.02, 0.6), (300000, 512, 512, 1236, 0.012, 0.6), (200000, 512, 512, 1236, 0.012, 0.6)]
if len(sys.argv) > 1 and sys.argv[1] == "aniso":  # the needle scenes of test_anisotropic_needles_images_bit_exact
    CASES = [(12000, 256, 192, 31, 0.03, 1.6), (12000, 200, 120, 32, 0.15, 1.2), (12000, 320, 256, 33, 0.006, 2.2)]
for (P, W, H, seed, sm, ss) in CASES:
    g = scene.make_gaussians(P, seed=seed, scale_median=sm, scale_sigma=ss)
    cam = scene.ring_cameras(5, W, H)[2]
    bg = torch.zeros(3)
    dL = (scene.upstream_grad(W, H, 5) * 50).to(dev)
    refs, ours = [], []
    for rep in range(2):
        refi, state = util.ref_forward(g, cam, bg, dev)
        refs.append(util.ref_backward(state, dL))
        (color, radii, depth), leaves, rs = util.ours_forward(g, cam, bg, dev, requires_grad=True)
        (color * dL).sum().backward()
        ours.append({n: leaves[l].grad.cpu().numpy() for l, n in PAIRS})
    of = util.oracle_forward(g, cam, bg)
    ob = util.oracle_backward(of, dL.cpu(), g, cam, bg)
    print(f"P={P} {W}x{H}")
    for _, n in PAIRS:
        l2 = lambda a, b: float(np.linalg.norm((a - b).astype(np.float64)) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))
        print(f"  {n:14s} max-norm: ref/ref {util.rel_err(refs[0][n], refs[1][n]):.2e} ours/ours {util.rel_err(ours[0][n], ours[1][n]):.2e} "
              f"ours/ref {util.rel_err(ours[0][n], refs[0][n]):.2e} ref/oracle {util.rel_err(refs[0][n], ob[n]):.2e} ours/oracle {util.rel_err(ours[0][n], ob[n]):.2e}"
              f" | L2: ref/ref {l2(refs[0][n], refs[1][n]):.2e} ours/ref {l2(ours[0][n], refs[0][n]):.2e} ours/oracle {l2(ours[0][n], ob[n]):.2e} ref/oracle {l2(refs[0][n], ob[n]):.2e}")
