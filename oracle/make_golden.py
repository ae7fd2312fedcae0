"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference
(oracle/_ref/libref_rast.so, rebuilt for sm_100a) on a B200.

Run on the GPU box:  python oracle/make_golden.py gpurun_out/golden
then copy the files into tests/golden/ and commit them. Each fixture holds the seeded inputs,
every intermediate of the reference's forward (K1..K6), its backward outputs (K7..K9) and, for
the apply_weights cases, weights/cnt. These are the vectors that pin oracle/splat_oracle.c.
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != os.path.join(ROOT, "oracle")]
sys.path.insert(0, ROOT)
from dge_b200 import scene  # noqa: E402
from oracle import ref  # noqa: E402
from tests import util  # noqa: E402

CASES = [
    # name, P, W, H, seed, scale_median, bg, view index of 3, sh_degree
    ("c1_2k_96x80", 2048, 96, 80, 11, 0.04, (0.0, 0.0, 0.0), 0, 3),
    ("c2_3k_120x72_bg", 3000, 120, 72, 12, 0.05, (0.3, 0.1, 0.7), 1, 3),
    ("c3_1k_64x64_deg1", 1024, 64, 64, 13, 0.08, (1.0, 1.0, 1.0), 2, 1),
]


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    dev = torch.device("cuda:0")
    for name, P, W, H, seed, sm, bgv, vi, deg in CASES:
        g = scene.make_gaussians(P, seed=seed, scale_median=sm)
        if deg < 3:
            g = g._replace(shs=g.shs[:, :(deg + 1) ** 2].contiguous())
        cam = scene.ring_cameras(3, W, H)[vi]
        bg = torch.tensor(bgv, dtype=torch.float32)
        inter, state = util.ref_forward(g, cam, bg, dev, sh_degree=deg)
        dL = scene.upstream_grad(W, H, seed + 100) * 50
        grads = util.ref_backward(state, dL.to(dev))
        # apply_weights with a binary disc mask (1 channel) accumulated over this view
        mask = scene.disc_mask(W, H, radius=min(W, H) * 0.3)
        weights = torch.zeros(P, 1, device=dev)
        cnt = torch.zeros(P, 1, dtype=torch.int32, device=dev)
        a = state["args"]
        e = torch.empty(0, device=dev)
        ref.apply_weights(torch.zeros(3, device=dev), a["means3D"], weights, a["opacity"], a["scales"], a["rotations"], 1.0, e,
                          a["viewmatrix"], a["projmatrix"], a["tan_fovx"], a["tan_fovy"], H, W, e, 0, a["campos"], False,
                          mask.to(dev), cnt, False)
        torch.cuda.synchronize()
        data = {f"in_{k}": getattr(g, k).numpy() for k in g._fields}
        data.update(in_view=cam.world_view_transform.numpy(), in_proj=cam.full_proj_transform.numpy(),
                    in_campos=cam.camera_center.numpy(), in_bg=bg.numpy(), in_W=W, in_H=H, in_deg=deg,
                    in_tanfovx=math.tan(cam.FoVx * 0.5), in_tanfovy=math.tan(cam.FoVy * 0.5), in_dL=dL.numpy(),
                    in_mask=mask.numpy())
        data.update({f"fw_{k}": v for k, v in inter.items()})
        data.update({f"bw_{k}": v for k, v in grads.items()})
        data.update(aw_weights=weights.cpu().numpy(), aw_cnt=cnt.cpu().numpy())
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **data)
        print(name, "R", inter["num_rendered"], "P_v", int((inter["radii"] > 0).sum()), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
