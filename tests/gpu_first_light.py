"""Manual GPU bring-up script (not a pytest file): product vs reference vs oracle, verbose."""
import sys, os, time, math, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dge_b200 import scene
from tests import util

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), "cpu cores", os.cpu_count(), flush=True)


def run(P, W, H, seed, V=2, scale_median=0.012, bgv=(0.0, 0.0, 0.0), check_oracle=True, grads=True):
    print(f"==== P={P} {W}x{H} seed={seed} scale={scale_median}", flush=True)
    g = scene.make_gaussians(P, seed=seed, scale_median=scale_median)
    cams = scene.ring_cameras(V, W, H)
    bg = torch.tensor(bgv, dtype=torch.float32)
    for ci, cam in enumerate(cams):
        try:
            (color, radii, depth), leaves, rs = util.ours_forward(g, cam, bg, dev, requires_grad=grads)
            torch.cuda.synchronize()
            mine = util.ours_intermediates(rs, g, dev)
            refi, state = util.ref_forward(g, cam, bg, dev)
            Pv = int((refi["radii"] > 0).sum())
            cnt = refi["ranges"][:, 1].astype(np.int64) - refi["ranges"][:, 0]
            print(f" cam{ci}: P_v={Pv} R={refi['num_rendered']} (ours {mine['num_rendered']}) R/Pv={refi['num_rendered']/max(Pv,1):.2f} max/tile={cnt.max()}")
            bad = util.compare_exact(mine, refi)
            print("  bit-exact vs reference:", "ALL OK" if not bad else bad)
            for n in ("out_color", "out_depth", "final_T"):
                print(f"   {n} maxabs {np.abs(mine[n].astype(np.float64)-refi[n]).max():.3e}", end="")
            print()
            if check_oracle:
                of = util.oracle_forward(g, cam, bg)
                bo = util.compare_exact(of, refi, names=("radii", "tiles_touched", "means2D", "depths", "conic_opacity", "rgb", "clamped", "keys", "point_list", "ranges", "n_contrib", "cov3D"))
                print("  oracle bit-exact vs reference:", "ALL OK" if not bo else bo,
                      " color maxabs", np.abs(of["out_color"] - refi["out_color"]).max(), "finalT", np.abs(of["final_T"] - refi["final_T"]).max())
            if grads:
                dL = scene.upstream_grad(W, H, seed + 1 + ci) * 50
                dLd = dL.to(dev)
                (color * dLd).sum().backward()
                torch.cuda.synchronize()
                rb = util.ref_backward(state, dLd)
                pairs = [("means3D", "dL_dmeans3D"), ("means2D", "dL_dmeans2D"), ("shs", "dL_dsh"), ("opacities", "dL_dopacity"),
                         ("scales", "dL_dscales"), ("rotations", "dL_drotations")]
                msg = []
                for leaf, name in pairs:
                    msg.append(f"{name}:{util.rel_err(leaves[leaf].grad.cpu().numpy(), rb[name]):.2e}")
                print("  grads rel vs reference:", " ".join(msg))
                if check_oracle:
                    ob = util.oracle_backward(of, dL, g, cam, bg)
                    msg = [f"{name}:{util.rel_err(rb[name], ob[name]):.2e}" for _, name in pairs]
                    print("  reference grads rel vs oracle:", " ".join(msg))
        except Exception:
            traceback.print_exc()
            torch.cuda.synchronize()


if __name__ == "__main__":
    run(16384, 256, 256, 1235)
    run(5000, 200, 120, 7, scale_median=0.05, bgv=(0.3, 0.1, 0.7))
    run(300000, 512, 512, 1236, V=2, check_oracle=True)
    run(1000000, 512, 512, 1236, V=2, check_oracle=False)
    # timing, config 2 shape
    g = scene.make_gaussians(1000000, seed=1236)
    cam = scene.ring_cameras(4, 512, 512)[1]
    bg = torch.zeros(3)
    for name in ("ours", "ref"):
        ts = []
        for it in range(6):
            torch.cuda.synchronize(); t0 = time.time()
            if name == "ours":
                (color, radii, depth), leaves, rs = util.ours_forward(g, cam, bg, dev, requires_grad=True)
                torch.cuda.synchronize(); t1 = time.time()
                color.sum().backward()
            else:
                refi_state = None
                from oracle import ref
                inter_state = util.ref_forward.__wrapped__ if hasattr(util.ref_forward, "__wrapped__") else None
                gd = util.to_dev(g, dev); cam_d = scene.camera_to(cam, dev); tfx, tfy = util.tans(cam)
                e = torch.empty(0, device=dev)
                torch.cuda.synchronize(); t0 = time.time()
                R, c, d, r, ge, bi, im = ref.rasterize_gaussians(bg.to(dev), gd.means3D, e, gd.opacities, gd.scales, gd.rotations, 1.0, e,
                    cam_d.world_view_transform, cam_d.full_proj_transform, tfx, tfy, 512, 512, gd.shs, 3, cam_d.camera_center, False, False)
                torch.cuda.synchronize(); t1 = time.time()
                ref.rasterize_gaussians_backward(bg.to(dev), gd.means3D, r, e, gd.scales, gd.rotations, 1.0, e, cam_d.world_view_transform,
                    cam_d.full_proj_transform, tfx, tfy, torch.ones(3, 512, 512, device=dev), gd.shs, 3, cam_d.camera_center, ge, R, bi, im, False)
            torch.cuda.synchronize(); t2 = time.time()
            ts.append((t1 - t0, t2 - t1))
        print(name, "fwd/bwd ms:", [(round(a * 1e3, 2), round(b * 1e3, 2)) for a, b in ts], flush=True)
