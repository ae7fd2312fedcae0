"""Parity at BASELINE.json's FULL sizes and of the BENCHED path itself, against the unmodified reference rebuilt
for sm_100a (oracle/_ref), run live on the same inputs:

 * the per-view API at config 2's full size (1 M Gaussians, 512x512) and at a config-4-shaped size
   (1 M Gaussians, 1264x832): every intermediate, the lists and the images bit-exact, gradients within 1e-4;
 * the per-step family exactly as bench.py drives it (fit.fit_step: fused activations, batched preprocess,
   PRUNED instance lists, segmented sorts, tile partition, grid.z blends, fused L1, moment-sum backward, per-Gaussian
   backward with the activations' backward, fused Adam) against DGE's step shape around the reference rasterizer:
   images bit-exact, loss, the six raw-parameter gradient tensors + the screen-space gradient, radii max;
 * a 2-rank NCCL run of the same step (needs two GPUs): replicas bit-identical, reduced gradients equal to the
   single-process ones.
Achieved errors are printed (pytest -s / the captured output of a failure) so that a regression inside the
tolerance is visible."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

from dge_b200 import fit, scene
from tests import util
from tests.test_parity_gpu import GRAD_PAIRS, GRAD_TOL, _ref_available
from tests.test_fit_gpu import _share_fused_activations

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("P,W,H,views,seed", [(1_000_000, 512, 512, (3, 11), 1236),      # BASELINE.json configs[1], full P
                                              (1_000_000, 1264, 832, (2,), 1238)])      # configs[3] image size, 1 M
def test_per_view_api_full_size_vs_reference(cuda, P, W, H, views, seed):
    if not _ref_available():
        pytest.skip("oracle/_ref/libref_rast.so not built")
    g = scene.make_gaussians(P, seed=seed)
    bg = torch.zeros(3)
    ring = scene.ring_cameras(20, W, H)
    for ci in views:
        cam = ring[ci]
        (color, radii, depth), leaves, rs = util.ours_forward(g, cam, bg, cuda, requires_grad=True)
        mine = util.ours_intermediates(rs, g, cuda)
        refi, state = util.ref_forward(g, cam, bg, cuda)
        assert mine["num_rendered"] == refi["num_rendered"]
        assert util.compare_exact(mine, refi) == {}
        assert np.array_equal(color.detach().cpu().numpy(), refi["out_color"])
        assert np.array_equal(depth.detach().cpu().numpy(), refi["out_depth"])
        assert np.array_equal(radii.cpu().numpy(), refi["radii"])
        dL = scene.upstream_grad(W, H, seed + ci) * 50
        (color * dL.to(cuda)).sum().backward()
        rb = util.ref_backward(state, dL.to(cuda))
        for leaf, name in GRAD_PAIRS:
            ok, msg = util.grad_ok(leaves[leaf].grad.cpu().numpy(), rb[name], None, GRAD_TOL)
            print(f"{W}x{H} P={P} view {ci} R={mine['num_rendered']} {name}: {msg}")
            assert ok, (name, msg)
        del leaves, color, radii, depth, mine, refi, state, rb
        torch.cuda.empty_cache()


@pytest.mark.parametrize("P,W,H,V,bgv", [(200_000, 512, 512, 3, 0.0), (60_000, 200, 136, 4, 0.3),
                                         (1_000_000, 512, 512, 2, 0.0),      # config 2 at full P
                                         (400_000, 1264, 832, 2, 0.0)])      # two-level tile partition (4108 tiles)
def test_benched_path_vs_reference(cuda, P, W, H, V, bgv):
    """fit.fit_step's default path (what bench.py times) against DGE's step around the reference rasterizer."""
    if not _ref_available():
        pytest.skip("oracle/_ref/libref_rast.so not built")
    import bench
    g = scene.make_gaussians(P, seed=4321)
    cams = [scene.camera_to(c, cuda) for c in scene.ring_cameras(max(V, 5), W, H)[:V]]
    gen = torch.Generator().manual_seed(7)
    targets = [torch.rand(3, H, W, generator=gen).to(cuda) for _ in range(V)]
    bg = torch.zeros(3, device=cuda) + bgv
    ours, twin = fit.FitModel(g, cuda), fit.FitModel(g, cuda)
    _share_fused_activations(twin)
    refr = bench.make_reference_rasterize()
    images = []

    def recording(rs, *a):
        out = refr(rs, *a)
        images.append(out[0].detach().clone())
        return out
    l_ref = fit.fit_step(twin, cams, targets, bg, global_batch=V, direct=False, rasterize=recording)
    l = fit.fit_step(ours, cams, targets, bg, global_batch=V)  # batched, prune_lists=True, fused everything
    torch.cuda.synchronize()
    vb = ours._batches[0]
    assert len(ours._batches) == 1 and vb.V == V
    for v in range(V):
        assert torch.equal(vb.color[v], images[v]), f"image of view {v} differs from the reference's"
    assert sum(vb.num_rendered) > 0
    print(f"loss ours {float(l):.7f} reference {float(l_ref):.7f}")
    assert abs(float(l) - float(l_ref)) <= 1e-5 * abs(float(l_ref))
    for name, sl in list(ours.slices.items()) + [("means2D", ours.means2D_slice)]:
        ok, msg = util.grad_ok(ours.flat_grad[sl].cpu().numpy(), twin.flat_grad[sl].cpu().numpy(), None, GRAD_TOL)
        print(f"{W}x{H} P={P} V={V} d/d{name}: {msg}")
        assert ok, (name, msg)
    assert torch.equal(ours.max_radii2D, twin.max_radii2D)  # max over the views of the reference's radii
    assert torch.equal(ours.denom, twin.denom)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        P, W, H, V = 100_000, 256, 256, 3 * world + 1  # uneven shares
        g = scene.make_gaussians(P, seed=5)
        cams = [scene.camera_to(c, dev) for c in scene.ring_cameras(V, W, H)]
        gen = torch.Generator().manual_seed(1)
        targets = [torch.rand(3, H, W, generator=gen).to(dev) for _ in range(V)]
        bg = torch.zeros(3, device=dev)
        mine = fit.shard_views(V, rank, world)
        model = fit.FitModel(g, dev)
        for _ in range(2):
            loss = fit.fit_step(model, [cams[i] for i in mine], [targets[i] for i in mine], bg, global_batch=V)
        torch.cuda.synchronize()
        torch.save({"flat": model.flat.cpu(), "grad": model.flat_grad.cpu(), "loss": float(loss),
                    "radii": model.max_radii2D.cpu(), "accum": model.xyz_gradient_accum.cpu()}, f"{out}.{rank}")
    finally:
        dist.destroy_process_group()


def test_two_ranks_nccl_equal_single_process(cuda, tmp_path):
    """The view-sharded step over NCCL on two GPUs: both replicas end bit-identical, and the reduced gradient,
    loss and statistics equal the single-process step over the whole batch."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    out = str(tmp_path / "rank")
    mp.spawn(_nccl_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    assert torch.equal(r0["flat"], r1["flat"]) and torch.equal(r0["grad"], r1["grad"])
    assert torch.equal(r0["radii"], r1["radii"]) and r0["loss"] == r1["loss"]
    P, W, H, V = 100_000, 256, 256, 7
    g = scene.make_gaussians(P, seed=5)
    cams = [scene.camera_to(c, cuda) for c in scene.ring_cameras(V, W, H)]
    gen = torch.Generator().manual_seed(1)
    targets = [torch.rand(3, H, W, generator=gen).to(cuda) for _ in range(V)]
    single = fit.FitModel(g, cuda)
    for _ in range(2):
        l1 = fit.fit_step(single, cams, targets, torch.zeros(3, device=cuda), global_batch=V)
    torch.cuda.synchronize()
    assert abs(l1.item() - r0["loss"]) <= 1e-5 * abs(l1.item())
    for name, sl in list(single.slices.items()) + [("means2D", single.means2D_slice)]:
        ok, msg = util.grad_ok(r0["grad"][sl].numpy(), single.flat_grad[sl].cpu().numpy())
        print(f"2 ranks vs 1: d/d{name}: {msg}")
        assert ok, (name, msg)
    assert torch.equal(r0["radii"], single.max_radii2D.cpu())
