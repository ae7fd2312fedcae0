"""The fit step on the GPU: the direct C-ABI path (fixed buffers, fused loss, accumulate-mode
backward, several streams) against the autograd path (per-view tensors, torch ops), and the
fused Adam against torch.optim.Adam (the reference's optimiser, gaussian_model.py:374)."""
import copy

import pytest
import torch

from dge_b200 import fit, scene
from tests import util

pytestmark = pytest.mark.gpu
P, W, H, V = 30000, 160, 128, 5


def _setup(cuda, fused, P=P, W=W, H=H, V=V, scale_median=0.03):
    g = scene.make_gaussians(P, seed=61, scale_median=scale_median)
    cams = [scene.camera_to(c, cuda) for c in scene.ring_cameras(V, W, H)]
    gen = torch.Generator().manual_seed(5)
    targets = [torch.rand(3, H, W, generator=gen).to(cuda) for _ in range(V)]
    return fit.FitModel(g, cuda, fused_adam=fused), cams, targets, torch.zeros(3, device=cuda)


def _share_fused_activations(model):
    """model.activations() -> the values of the fused activation kernel (dge_fit_activate) with torch's backward
    formulas. The autograd path then sees bit-identical activated tensors to the direct path's: torch's own
    exp / sigmoid / normalize differ from the kernel's by an ulp, which the L1 loss' sign() and the blend's
    alpha >= 1/255 decisions turn into 1e-5 ... 1e-3 gradient differences that have nothing to do with the
    rasterizer (measured with tests/gpu_bg_diag.py; the activations themselves are compared in
    test_fused_activations_match_torch)."""
    class Acts(torch.autograd.Function):
        @staticmethod
        def forward(ctx, f_dc, f_rest, op, sc, rot):
            a = model.activations_fused()
            out = [a[k].clone() for k in ("shs", "opacities", "scales", "rotations")]
            ctx.save_for_backward(out[1], out[2], out[3], rot)
            return tuple(out)

        @staticmethod
        def backward(ctx, g_sh, g_op, g_sc, g_rot):
            o, s, q, raw = ctx.saved_tensors
            nrm = raw.norm(dim=1, keepdim=True).clamp_min(1e-12)
            return (g_sh[:, :1].contiguous(), g_sh[:, 1:].contiguous(), g_op * o * (1 - o), g_sc * s,
                    (g_rot - q * (q * g_rot).sum(1, keepdim=True)) / nrm)

    def acts():
        p = model.params
        shs, op, sc, rot = Acts.apply(p["f_dc"], p["f_rest"], p["opacity"], p["scaling"], p["rotation"])
        return {"means3D": p["xyz"], "shs": shs, "opacities": op, "scales": sc, "rotations": rot}
    model.activations = acts


def test_fused_activations_match_torch(cuda):
    """dge_fit_activate against GaussianModel's activations (gaussian_model.py:221-258) as torch evaluates them,
    after an optimiser step has made the raw quaternions non-unit: equal to a few ulp."""
    model, cams, targets, bg = _setup(cuda, True)
    fit.fit_step(model, cams, targets, bg, global_batch=V)
    want, got = model.activations(), model.activations_fused()
    torch.cuda.synchronize()
    for k in ("shs", "opacities", "scales", "rotations"):
        w, g = want[k].detach(), got[k]
        err = float(((g - w).abs() / w.abs().clamp_min(1e-30)).max())
        print(f"activations {k}: max relative difference {err:.2e}")
        assert err <= 5e-7, (k, err)
    assert torch.equal(got["shs"], want["shs"].detach())  # a concatenation: exact


def test_update_stats_kernel_equals_the_torch_statements(cuda):
    """dge_fit_update_stats against the statements it replaces (threestudio/systems/DGE.py:266-284,
    gaussian_model.py:811-815): max_radii2D and denom exactly, the accumulated gradient norm to an ulp; rows of
    Gaussians no view saw are untouched. P is not a multiple of the block size."""
    from dge_b200 import _lib as L
    Pn = 100_003
    gen = torch.Generator().manual_seed(21)
    radii = torch.randint(0, 40, (Pn,), generator=gen, dtype=torch.int32)
    radii = torch.where(torch.rand(Pn, generator=gen) < 0.3, torch.zeros_like(radii), radii).to(cuda)
    g2 = (torch.randn(Pn, 3, generator=gen) * 1e-3).to(cuda)
    max_r = torch.randint(0, 40, (Pn,), generator=gen, dtype=torch.int32).to(cuda)
    accum, denom = torch.rand(Pn, 1, generator=gen).to(cuda), torch.randint(0, 9, (Pn, 1), generator=gen).float().to(cuda)
    vis = radii > 0
    want_r = torch.where(vis, torch.maximum(max_r, radii), max_r)
    gn = g2[:, :2].norm(dim=-1, keepdim=True)
    want_a = accum + torch.where(vis[:, None], gn, torch.zeros_like(gn))
    want_d = denom + vis[:, None].float()
    L.check(L.load().dge_fit_update_stats(Pn, radii.data_ptr(), g2.data_ptr(), max_r.data_ptr(), accum.data_ptr(),
                                          denom.data_ptr(), L.stream_ptr(cuda)), "update stats")
    torch.cuda.synchronize()
    assert torch.equal(max_r, want_r) and torch.equal(denom, want_d)
    torch.testing.assert_close(accum, want_a, rtol=2e-7, atol=0)
    assert torch.equal(accum[~vis], want_a[~vis])


@pytest.mark.parametrize("streams,bgv,batched", [(1, 0.0, False), (3, 0.0, False), (2, 0.4, False), (1, 0.0, True),
                                                 (1, 0.4, True), (2, 0.0, True)])
def test_direct_path_matches_autograd_path(cuda, streams, bgv, batched):
    ref_model, cams, targets, bg = _setup(cuda, True)
    model, _, _, _ = _setup(cuda, True)
    _share_fused_activations(ref_model)
    bg = bg + bgv
    for step in range(2):
        l_ref = fit.fit_step(ref_model, cams, targets, bg, global_batch=V, direct=False, num_streams=1)
        g_ref = ref_model.flat_grad.clone()
        l = fit.fit_step(model, cams, targets, bg, global_batch=V, direct=True, num_streams=streams, batched=batched,
                         num_chunks=streams)
        torch.cuda.synchronize()
        assert abs(float(l) - float(l_ref)) <= 1e-5 * abs(float(l_ref))
        for name, sl in list(model.slices.items()) + [("means2D", model.means2D_slice)]:
            ok, msg = util.grad_ok(model.flat_grad[sl].cpu().numpy(), g_ref[sl].cpu().numpy())
            print(f"step {step} {name}: {msg}")
            assert ok, (step, name, msg)
        assert torch.equal(model.max_radii2D, ref_model.max_radii2D)
        assert torch.equal(model.denom, ref_model.denom)
        torch.testing.assert_close(model.xyz_gradient_accum, ref_model.xyz_gradient_accum, rtol=1e-3,
                                   atol=1e-4 * float(ref_model.xyz_gradient_accum.max()))
        # Adam normalises tiny gradients to +-lr steps: compare the reduced gradients above and
        # keep the replicas in lock-step for the next iteration
        model.flat.copy_(ref_model.flat)
        model.exp_avg.copy_(ref_model.exp_avg)
        model.exp_avg_sq.copy_(ref_model.exp_avg_sq)


@pytest.mark.parametrize("P,W,H,V,sm", [(30000, 160, 128, 5, 0.03),      # 80 tiles: 7 tile-id bits
                                        (250000, 512, 512, 3, 0.012),    # config-2 shape: 1024 tiles, 10 bits
                                        (60000, 1264, 832, 2, 0.03),     # config-4 shape: 4108 tiles (generic sort)
                                        (30001, 33, 17, 1, 0.1),         # one view, ragged image, odd P
                                        (3001, 100, 60, 64, 0.05),       # the largest batch
                                        (1, 64, 64, 2, 0.1),             # a single Gaussian
                                        (500, 17, 150, 3, 0.05),         # tall, narrow, ragged
                                        (2000, 129, 65, 5, 0.5)])        # Gaussians larger than the image
@pytest.mark.parametrize("prune", [False, True])
def test_batched_views_bit_identical_to_per_view(cuda, P, W, H, V, sm, prune):
    """dge_fit_views_forward (all views of the step per launch: batched preprocess, segmented sorts,
    single-pass tile partition, grid.z = view) must reproduce the per-view API bit for bit: images,
    depth, instance lists, tile ranges, and the max of the radii."""
    import ctypes as C
    from dge_b200 import _lib as L
    from dge_b200 import diff_gaussian_rasterization as dgr
    lib = L.load()
    model, cams, targets, bg = _setup(cuda, True, P, W, H, V, sm)
    bg = bg + 0.25
    fit.fit_step(model, cams, targets, bg, global_batch=V, batched=True, update_stats=False, prune_lists=prune)
    vb = model._batches[0]
    # the step above moved the parameters (Adam): the per-view API renders an untouched twin
    model2, _, _, _ = _setup(cuda, True, P, W, H, V, sm)
    a2 = model2.activations_fused()
    radii_max = torch.zeros(P, dtype=torch.int32, device=cuda)
    total_R = 0
    for v, cam in enumerate(cams):
        rs = scene.raster_settings(cam, bg, 3, module=dgr)
        e = torch.empty(0, device=cuda)
        R, color, depth, radii, geom, binning, img = dgr._forward_call(rs, a2["means3D"], e, a2["opacities"],
                                                                       a2["scales"], a2["rotations"], e, a2["shs"])
        assert torch.equal(color, vb.color[v]), v
        assert torch.equal(depth, vb.depth[v]), v
        if prune:
            # pruned lists (prune_lists: tiles outside the alpha >= 1/255 box are not emitted): fewer
            # instances, identical images; the lists themselves are compared with pruning off
            assert vb.num_rendered[v] <= R, v
            radii_max = torch.maximum(radii_max, radii)
            total_R += vb.num_rendered[v]
            continue
        assert R == vb.num_rendered[v], v
        # the instance list and the tile ranges of the view inside the batch's arenas
        bp, ip = (C.c_void_p * 2)(), (C.c_void_p * 3)()
        lib.dge_binning_pointers(binning.data_ptr(), R, W, H, bp)
        lib.dge_image_pointers(img.data_ptr(), W, H, ip)
        pl = binning[bp[0] - binning.data_ptr():][:4 * R].view(torch.int32)
        T = ((W + 15) // 16) * ((H + 15) // 16)
        rg = img[ip[2] - img.data_ptr():][:8 * T].view(torch.int32)
        assert torch.equal(pl, vb.binning[:4 * (total_R + R)].view(torch.int32)[total_R:]), v
        istride = (lib.dge_image_bytes(W, H) + 255) // 256 * 256
        assert torch.equal(rg, vb.img[v * istride + (ip[2] - img.data_ptr()):][:8 * T].view(torch.int32)), v
        radii_max = torch.maximum(radii_max, radii)
        total_R += R
    assert torch.equal(radii_max, vb.radii_max)


def test_batched_with_empty_views(cuda):
    """Views that see nothing (camera looking away: every Gaussian culled, num_rendered == 0) inside a
    batch, and a batch made only of such views: images are the background, gradients are exact zeros for
    those views, and the other views are unaffected."""
    g_model, cams, targets, bg = _setup(cuda, True)
    away = scene.camera_to(scene.look_at_camera((0.0, 0.0, 6.0), W, H, target=(0.0, 0.0, 12.0)), cuda)
    bg = bg + 0.5
    # mixed batch: [cam0, away, cam1]
    mixed = [cams[0], away, cams[1]]
    tg = [targets[0], targets[1], targets[2]]
    a, _, _, _ = _setup(cuda, True)
    b, _, _, _ = _setup(cuda, True)
    la = fit.fit_step(a, mixed, tg, bg, global_batch=3, batched=True, update_stats=False)
    lb = fit.fit_step(b, mixed, tg, bg, global_batch=3, batched=False, num_streams=1, update_stats=False)
    torch.cuda.synchronize()
    vb = a._batches[0]
    assert list(vb.num_rendered)[1] == 0 and list(vb.num_rendered)[0] > 0
    assert torch.equal(vb.color[1], bg.view(3, 1, 1).expand(3, H, W))
    assert abs(float(la) - float(lb)) <= 1e-5 * abs(float(lb))
    for name, sl in a.slices.items():
        ok, msg = util.grad_ok(a.flat_grad[sl].cpu().numpy(), b.flat_grad[sl].cpu().numpy())
        assert ok, (name, msg)
    # a batch in which no view sees anything
    c, _, _, _ = _setup(cuda, True)
    lc = fit.fit_step(c, [away, away], [targets[0], targets[1]], bg, global_batch=2, batched=True, update_stats=False)
    torch.cuda.synchronize()
    assert sum(c._batches[0].num_rendered) == 0
    assert not c.flat_grad.any()
    assert torch.isfinite(lc)


def test_densify_then_fit_on_gpu(cuda):
    """SURVEY.md §8f N4 on the device: fit steps accumulate the statistics, densify_and_prune rebuilds the
    flat buffers for the new Gaussian count, the next (batched) fit step runs on them; two replicas with
    generators seeded alike stay bit-identical (no broadcast needed between ranks)."""
    a, cams, targets, bg = _setup(cuda, True)
    b, _, _, _ = _setup(cuda, True)
    for m in (a, b):
        for _ in range(2):
            fit.fit_step(m, cams, targets, bg, global_batch=V, batched=True)
        assert float(m.denom.sum()) > 0
    # replicas of a multi-GPU fit hold identical state (they apply the same all-reduced gradients); two
    # independent runs on one GPU differ in the last bits (float atomics), so give b a's state
    for name in ("flat", "exp_avg", "exp_avg_sq", "xyz_gradient_accum", "denom", "max_radii2D"):
        getattr(b, name).copy_(getattr(a, name))
    counts = []
    for m in (a, b):
        gen = torch.Generator(device=cuda).manual_seed(3)
        # thresholds picked so that this small scene clones, splits and prunes something
        counts.append(m.densify_and_prune(1e-7, 1.0, 0.05, 4.0, 0, generator=gen))
    assert counts[0] == counts[1]
    P0, P1, P2, P3 = counts[0]
    assert P1 > P0 and P2 > P1 and P3 < P2 and a.P == P3
    assert torch.equal(a.flat, b.flat) and torch.equal(a.exp_avg_sq, b.exp_avg_sq)
    la = fit.fit_step(a, cams, targets, bg, global_batch=V, batched=True)
    lb = fit.fit_step(b, cams, targets, bg, global_batch=V, batched=True)
    torch.cuda.synchronize()
    assert torch.isfinite(la) and abs(float(la) - float(lb)) <= 1e-5 * abs(float(la))
    assert a._batches[0].radii_max.shape[0] == P3


def test_host_inputs_equal_resident(cuda):
    a, cams, targets, bg = _setup(cuda, True)
    b, _, _, _ = _setup(cuda, True)
    cams_h = [scene.Camera(*[t.cpu().pin_memory() if isinstance(t, torch.Tensor) else t for t in c]) for c in cams]
    targets_h = [t.cpu().pin_memory() for t in targets]
    for batched in (True, False):
        la = fit.fit_step(a, cams, targets, bg, global_batch=V, num_streams=2, batched=batched)
        lb = fit.fit_step(b, cams_h, targets_h, bg, global_batch=V, num_streams=2, host_inputs=True, batched=batched)
        torch.cuda.synchronize()
        assert abs(float(la) - float(lb)) <= 1e-6 * abs(float(la))
        for name, sl in a.slices.items():
            ok, msg = util.grad_ok(b.flat_grad[sl].cpu().numpy(), a.flat_grad[sl].cpu().numpy())
            assert ok, (name, msg)
        b.flat.copy_(a.flat), b.exp_avg.copy_(a.exp_avg), b.exp_avg_sq.copy_(a.exp_avg_sq)


def test_fused_adam_matches_torch_adam(cuda):
    fused, cams, targets, bg = _setup(cuda, True)
    stock, _, _, _ = _setup(cuda, False)
    mask = (torch.arange(P, device=cuda) % 3 != 0)
    fused.set_grad_mask(mask)
    stock.set_grad_mask(mask)
    gen = torch.Generator(device=cuda).manual_seed(1)
    for step in range(3):
        g = torch.randn(fused.flat_grad.shape, device=cuda, generator=gen) * 1e-3
        fused.flat_grad.copy_(g)
        stock.flat_grad.copy_(g)
        fused.adam_step()
        stock.adam_step()
    torch.cuda.synchronize()
    torch.testing.assert_close(fused.flat, stock.flat, rtol=2e-5, atol=2e-7)


def test_render_views_with_semantic_channel(cuda):
    """fit.render_views: all views per launch, and the edit-mask "semantic" image as a fourth blended
    channel of the SAME launch — bit-identical to the reference's way, a second per-view render with
    override_color = mask repeated three times (threestudio/systems/DGE.py:198-204)."""
    from dge_b200 import diff_gaussian_rasterization as dgr
    model, cams, _, bg = _setup(cuda, True, 40000, 200, 136, 4, 0.03)
    bg = bg + torch.tensor([0.1, 0.3, 0.2], device=cuda)
    a = {k: v.detach() for k, v in model.activations().items()}
    Pn = a["means3D"].shape[0]
    mask = (torch.rand(Pn, generator=torch.Generator().manual_seed(4)) < 0.4).to(cuda)
    color, depth, sem, radii_max = fit.render_views(a["means3D"], a["shs"], a["opacities"], a["scales"], a["rotations"],
                                                    cams, bg, extra=mask.float())
    rmax = torch.zeros(Pn, dtype=torch.int32, device=cuda)
    for v, cam in enumerate(cams):
        rs = scene.raster_settings(cam, bg, 3, module=dgr)
        zero2d = torch.zeros_like(a["means3D"])
        c1, r1, d1 = dgr.GaussianRasterizer(rs)(means3D=a["means3D"], means2D=zero2d, shs=a["shs"], colors_precomp=None,
                                                opacities=a["opacities"], scales=a["scales"], rotations=a["rotations"],
                                                cov3D_precomp=None)
        c2, _, _ = dgr.GaussianRasterizer(rs)(means3D=a["means3D"], means2D=zero2d, shs=None,
                                              colors_precomp=mask[..., None].float().repeat(1, 3),
                                              opacities=a["opacities"], scales=a["scales"], rotations=a["rotations"],
                                              cov3D_precomp=None)
        assert torch.equal(c1, color[v]) and torch.equal(d1, depth[v]), v
        assert torch.equal(c2, sem[v]), v
        rmax = torch.maximum(rmax, r1)
    assert torch.equal(rmax, radii_max)
    assert 0 < int((torch.norm(sem, dim=1) > 0.8).sum()) < sem[:, 0].numel()


@pytest.mark.parametrize("CH", [1, 3])
def test_batched_mask_backprojection_equals_per_view(cuda, CH):
    """fit.backproject_masks (dge_fit_views_apply_weights: all views per launch) against the per-view
    GaussianRasterizer.apply_weights loop of DGE.update_mask: binary masks, so weights and cnt are exact."""
    from dge_b200 import diff_gaussian_rasterization as dgr
    model, cams, _, bg = _setup(cuda, True, 40000, 200, 136, 6, 0.03)
    a = {k: v.detach() for k, v in model.activations().items()}
    Pn = a["means3D"].shape[0]
    gen = torch.Generator().manual_seed(9)
    masks = [(torch.rand(CH, 136, 200, generator=gen) > 0.6).float().to(cuda) for _ in cams]
    w1 = torch.zeros(Pn, CH, device=cuda)
    c1 = torch.zeros(Pn, 1, dtype=torch.int32, device=cuda)
    for cam, m in zip(cams, masks):
        rs = scene.raster_settings(cam, bg, 0, module=dgr)
        dgr.GaussianRasterizer(rs).apply_weights(a["means3D"], None, a["opacities"], None, w1, a["scales"], a["rotations"],
                                                 None, c1, m)
    w2 = torch.zeros(Pn, CH, device=cuda)
    c2 = torch.zeros(Pn, 1, dtype=torch.int32, device=cuda)
    counts = fit.backproject_masks(a["means3D"], a["opacities"], a["scales"], a["rotations"], cams, masks, w2, c2)
    torch.cuda.synchronize()
    assert len(counts) == len(cams) and min(counts) > 0
    assert int(c1.sum()) > 0
    assert torch.equal(c1, c2)
    assert torch.equal(w1, w2)


def test_local_edit_flow_mask_backprojection_then_masked_fit(cuda):
    """BASELINE.json config 3 at small scale (DGE.update_mask, DGE.py:101-165 + masked fit):
    apply_weights over the views with a binary mask -> weights/cnt -> selection at mask_thres ->
    grad mask; after a fit step only selected Gaussians move (rotation is not masked,
    gaussian_model.py:848)."""
    from dge_b200 import diff_gaussian_rasterization as dgr
    model, cams, targets, bg = _setup(cuda, True)
    a = {k: v.detach() for k, v in model.activations().items()}
    weights = torch.zeros(P, 1, device=cuda)
    cnt = torch.zeros(P, 1, dtype=torch.int32, device=cuda)
    mask_img = scene.disc_mask(W, H, radius=30).to(cuda)
    for cam in cams:
        rs = scene.raster_settings(cam, bg, 0, module=dgr)
        dgr.GaussianRasterizer(rs).apply_weights(a["means3D"], None, a["opacities"], None, weights, a["scales"],
                                                 a["rotations"], None, cnt, mask_img)
    sel = ((weights / (cnt + 1e-7)) > 0.8).view(-1)          # DGE.py:149-152, dge.yaml mask_thres
    assert 0 < int(sel.sum()) < P
    model.set_grad_mask(sel)
    before = model.flat.clone()
    fit.fit_step(model, cams, targets, bg, global_batch=V, num_streams=2)
    torch.cuda.synchronize()
    moved = (model.flat != before)
    for name in fit.MASKED_GROUPS:
        sl = model.slices[name]
        k = (sl.stop - sl.start) // P
        m = moved[sl].view(P, k).any(1)
        assert not m[~sel].any(), name
        assert m[sel].any(), name
    sl = model.slices["rotation"]
    assert moved[sl].view(P, 4).any(1)[~sel].any()              # rotation is not masked in the reference


def test_per_gaussian_backward_over_ranges_is_bit_identical(cuda):
    """dge_fit_backward_geom_raw over consecutive Gaussian ranges (offset pointers; what fit_step does on
    several GPUs so that a finished range's all-reduce overlaps the next range's kernel) writes exactly the
    rows one launch over all Gaussians writes."""
    from dge_b200 import _lib as L
    lib = L.load()
    Pn, Vn, Wn, Hn = 5000, 6, 128, 96
    model = fit.FitModel(scene.make_gaussians(Pn, seed=9), cuda)
    a = model.activations_fused()
    cams = torch.stack([fit.camera_record(scene.camera_to(c, cuda)) for c in scene.ring_cameras(Vn, Wn, Hn)]).to(cuda)
    gen = torch.Generator().manual_seed(4)
    acc = (torch.randn(Vn, Pn, 12, generator=gen) * 1e-3).to(cuda)
    flags = torch.randint(0, 16, (Vn, Pn), generator=gen, dtype=torch.int32).to(cuda)
    flags = torch.where(torch.rand(Vn, Pn, generator=gen).to(cuda) < 0.5, flags | 1, torch.zeros_like(flags))
    flags = flags.to(torch.uint8).contiguous()  # bit 0 visible, bits 1-3 SH clamp mask (include/dge_b200.h)
    st = L.stream_ptr(cuda)

    def run(cuts, fill):
        out = {k: torch.full_like(v.grad, fill) for k, v in model.params.items()}
        m2d = torch.full((Pn, 3), fill, device=cuda)
        for first, stop in zip(cuts[:-1], cuts[1:]):
            f = 4 * first
            L.check(lib.dge_fit_backward_geom_raw(
                stop - first, 3, Vn, cams.data_ptr(), Wn, Hn, 1.0, acc.data_ptr() + 12 * f, Pn * 12,
                flags.data_ptr() + first, Pn, a["means3D"].data_ptr() + 3 * f, a["shs"].data_ptr() + 48 * f, a["opacities"].data_ptr() + f,
                a["scales"].data_ptr() + 3 * f, a["rotations"].data_ptr() + 4 * f,
                model.params["rotation"].data_ptr() + 4 * f, out["xyz"].data_ptr() + 3 * f, m2d.data_ptr() + 3 * f,
                out["f_dc"].data_ptr() + 3 * f, out["f_rest"].data_ptr() + 45 * f, out["opacity"].data_ptr() + f,
                out["scaling"].data_ptr() + 3 * f, out["rotation"].data_ptr() + 4 * f, st), "geom raw")
        torch.cuda.synchronize()
        return {**out, "m2d": m2d}

    whole = run([0, Pn], 0.0)
    parts = run([0, 1280, 1408, 3840, Pn], 7.0)   # every row is written: the fill value never survives
    for k in whole:
        assert torch.equal(whole[k].view(torch.int32), parts[k].view(torch.int32)), k


def test_dge_fit_adapter_training_step(cuda):
    """fit.DGEFitAdapter: one DGE.training_step-shaped iteration (threestudio/systems/DGE.py:170-296, 617-699) —
    a batch of views with a local edit mask, a TORCH loss on top of the returned images (L1 plus a smooth
    term standing in for LPIPS), loss.backward(), on_before_optimizer_step — through the per-step family, against
    the same iteration through the per-view API the way DGE.forward drives it (render() twice per view: SH colours,
    then override_color = mask repeated three times)."""
    from dge_b200 import diff_gaussian_rasterization as dgr
    Pn, Wn, Hn, Vn = 40000, 200, 136, 20
    g = scene.make_gaussians(Pn, seed=17, scale_median=0.03)
    cams = [scene.camera_to(c, cuda) for c in scene.ring_cameras(Vn, Wn, Hn)]
    gen = torch.Generator().manual_seed(5)
    targets = torch.rand(Vn, Hn, Wn, 3, generator=gen).to(cuda)
    bg = torch.zeros(3, device=cuda)
    mask = (torch.rand(Pn, generator=gen) < 0.4).to(cuda)           # gaussian.mask of a local edit

    def loss_fn(images):  # [V,H,W,3]; any torch loss: L1 + a smooth second term
        return 10.0 * (images - targets).abs().mean() + 3.0 * ((images - targets) ** 2).mean()

    # --- the per-step family behind the adapter
    model = fit.FitModel(g, cuda)
    ad = fit.DGEFitAdapter(model)
    out = ad.forward(cams, bg, mask=mask)
    loss = loss_fn(out["comp_rgb"])
    loss.backward()
    ad.on_before_optimizer_step()
    torch.cuda.synchronize()

    # --- DGE.forward's loop over render() on the per-view API (twin model, same fused activations)
    twin = fit.FitModel(g, cuda)
    _share_fused_activations(twin)
    a = twin.activations()
    images, depths, sems, vsp, radii = [], [], [], [], None
    for cam in cams:
        rs = scene.raster_settings(cam, bg, 3, module=dgr)
        pts = torch.zeros_like(a["means3D"], requires_grad=True)
        img, r, d = dgr.GaussianRasterizer(rs)(means3D=a["means3D"], means2D=pts, shs=a["shs"], colors_precomp=None,
                                               opacities=a["opacities"], scales=a["scales"], rotations=a["rotations"],
                                               cov3D_precomp=None)
        sem, _, _ = dgr.GaussianRasterizer(rs)(means3D=a["means3D"], means2D=torch.zeros_like(pts), shs=None,
                                               colors_precomp=mask[..., None].float().repeat(1, 3),
                                               opacities=a["opacities"], scales=a["scales"], rotations=a["rotations"],
                                               cov3D_precomp=None)
        images.append(img.permute(1, 2, 0)); depths.append(d.permute(1, 2, 0)); sems.append(sem.detach()); vsp.append(pts)
        radii = r if radii is None else torch.max(r, radii)
    ref_images = torch.stack(images, 0)
    twin.zero_grad()
    loss_ref = loss_fn(ref_images)
    loss_ref.backward()
    torch.cuda.synchronize()
    # forward products: bit-identical
    assert torch.equal(out["comp_rgb"].detach(), ref_images.detach())
    assert torch.equal(out["depth"], torch.stack(depths, 0))
    assert torch.equal(out["semantic_render"], torch.stack(sems, 0))
    assert torch.equal(out["masks"], torch.norm(torch.stack(sems, 0), dim=1) > 0.8)
    assert torch.equal(out["radii"], radii) and torch.equal(out["visibility_filter"], radii > 0)
    assert abs(float(loss.detach()) - float(loss_ref.detach())) <= 1e-6 * abs(float(loss_ref.detach()))
    # gradients: raw parameters + the summed screen-space gradient (DGE.py:269-276)
    vs_grad = sum(v.grad for v in vsp)
    for name, sl in model.slices.items():
        ok, msg = util.grad_ok(model.flat_grad[sl].cpu().numpy(), twin.flat_grad[sl].cpu().numpy())
        print(f"adapter d/d{name}: {msg}")
        assert ok, (name, msg)
    ok, msg = util.grad_ok(model.means2D.grad.cpu().numpy(), vs_grad.cpu().numpy())
    assert ok, ("viewspace", msg)
    # statistics of on_before_optimizer_step (gaussian_model.py:811-815)
    vis = radii > 0
    assert torch.equal(model.denom.view(-1) > 0, vis)
    want = torch.where(vis, vs_grad[:, :2].norm(dim=-1), torch.zeros(Pn, device=cuda))
    torch.testing.assert_close(model.xyz_gradient_accum.view(-1), want, rtol=1e-3, atol=1e-4 * float(want.max()))
    assert torch.equal(model.max_radii2D, torch.where(vis, radii, torch.zeros_like(radii)))
    before = model.flat.clone()
    ad.optimizer_step()
    assert not torch.equal(before, model.flat) and model.step_count == 1


@pytest.mark.parametrize("chunks", [1, 2])
def test_prefetched_front_half_gives_the_same_step(cuda, chunks):
    """fit_step(next_cameras=...): the next step's projection / depth sort / binning launched at the end of a step
    (geometry only, before the features are stepped), completed with dge_fit_views_colour at the start of the
    next — the step it produces must be the unpipelined one: images bit-identical, loss and gradients equal."""
    a, cams, targets, bg = _setup(cuda, True, 60000, 200, 136, 6, 0.03)
    b, _, _, _ = _setup(cuda, True, 60000, 200, 136, 6, 0.03)
    other = [scene.camera_to(c, cuda) for c in scene.ring_cameras(9, 200, 136)[3:9]]
    fit.fit_step(a, cams, targets, bg, global_batch=6, num_chunks=chunks, next_cameras=other)
    assert a._front is not None and not a._front["coloured"]           # geometry-only front half of `other`
    for name in ("flat", "exp_avg", "exp_avg_sq", "xyz_gradient_accum", "denom", "max_radii2D"):
        getattr(b, name).copy_(getattr(a, name))
    b.step_count = a.step_count
    la = fit.fit_step(a, other, targets, bg, global_batch=6, num_chunks=chunks)   # finds the prefetched front half
    lb = fit.fit_step(b, other, targets, bg, global_batch=6, num_chunks=chunks)
    torch.cuda.synchronize()
    for va, vb in zip(a._batches, b._batches):
        assert torch.equal(va.color, vb.color) and torch.equal(va.depth, vb.depth)
        assert list(va.num_rendered) == list(vb.num_rendered)
        assert torch.equal(va.flags, vb.flags)
    assert abs(float(la) - float(lb)) <= 1e-6 * abs(float(lb))
    for name, sl in list(a.slices.items()) + [("means2D", a.means2D_slice)]:
        ok, msg = util.grad_ok(a.flat_grad[sl].cpu().numpy(), b.flat_grad[sl].cpu().numpy())
        assert ok, (name, msg)
    assert torch.equal(a.max_radii2D, b.max_radii2D) and torch.equal(a.denom, b.denom)
    # a front half prefetched for other cameras than the next call's is not used
    fit.fit_step(a, other, targets, bg, global_batch=6, num_chunks=chunks, next_cameras=cams)
    assert not a._front["coloured"]
    lc = fit.fit_step(a, other, targets, bg, global_batch=6, num_chunks=chunks)   # `cams` was prefetched
    assert torch.isfinite(lc) and a._front["coloured"]


def _state_for_densify(cuda, seed_mask=True):
    model, cams, targets, bg = _setup(cuda, True, 40000, 200, 136, 4, 0.05)
    if seed_mask:
        model.set_grad_mask(torch.rand(model.P, generator=torch.Generator().manual_seed(8)) < 0.7)
    for _ in range(3):
        fit.fit_step(model, cams, targets, bg, global_batch=4)
    return model


@pytest.mark.parametrize("masked,max_screen", [(True, 5), (False, 5), (True, 0)])
def test_device_densify_equals_torch_path(cuda, masked, max_screen):
    """csrc/densify.cu (dge_densify_select / dge_densify_gather: decisions, compaction, clone / split appends and
    Adam-state surgery as two kernels over the flat buffers) against the torch path of FitModel.densify_and_prune,
    which tests/test_densify.py pins bit-exactly to the reference's own code: same counts, same rows in the same
    order — parameters, both Adam moments, edit mask identical; only the split children's positions (a 3x3
    product whose accumulation order torch leaves to cuBLAS) are compared to an ulp."""
    a = _state_for_densify(cuda, masked)
    b = _state_for_densify(cuda, masked)
    for name in ("flat", "exp_avg", "exp_avg_sq", "xyz_gradient_accum", "denom", "max_radii2D"):
        getattr(b, name).copy_(getattr(a, name))
    P0 = a.P
    scale_med = float(torch.exp(a.params["scaling"].detach()).max(dim=1).values.median())
    extent = scale_med / 0.01          # percent_dense * extent = the median size: clones AND splits occur
    gmed = float((a.xyz_gradient_accum / a.denom.clamp_min(1)).median())
    args = (gmed, 0.5, 0.3, extent, max_screen)
    ca = a.densify_and_prune(*args, generator=torch.Generator(device=cuda).manual_seed(3), device_kernels=True)
    cb = b.densify_and_prune(*args, generator=torch.Generator(device=cuda).manual_seed(3), device_kernels=False)
    torch.cuda.synchronize()
    print("densify counts (before, after clone, after split, after prune):", ca)
    assert ca == cb and ca[1] > ca[0] and ca[2] > ca[1] and ca[3] < ca[2] and a.P == b.P == ca[3] != P0
    n_child_rows = None
    for name in a.params:
        pa, pb = a.params[name].detach(), b.params[name].detach()
        if name == "xyz":
            same = (pa == pb).all(dim=1)
            n_child_rows = int((~same).sum())
            torch.testing.assert_close(pa, pb, rtol=2e-6, atol=1e-7)
            first_diff = int(torch.nonzero(~same)[0]) if n_child_rows else a.P
            assert same[:first_diff].all()   # originals and clones are copies: exact; differences only among the children
        else:
            assert torch.equal(pa, pb), name
        for ma, mb in zip(a.adam_state(name), b.adam_state(name)):
            assert torch.equal(ma, mb), name
    print(f"children whose position differs in the last bit from torch's bmm: {n_child_rows} of {ca[3]}")
    assert (a.grad_mask is None) == (b.grad_mask is None)
    if masked:
        assert torch.equal(a.grad_mask, b.grad_mask)
    assert not a.xyz_gradient_accum.any() and not a.denom.any() and not a.max_radii2D.any()
    # the rebuilt model keeps fitting
    _, cams, targets, bg = _setup(cuda, True, 40000, 200, 136, 4, 0.05)
    la = fit.fit_step(a, cams, targets, bg, global_batch=4)
    assert torch.isfinite(la)


def test_densify_fixtures_on_device(cuda):
    """The reference's own densification (fixtures of oracle/make_densify_golden.py, produced by its code on the
    CPU) through the device kernels: counts, parameters, Adam moments and mask as in the fixture (positions and
    log-scales of the split children to an ulp: CPU vs GPU libm)."""
    import glob, os
    import numpy as np
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "densify_*.npz")))
    assert files
    names = [g[0] for g in fit.GROUPS]
    for path in files:
        z = np.load(path)
        max_grad, pct, min_opacity, extent, max_screen, percent_dense = (float(x) for x in z["hyper"])
        P = z["in_xyz"].shape[0]
        model = fit.FitModel(scene.make_gaussians(P, seed=1), cuda)
        model.step_count = int(z["adam_step"][0])
        model._allocate({n: torch.from_numpy(z["in_" + n]) for n in names}, {n: torch.from_numpy(z["in_m_" + n]) for n in names},
                        {n: torch.from_numpy(z["in_v_" + n]) for n in names})
        model.xyz_gradient_accum = torch.from_numpy(z["in_xyz_gradient_accum"]).to(cuda)
        model.denom = torch.from_numpy(z["in_denom"]).to(cuda)
        model.max_radii2D = torch.from_numpy(z["in_max_radii2D"]).to(torch.int32).to(cuda)
        model.set_grad_mask(torch.from_numpy(z["in_mask"]))
        counts = model.densify_and_prune(max_grad, pct, min_opacity, extent, int(max_screen), percent_dense=percent_dense,
                                         normal_samples=torch.from_numpy(z["normal_samples"]), device_kernels=True)
        assert list(counts) == [int(c) for c in z["counts"]], path
        for n in names:
            got = model.params[n].detach().cpu().numpy()
            if n in ("xyz", "scaling"):
                np.testing.assert_allclose(got, z["out_" + n], rtol=3e-6, atol=1e-6, err_msg=n)
            else:
                assert np.array_equal(got, z["out_" + n]), (path, n)
            m, v = model.adam_state(n)
            assert np.array_equal(m.cpu().numpy(), z["out_m_" + n]) and np.array_equal(v.cpu().numpy(), z["out_v_" + n]), n
        assert np.array_equal(model.grad_mask.bool().cpu().numpy(), z["out_mask"])
