"""Host logic of the view-sharded fit step on CPU: world_size-2 gloo run == single-process run.

The rasterizer is replaced by a small differentiable stand-in (the CUDA library cannot run
here); what is under test is dge_b200/fit.py: view sharding, the flat gradient buffer that
autograd accumulates into and the collective reduces in place, densification statistics and
the reference-style Adam path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dge_b200 import fit, scene

P, W, H, V = 257, 24, 16, 6


def fake_rasterize(rs, means3D, means2D, shs, opacities, scales, rotations):
    """Differentiable in every input, depends on the camera; returns (color, radii, depth)."""
    view = rs.viewmatrix
    pv = torch.cat([means3D, torch.ones_like(means3D[:, :1])], 1) @ view
    wgt = torch.sigmoid(pv[:, 2:3]) * opacities * scales.prod(1, keepdim=True).sqrt() * (1 + rotations[:, :1])
    base = (wgt * shs.mean(1)).sum(0) + (means2D[:, :2] * pv[:, :2]).sum()
    ramp = torch.linspace(0, 1, rs.image_height * rs.image_width).view(1, rs.image_height, rs.image_width)
    color = base.view(3, 1, 1) * ramp + rs.bg.view(3, 1, 1)
    radii = (pv[:, 2] > 3.9).to(torch.int32) * (1 + (pv[:, 0].abs() * 10).to(torch.int32))
    return color, radii, color[:1].detach()


def _data():
    g = scene.make_gaussians(P, seed=5, scale_median=0.05)
    cams = scene.ring_cameras(V, W, H)
    gen = torch.Generator().manual_seed(3)
    targets = [torch.rand(3, H, W, generator=gen) for _ in range(V)]
    return g, cams, targets


def _run_steps(rank, world, steps=3, batch=V):
    g, cams, targets = _data()
    model = fit.FitModel(g, torch.device("cpu"), fused_adam=False)
    mine = fit.shard_views(batch, rank, world)
    grad_ptr = model.flat_grad.data_ptr()
    losses = []
    for _ in range(steps):
        loss = fit.fit_step(model, [cams[i] for i in mine], [targets[i] for i in mine], torch.zeros(3), global_batch=batch,
                            rasterize=fake_rasterize, image_size=(W, H))
        losses.append(float(loss))
    assert model.flat_grad.data_ptr() == grad_ptr
    for name, p in model.params.items():  # .grad still aliases the flat buffer
        assert p.grad.data_ptr() == model.flat_grad[model.slices[name]].data_ptr(), name
    return model, losses


def _worker(rank, world, port, q, batch=V):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model, losses = _run_steps(rank, world, batch=batch)
        q.put((rank, model.flat.numpy().copy(), model.xyz_gradient_accum.numpy().copy(), model.denom.numpy().copy(),
               model.max_radii2D.numpy().copy(), losses))  # numpy: pickled by value, no shared-memory handles
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_views():
    assert fit.shard_views(20, 0, 8) == [0, 8, 16]
    assert sorted(sum((fit.shard_views(20, r, 8) for r in range(8)), [])) == list(range(20))


def test_view_sampler_draws_the_reference_batches():
    """fit.ViewSampler against the reference's own statements (threestudio/data/gs_load.py:218-222, 256-272,
    286-292) executed on Python's global generator, over several refills of the stack and a re-seed."""
    import random

    def reference_batches(total, max_view, batch, steps, seed=None):
        if seed is None:
            random.seed(0)  # gs_load.py:218
        else:
            random.seed(seed)  # gs_load.py:287
        n2n = random.sample(range(0, total), min(total, max_view))
        stack = n2n.copy()
        out = []
        for _ in range(steps):
            idx = []
            for _ in range(batch):
                if not stack:
                    stack = n2n.copy()
                v = random.choice(stack)
                stack.remove(v)
                idx.append(v)
            out.append(idx)
        return n2n, out

    for total, max_view, batch in [(37, 20, 5), (12, 48, 5), (100, 30, 7)]:
        n2n, want = reference_batches(total, max_view, batch, 11)
        s = fit.ViewSampler(total, max_view, batch)
        random.seed(12345)  # whatever else touches the global generator must not matter
        assert s.n2n_view_index == n2n
        assert [s.next_batch() for _ in range(11)] == want
        n2n, want = reference_batches(total, max_view, batch, 4, seed=3)
        s.update_cameras(3)
        assert s.n2n_view_index == n2n and [s.next_batch() for _ in range(4)] == want
    # two ranks: same batch, disjoint shares that cover it
    a, b = fit.ViewSampler(37, 20, 5), fit.ViewSampler(37, 20, 5)
    (ba, sa), (bb, sb) = a.next_shard(0, 2), b.next_shard(1, 2)
    assert ba == bb and sorted(sa + sb) == sorted(ba) and not set(sa) & set(sb)


@pytest.mark.parametrize("fused", [False])
def test_checkpoint_resume_continues_identically(fused, tmp_path):
    """state_dict / load_state_dict: a fit resumed from a checkpoint (through torch.save) takes exactly the
    steps the uninterrupted one takes — parameters, Adam moments and step, xyz rate, statistics, mask."""
    g, cams, targets = _data()
    bg = torch.zeros(3)
    mask = (torch.arange(P) % 3 != 0)

    def steps(model, n, first):
        for it in range(first, first + n):
            model.update_learning_rate(it)
            fit.fit_step(model, cams, targets, bg, global_batch=V, rasterize=fake_rasterize)

    a = fit.FitModel(g, torch.device("cpu"), fused_adam=fused)
    a.training_setup(40, spatial_lr_scale=2.0)
    a.set_grad_mask(mask)
    steps(a, 3, 0)
    torch.save(a.state_dict(), tmp_path / "fit.pt")
    steps(a, 3, 3)
    b = fit.FitModel(scene.make_gaussians(7, seed=1), torch.device("cpu"), fused_adam=fused)  # another size
    b.load_state_dict(torch.load(tmp_path / "fit.pt", weights_only=False))
    assert b.P == P and b.step_count == 3
    steps(b, 3, 3)
    assert torch.equal(a.flat, b.flat)
    for nm, _, _ in fit.GROUPS:
        for x, y in zip(a.adam_state(nm), b.adam_state(nm)):
            assert torch.equal(x, y), nm
    assert torch.equal(a.xyz_gradient_accum, b.xyz_gradient_accum) and torch.equal(a.denom, b.denom)
    assert torch.equal(a.max_radii2D, b.max_radii2D) and torch.equal(a.grad_mask, b.grad_mask)
    assert a.lrs == b.lrs and a.step_count == b.step_count == 6


def test_black_background_flag_follows_the_tensor_not_its_address():
    """fit._background_is_black: cached per tensor object AND version counter (the flag compiles the background
    term out of the blend backward; a stale 'black' would silently drop it)."""
    class M:
        pass
    m = M()
    bg = torch.zeros(3)
    assert fit._background_is_black(m, bg) == 1
    bg[1] = 0.5  # in place: same object, same address
    assert fit._background_is_black(m, bg) == 0
    other = torch.zeros(3)
    assert fit._background_is_black(m, other) == 1
    assert fit._background_is_black(m, torch.ones(3)) == 0


def test_flat_views_are_leaves():
    g, _, _ = _data()
    m = fit.FitModel(g, torch.device("cpu"), fused_adam=False)
    assert all(p.is_leaf and p.requires_grad for p in m.params.values())
    assert sum(p.numel() for p in m.params.values()) == fit.FLOATS_PER_GAUSSIAN * P
    a = m.activations()
    torch.testing.assert_close(a["scales"], g.scales, rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(a["opacities"], g.opacities, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(a["shs"], g.shs)


@pytest.mark.parametrize("batch", [V, 1])  # batch 1 on 2 ranks: rank 1's share is EMPTY and it still joins the collectives
def test_two_ranks_match_single_process(batch):
    single, losses1 = _run_steps(0, 1, batch=batch)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, batch)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, flat, accum, denom, radii, losses in res:
        flat, accum, denom, radii = (torch.from_numpy(a) for a in (flat, accum, denom, radii))
        torch.testing.assert_close(flat, single.flat, rtol=2e-5, atol=1e-6)      # replicas stay identical
        torch.testing.assert_close(accum, single.xyz_gradient_accum, rtol=1e-4, atol=1e-7)
        assert torch.equal(denom, single.denom)
        assert torch.equal(radii, single.max_radii2D)
        assert losses == pytest.approx(losses1, rel=1e-5)
    assert (res[0][1] == res[1][1]).all()  # bit-identical across ranks (same reduced gradient)
