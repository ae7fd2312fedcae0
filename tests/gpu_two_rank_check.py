"""Diagnostic for N GPUs (not a pytest file; run under torchrun): the view-sharded fit step with the
pipelined NCCL all-reduce + fused Adam leaves IDENTICAL replicas on every rank, and the reduced gradient
equals the one a single process computes over the whole batch (1e-4, float atomics)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from dge_b200 import fit, scene
from tests import util

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
P, W, H, V = 100000, 256, 256, 4 * world
g = scene.make_gaussians(P, seed=5)
cams = [scene.camera_to(c, dev) for c in scene.ring_cameras(V, W, H)]
gen = torch.Generator().manual_seed(1)
targets = [torch.rand(3, H, W, generator=gen).to(dev) for _ in range(V)]
bg = torch.zeros(3, device=dev)
mine = fit.shard_views(V, rank, world)
model = fit.FitModel(g, dev)
for _ in range(2):
    loss = fit.fit_step(model, [cams[i] for i in mine], [targets[i] for i in mine], bg, global_batch=V)
grad = model.flat_grad.clone()
flat = model.flat.clone()
# every replica identical
ref = flat.clone()
dist.broadcast(ref, 0)
same = torch.equal(ref, flat)
ok = torch.tensor([int(same)], device=dev)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print("replicas identical:", bool(ok.item()), "loss", float(loss))
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    # single-process run over the whole batch (no process group any more)
    single = fit.FitModel(g, dev)
    for _ in range(2):
        l1 = fit.fit_step(single, cams, targets, bg, global_batch=V)
    res = {}
    for name, sl in single.slices.items():
        res[name] = util.grad_ok(grad[sl].cpu().numpy(), single.flat_grad[sl].cpu().numpy())
    print("loss single", float(l1), {k: v[0] for k, v in res.items()})
    assert bool(ok.item()) and all(v[0] for v in res.values()), res
    assert abs(float(l1) - float(loss)) <= 1e-5 * abs(float(l1))
    print("two-rank check ok")
