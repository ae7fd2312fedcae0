#!/usr/bin/env python
"""bench.py — fwd+bwd views/s of the DGE 3D-fit step (BASELINE.json metric).

Workload (config 2 of BASELINE.json / SURVEY.md §8d): 1 M synthetic Gaussians (randgauss-v1),
SH degree 3, a batch of 20 edited views per GPU at 512x512, L1 loss to fixed random targets,
forward + backward through the rasterizer for every view, one Adam step over the 59
floats/Gaussian per step. One process per GPU; with N>1 each rank fits 20 views per step
against its replica and the gradients + densification statistics are all-reduced (weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0). `value` = views/s with the step's inputs resident in HBM;
`e2e` = the same step through the public API with cameras/targets in pinned HOST memory and
the loss read back every step. Besides the K-step total (the contract's number) every step is
bracketed by its own CUDA events and the median is reported next to it.

`--impl reference` is the reference's OWN program shape with nothing of this repository's product
in the process: DGE's training step (per-view render() with the activations of GaussianModel,
stacked L1 loss, one backward, on_before_optimizer_step's statistics, torch.optim.Adam) around
the UNMODIFIED reference rasterizer rebuilt for sm_100a (oracle/_ref/libref_rast.so; the reference
ships no CPU rasterizer). It imports dge_b200.scene (the synthetic scene, pure torch) and oracle/
only; libdge_b200.so is never loaded in that process.

`extras` (ours): the same JSON line also carries the drop-in boundary of SURVEY.md §8d (`dropin`: the
reference arm's loop with ONLY the import swapped), config 3 (mask back-projection), config 4 as a
strong-scaling point (64 views split over the N ranks) and one config 5 point, each with its own
P_visible / R.
"""
import argparse
import json
import math
import os
import statistics
import sys
import threading
import time

# Several GPUs: which transport / algorithm family NCCL sets up (NVLS on NVSwitch?) goes on the JSON line. Rank 0's one-off
# INIT lines are sent to a file — nothing is logged per collective, the timed steps are not disturbed. The variables have
# to be in the environment before the NCCL library first looks at them, hence before torch is imported.
NCCL_LOG = None
NCCL_DEBUG_WAS = os.environ.get("NCCL_DEBUG")
if (int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("RANK", "0") == "0"
        and (NCCL_DEBUG_WAS or "").upper() not in ("INFO", "TRACE") and "NCCL_DEBUG_FILE" not in os.environ):
    import tempfile
    NCCL_LOG = os.path.join(tempfile.gettempdir(), f"dge_b200_nccl_{os.getpid()}.log")
    os.environ.update(NCCL_DEBUG="INFO", NCCL_DEBUG_SUBSYS="INIT", NCCL_DEBUG_FILE=NCCL_LOG)

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: P, W, H, views per GPU per step, seed
    "config2": dict(P=1_000_000, W=512, H=512, V=20, seed=1236,
                    desc="DGE global edit fit: 1M Gaussians, 20 views/step/GPU at 512x512, fwd+bwd+Adam"),
    # BASELINE.json configs[3]: use_original_resolution, 64 views sharded over the GPUs (8 per GPU at 8 GPUs)
    "config4": dict(P=3_000_000, W=1264, H=832, V=8, seed=1238,
                    desc="use_original_resolution: 3M Gaussians at 1264x832, 8 views/step/GPU, fwd+bwd+Adam"),
    # BASELINE.json configs[4] (one point of the sweep): 1080p, sort/duplicate stress
    "config5": dict(P=2_000_000, W=1920, H=1080, V=8, seed=1239,
                    desc="densification sweep point: 2M Gaussians at 1920x1080, 8 views/step/GPU, fwd+bwd+Adam"),
    # BASELINE.json configs[2], first half: DGE.update_mask (DGE.py:101-165) — mask back-projection over 40 views
    "config3": dict(P=1_000_000, W=512, H=512, V=40, seed=1237,
                    desc="DGE local edit: mask back-projection (apply_weights) of a disc mask over 40 views of 512x512, 1M Gaussians"),
    "config1": dict(P=16_384, W=256, H=256, V=1, seed=1235, desc="16k Gaussians, one 256x256 view (parity config)"),
    "tiny": dict(P=50_000, W=256, H=256, V=4, seed=1240, desc="smoke-sized"),
    "hostbound": dict(P=2_000, W=64, H=64, V=20, seed=1241, desc="negligible GPU work: measures host overhead per view"),
}
METRIC = "fwd+bwd views/s @1M Gaussians 512^2 (DGE 3D-fit step incl. Adam)"
# the reference's learning rates (gaussiansplatting/arguments/__init__.py:72-81; configs/dge.yaml scales them by 1)
LRS = {"xyz": 0.00016, "f_dc": 0.0125, "f_rest": 0.0125 / 20.0, "opacity": 0.05, "scaling": 0.005, "rotation": 0.001}
LAMBDA_L1 = 10.0  # configs/dge.yaml:62


# --------------------------------------------------------------------------- clocks -----
class ClockSampler:
    """SM clock and throttle reasons through NVML in a background thread. On this driver every NVML
    query stalls the launching thread for ~25 ms, so the sampler runs over K IDENTICAL steps right
    after the timed regions; the clock DURING the timed regions is measured on the device itself
    (clock_probe below), which needs no driver query."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index, period=0.1):
        self.gpu, self.period, self.sm, self.bits, self.ok = gpu_index, period, [], 0, False
        self.stop_flag = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
        except Exception as ex:
            self.err = str(ex)

    def _loop(self):
        while not self.stop_flag.is_set():
            try:
                self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.bits |= self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self.stop_flag.set()
        self.t.join(timeout=2)
        sm = sorted(self.sm)
        return {"sm_mhz": float(sm[len(sm) // 2]) if sm else None, "sm_max_mhz": float(self.max),
                "reasons": sorted(n for b, n in self.REASONS.items() if self.bits & b), "samples": len(sm)}


class ClockProbe:
    """Average SM clock of a timed region, measured ON THE DEVICE: a tiny kernel stores every SM's cycle
    counter and the global nanosecond timer at the start and at the end of the region (same stream);
    MHz = delta cycles / delta ns per SM, median over the SMs both probes reached."""

    def __init__(self, launch, dev):
        self.launch = launch
        self.a = torch.zeros(256, 2, dtype=torch.int64, device=dev)
        self.b = torch.zeros(256, 2, dtype=torch.int64, device=dev)

    def begin(self):
        self.a.zero_()
        self.b.zero_()
        self.launch(self.a.data_ptr())

    def end(self):
        self.launch(self.b.data_ptr())

    def mhz(self):
        a, b = self.a.cpu(), self.b.cpu()
        ok = (a[:, 1] > 0) & (b[:, 1] > a[:, 1])
        if int(ok.sum()) == 0:
            return None
        f = (b[ok, 0] - a[ok, 0]).double() / (b[ok, 1] - a[ok, 1]).double() * 1e3
        return float(f.median())


def timed_steps(step_fn, K, barrier, probe=None):
    """K steps between barriers (the contract's bracket); every step also has its own event pair.
    Returns (total ms, [ms of each step])."""
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    barrier()
    if probe is not None:
        probe.begin()
    ev[0].record()
    for k in range(K):
        step_fn()
        ev[k + 1].record()
    if probe is not None:
        probe.end()
    barrier()
    return ev[0].elapsed_time(ev[K]), [ev[k].elapsed_time(ev[k + 1]) for k in range(K)]


def nvml_clocks_under_load(step_fn, barrier, gpu_index, ms_per_step, K, sync_count=None):
    sampler = ClockSampler(gpu_index, period=0.05)
    sampler.start()
    # at least ~0.6 s of load so that the median rests on a dozen samples
    n = max(K, min(200, int(math.ceil(600.0 / max(ms_per_step, 0.05)))))
    if sync_count is not None:
        n = sync_count(n)
    for _ in range(n):
        step_fn()
    barrier()
    return sampler.stop()


# ------------------------------------------------- the reference's program shape -----
def make_reference_rasterize():
    """autograd.Function around the reference's _C calls, mirroring
    DGR/diff_gaussian_rasterization/__init__.py:50-225 (the reference's own binding cannot be
    imported: its _C is a torch extension that needs ~5 min of nvcc per build)."""
    from oracle import ref

    class RefRasterize(torch.autograd.Function):
        @staticmethod
        def forward(ctx, means3D, means2D, sh, opacities, scales, rotations, rs):
            e = torch.empty(0, dtype=torch.float32, device=means3D.device)
            R, color, depth, radii, geom, binning, img = ref.rasterize_gaussians(
                rs.bg, means3D, e, opacities, scales, rotations, rs.scale_modifier, e, rs.viewmatrix, rs.projmatrix,
                rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos,
                rs.prefiltered, rs.debug)
            ctx.rs, ctx.R = rs, R
            RefRasterize.last_R = R
            ctx.save_for_backward(means3D, scales, rotations, radii, sh, geom, binning, img)
            return color, radii, depth

        @staticmethod
        def backward(ctx, grad_color, _gr, _gd):
            rs = ctx.rs
            means3D, scales, rotations, radii, sh, geom, binning, img = ctx.saved_tensors
            e = torch.empty(0, dtype=torch.float32, device=means3D.device)
            (g_m2d, _g_col, g_op, g_m3d, _g_cov, g_sh, g_sc, g_rot, _g_con) = ref.rasterize_gaussians_backward(
                rs.bg, means3D, radii, e, scales, rotations, rs.scale_modifier, e, rs.viewmatrix, rs.projmatrix,
                rs.tanfovx, rs.tanfovy, grad_color, sh, rs.sh_degree, rs.campos, geom, ctx.R, binning, img, rs.debug)
            return g_m3d, g_m2d, g_sh, g_op, g_sc, g_rot, None

    def rasterize(rs, means3D, means2D, shs, opacities, scales, rotations):
        return RefRasterize.apply(means3D, means2D, shs, opacities, scales, rotations, rs)

    rasterize.cls = RefRasterize
    return rasterize


class Settings:
    """GaussianRasterizationSettings as gaussian_renderer.render() fills them
    (gaussiansplatting/gaussian_renderer/__init__.py:72-88); a plain object so that the reference arm needs
    no binding module."""
    __slots__ = ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix",
                 "projmatrix", "sh_degree", "campos", "prefiltered", "debug")

    def __init__(self, cam, bg, sh_degree):
        self.image_height, self.image_width = int(cam.image_height), int(cam.image_width)
        self.tanfovx, self.tanfovy = math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5)
        self.bg, self.scale_modifier = bg, 1.0
        self.viewmatrix, self.projmatrix, self.campos = cam.world_view_transform, cam.full_proj_transform, cam.camera_center
        self.sh_degree, self.prefiltered, self.debug = sh_degree, False, False


class DGEStyleModel:
    """GaussianModel's optimisable state and optimiser as the reference holds them
    (gaussiansplatting/scene/gaussian_model.py:60-75, 336-380): six separate leaf tensors, six Adam groups."""

    def __init__(self, g, dev):
        p = lambda t: torch.nn.Parameter(t.to(dev).contiguous().requires_grad_(True))
        self._xyz = p(g.means3D)
        self._features_dc = p(g.shs[:, :1, :])
        self._features_rest = p(g.shs[:, 1:, :])
        self._opacity = p(torch.log(g.opacities / (1 - g.opacities)))
        self._scaling = p(torch.log(g.scales))
        self._rotation = p(g.rotations)
        P = g.means3D.shape[0]
        self.xyz_gradient_accum = torch.zeros(P, 1, device=dev)
        self.denom = torch.zeros(P, 1, device=dev)
        self.max_radii2D = torch.zeros(P, device=dev)
        groups = [{"params": [self._xyz], "lr": LRS["xyz"], "name": "xyz"},
                  {"params": [self._features_dc], "lr": LRS["f_dc"], "name": "f_dc"},
                  {"params": [self._features_rest], "lr": LRS["f_rest"], "name": "f_rest"},
                  {"params": [self._opacity], "lr": LRS["opacity"], "name": "opacity"},
                  {"params": [self._scaling], "lr": LRS["scaling"], "name": "scaling"},
                  {"params": [self._rotation], "lr": LRS["rotation"], "name": "rotation"}]
        self.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)


def dge_style_step(m, rasterize, make_settings, cams, targets, bg, global_batch, host, dev):
    """One DGE training step as the reference runs it: DGE.forward's camera loop over render()
    (threestudio/systems/DGE.py:170-239; gaussiansplatting/gaussian_renderer/__init__.py:45-150 with the
    activations of gaussian_model.py:221-258 evaluated for EVERY view), the stacked L1 loss (DGE.py:672), one
    backward, on_before_optimizer_step's statistics (DGE.py:266-284, gaussian_model.py:811-815), Adam."""
    images, vsp, radii_max, tg = [], [], None, []
    for cam, target in zip(cams, targets):
        if host:
            cam = cam._replace(world_view_transform=cam.world_view_transform.to(dev, non_blocking=True),
                               full_proj_transform=cam.full_proj_transform.to(dev, non_blocking=True),
                               camera_center=cam.camera_center.to(dev, non_blocking=True))
            target = target.to(dev, non_blocking=True)
        screenspace_points = torch.zeros_like(m._xyz, dtype=m._xyz.dtype, requires_grad=True, device=dev) + 0
        screenspace_points.retain_grad()
        rs = make_settings(cam, bg, 3)
        shs = torch.cat((m._features_dc, m._features_rest), dim=1)
        color, radii, _depth = rasterize(rs, m._xyz, screenspace_points, shs, torch.sigmoid(m._opacity),
                                         torch.exp(m._scaling), torch.nn.functional.normalize(m._rotation))
        vsp.append(screenspace_points)
        radii_max = radii if radii_max is None else torch.max(radii, radii_max)
        images.append(color.permute(1, 2, 0))
        tg.append(target.permute(1, 2, 0))
    images, tg = torch.stack(images, 0), torch.stack(tg, 0)
    # mean over the GLOBAL batch (this arm is single-GPU: global_batch == len(cams))
    loss = LAMBDA_L1 * (images - tg).abs().sum() / float(global_batch * images[0].numel())
    loss.backward()
    with torch.no_grad():
        grad = torch.zeros_like(vsp[0])
        for v in vsp:
            grad = grad + v.grad
        vis = radii_max > 0
        m.max_radii2D[vis] = torch.max(m.max_radii2D[vis], radii_max[vis].float())
        m.xyz_gradient_accum[vis] += torch.norm(grad[vis, :2], dim=-1, keepdim=True)
        m.denom[vis] += 1
    m.optimizer.step()
    m.optimizer.zero_grad(set_to_none=True)
    return loss.detach()


def fit_restorer(model):
    """restore() puts a fit.FitModel back to the state it has now (parameters, Adam, statistics): the fit moves the
    Gaussians — and with them the instances per view — step by step, so every timed region starts from the SAME
    state, the initial scene, and `value` and `e2e` time the same steps."""
    snap = {k: getattr(model, k).clone() for k in ("flat", "exp_avg", "exp_avg_sq", "xyz_gradient_accum", "denom",
                                                     "max_radii2D")}

    def restore():
        torch.cuda.synchronize()
        for k, v in snap.items():
            getattr(model, k).copy_(v)
        model.step_count = 0
        model.parameters_changed()  # a prefetched front half belongs to the old parameters
    return restore


def dge_restorer(model):
    """The same for a DGEStyleModel (reference arm, drop-in loop)."""
    tensors = [model._xyz, model._features_dc, model._features_rest, model._opacity, model._scaling, model._rotation,
               model.xyz_gradient_accum, model.denom, model.max_radii2D]
    snap = [t.detach().clone() for t in tensors]

    def restore():
        torch.cuda.synchronize()
        with torch.no_grad():
            for t, v in zip(tensors, snap):
                t.copy_(v)
        model.optimizer.state.clear()  # Adam restarts at step 0 with zero moments
    return restore


class Workload:
    """The synthetic scene, this rank's cameras (pinned host and device copies) and its target images."""

    def __init__(self, cfg, dev, views_global, mine, device_targets=False):
        from dge_b200 import scene  # synthetic scene and camera conventions only: pure torch, loads no library
        self.cfg, self.dev = cfg, dev
        P, W, H = cfg["P"], cfg["W"], cfg["H"]
        self.g = scene.make_gaussians(P, seed=cfg["seed"])
        self.ring = scene.ring_cameras(views_global, W, H)
        self.mine = list(mine)
        pin = lambda t: t.pin_memory() if isinstance(t, torch.Tensor) else t
        self.cams_host = [scene.Camera(*[pin(t) for t in self.ring[i]]) for i in self.mine]
        self.cams_dev = [scene.camera_to(c, dev) for c in self.cams_host]
        if device_targets:  # large configs: draw on the device (seeded), keep a pinned host copy for e2e
            gen = torch.Generator(device=dev).manual_seed(cfg["seed"] + 17)
            allt = torch.rand(views_global, 3, H, W, device=dev, generator=gen)
            self.targets_stacked = allt[self.mine].contiguous()
            del allt
            host = torch.empty(self.targets_stacked.shape, dtype=torch.float32).pin_memory()
            host.copy_(self.targets_stacked)
            self.targets_host = list(host.unbind(0))
            self.targets_all = None
        else:
            gen = torch.Generator().manual_seed(cfg["seed"] + 17)
            self.targets_all = [torch.rand(3, H, W, generator=gen) for _ in range(views_global)]
            self.targets_host = [self.targets_all[i].pin_memory() for i in self.mine]
            self.targets_stacked = torch.stack([t.to(dev) for t in self.targets_host])
        self.targets_dev = list(self.targets_stacked.unbind(0))
        self.bg = torch.zeros(3, device=dev)


def summarize(total_ms, per_step, views_per_step):
    med = statistics.median(per_step)
    return {"ms_per_step": total_ms / len(per_step), "ms_per_step_median": med, "ms_per_step_min": min(per_step),
            "ms_per_step_max": max(per_step), "value": views_per_step * len(per_step) / (total_ms * 1e-3),
            "value_median": views_per_step / (med * 1e-3)}


# ------------------------------------------------------------------ reference arm -----
def run_reference(args, cfg):
    """Nothing of this repository's product is imported here: dge_b200.scene (pure torch) and oracle/ only."""
    from oracle import ref
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    P, W, H, V = cfg["P"], cfg["W"], cfg["H"], cfg["V"]
    wl = Workload(cfg, dev, V, range(V))
    model = DGEStyleModel(wl.g, dev)
    rasterize = make_reference_rasterize()
    lib = ref.load()
    probe = ClockProbe(lambda p: lib.ref_clock_probe(p), dev)

    def step(host):
        return dge_style_step(model, rasterize, Settings, wl.cams_host if host else wl.cams_dev,
                              wl.targets_host if host else wl.targets_dev, wl.bg, V, host, dev)

    barrier = torch.cuda.synchronize
    restore = dge_restorer(model)  # every timed region starts from the initial scene

    W_ = max(args.warmup, 3)
    for _ in range(W_):
        step(False)
    barrier()
    # workload statistics of this arm's own forward (untimed), on the initial scene
    restore()
    with torch.no_grad():
        Rs, Pvs = [], []
        for cam in wl.cams_dev[:4]:
            _c, radii, _d = rasterize(Settings(cam, wl.bg, 3), model._xyz, None,
                                      torch.cat((model._features_dc, model._features_rest), dim=1),
                                      torch.sigmoid(model._opacity), torch.exp(model._scaling),
                                      torch.nn.functional.normalize(model._rotation))
            Rs.append(int(rasterize.cls.last_R))
            Pvs.append(int((radii > 0).sum()))
    stats = {"P": P, "P_visible": sum(Pvs) / len(Pvs), "R": sum(Rs) / len(Rs), "views_sampled": len(Rs)}
    restore()
    for _ in range(W_):
        step(False)
    total, per = timed_steps(lambda: step(False), args.steps, barrier, probe)
    mhz = probe.mhz()
    res = summarize(total, per, V)
    restore()
    for _ in range(W_):
        step(True).item()
    last = [None]

    def e2e_step():
        last[0] = float(step(True).item())
    total2, per2 = timed_steps(e2e_step, args.steps, barrier, probe)
    mhz2 = probe.mhz()
    e2e = summarize(total2, per2, V)
    clocks = nvml_clocks_under_load(lambda: step(False), barrier, dev.index or 0, res["ms_per_step"], args.steps)
    clocks.update(sm_mhz_nvml=clocks.get("sm_mhz"), sm_mhz=mhz if mhz is not None else clocks.get("sm_mhz"),
                  sm_mhz_e2e=mhz2,
                  sampled="sm_mhz / sm_mhz_e2e: measured on the device over the timed regions (per-SM clock64 / "
                          "globaltimer probes at both ends); reasons and sm_mhz_nvml: NVML over identical steps run "
                          "right after them")
    extras = {}
    if args.extras:
        del model
        torch.cuda.empty_cache()
        extras["boundary"] = measure_boundary(rasterize, Settings, wl, dev, max(3, min(args.steps, 6)), 3)
    assert "dge_b200._lib" not in sys.modules and "dge_b200.fit" not in sys.modules
    out = {
        "metric": METRIC, "value": res["value"], "unit": "views/s", "n_gpus": 1, "steps": args.steps, "warmup": W_,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args.config, cfg, 1),
        "workload_stats": stats, "timing": {"resident": res, "e2e": e2e}, "clocks": clocks,
        "e2e": {"value": e2e["value"], "unit": "views/s", "h2d_bytes_per_step": V * (3 * H * W * 4 + (16 + 16 + 3) * 4),
                "d2h_bytes_per_step": 4, "ms_per_step": e2e["ms_per_step"], "loss": last[0]},
        "gpu_launches": None, "impl": "reference", "extras": extras,
        "path": "DGE training step (render() per view with per-view activations, stacked L1, one backward, "
                "on_before_optimizer_step statistics, torch.optim.Adam) around the unmodified reference rasterizer",
        "cpu_baseline": {"value": res["value"], "unit": "views/s", "cores": os.cpu_count(), "kind": "reference",
                         "sample": "the reference ships no CPU rasterizer; this arm runs its UNMODIFIED CUDA code rebuilt "
                                   "for sm_100a (oracle/_ref) on one B200 of the same box inside DGE's own step shape"},
    }
    print(json.dumps(out), flush=True)
    return 0


def workload_config(name, cfg, n_gpus, views_global=None):
    """What defines the workload — identical in both arms."""
    V = cfg["V"]
    return {"workload": f"{name}: {cfg['desc']}", "gaussians": cfg["P"], "resolution": [cfg["W"], cfg["H"]],
            "views_per_step_per_gpu": V, "global_batch": views_global if views_global is not None else V * n_gpus,
            "sh_degree": 3, "loss": "10 * L1 (mean over the global batch)", "optimizer": "Adam, 6 groups, eps 1e-15",
            "scene": "randgauss-v1",
            "l2": "per-view working set (inputs 236 MB + scratch) exceeds the 126 MB L2; no explicit flush"}


# -------------------------------------------------------------------- cpu baseline -----
def cpu_baseline(cfg, g, cams, targets, budget_s=12.0):
    """The CPU restatement (oracle/, kind "port") on the first views of the same workload, fwd+bwd:
    whole views until `budget_s` seconds of CPU work have been spent (at least one)."""
    from oracle import oracle
    import numpy as np
    a = dict(shs=g.shs.numpy(), scales=g.scales.numpy(), rotations=g.rotations.numpy())
    n, t0 = 0, time.perf_counter()
    for cam, target in zip(cams, targets):
        tfx, tfy = math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5)
        common = (cam.world_view_transform.numpy(), cam.full_proj_transform.numpy(), cam.camera_center.numpy(),
                  np.zeros(3, np.float32), cfg["W"], cfg["H"], tfx, tfy)
        fw = oracle.forward(g.means3D.numpy(), g.opacities.numpy(), *common, **a)
        dL = np.sign(fw["out_color"] - target.numpy()).astype(np.float32) / (3.0 * cfg["W"] * cfg["H"])
        oracle.backward(fw, dL, g.means3D.numpy(), *common, **a)
        n += 1
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "views/s", "cores": int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1)),
            "kind": "port",
            "sample": f"{n} views fwd+bwd of the same workload ({cfg['P']} Gaussians, {cfg['W']}x{cfg['H']}) with oracle/splat_oracle.c "
                      f"(OpenMP over Gaussians and pixels); {dt:.1f} s"}


def algorithmic_bytes(stage, P, Pv, R, Npix, T, n_pass=6):
    """SURVEY.md §8d, per view, with the reference's stage decomposition."""
    return {
        "preprocess": 20 * P + 291 * Pv,
        "binning": 16 * P + 12 * Pv + (28 + 24 * n_pass) * R + 16 * T,  # K2+K3+K4+K5 (depth sort + binning here)
        "render_fwd": 44 * R + 24 * Npix + 8 * T,
        "render_bwd": 40 * R + 20 * Npix + 8 * T + 44 * Pv,
        "geom_bwd": 8 * P + 615 * Pv + 284 * (P - Pv),
    }[stage]


STAGE_NAMES = ["preprocess", "depth_sort", "binning", "render_fwd", "render_bwd", "geom_bwd", "apply_weights"]
KERNEL_OF_STAGE = {"preprocess": "preprocess_kernel", "render_fwd": "render_forward_kernel",
                   "render_bwd": "render_backward_kernel", "geom_bwd": "geom_backward_kernel",
                   "binning": "onesweep_kernel+expand_kernel", "depth_sort": "onesweep_kernel"}


# ----------------------------------------------------------------------- our arm -----
class Dist:
    """torch.distributed plumbing of one bench process (NCCL, one rank per GPU)."""

    def __init__(self, dev):
        import torch.distributed as dist
        self.dist, self.dev = dist, dev
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.nccl_log = NCCL_LOG
        if self.world > 1:
            dist.init_process_group("nccl", device_id=dev)

    def nccl_summary(self):
        """{version, nvls, channels...} read off rank 0's NCCL INIT log; None on one GPU or when NCCL_DEBUG was set
        by the caller."""
        if not self.nccl_log:
            return None
        if not os.path.exists(self.nccl_log):
            return {"version": ".".join(str(v) for v in torch.cuda.nccl.version()), "log": "no INIT log was written"}
        import re
        text = open(self.nccl_log, errors="replace").read()
        out = {"version": ".".join(str(v) for v in torch.cuda.nccl.version()), "nvls": False,
               "NCCL_DEBUG_before": NCCL_DEBUG_WAS}
        m = re.search(r"NVLS multicast support is (\w+)", text)
        if m:
            out["nvls_multicast"] = m.group(1)
        m = re.search(r"(\d+) coll channels, (\d+) collnet channels, (\d+) nvls channels, (\d+) p2p channels", text)
        if m:
            out.update(coll_channels=int(m.group(1)), nvls_channels=int(m.group(3)), p2p_channels=int(m.group(4)))
            out["nvls"] = int(m.group(3)) > 0
        elif re.search(r"Connected NVLS", text):
            out["nvls"] = True
        out["init_lines"] = [l.split("NCCL INFO ", 1)[-1][:160] for l in text.splitlines()
                             if re.search(r"NVLS|coll channels|Connected all|threadThresholds", l)][:8]
        return out

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_(self, values):
        t = torch.tensor(values, device=self.dev, dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def sample_view_stats(model, wl, n=4):
    """P_visible and R (the REFERENCE's instance count: unpruned lists) of this rank's first views."""
    from dge_b200 import scene
    from dge_b200 import diff_gaussian_rasterization as dgr
    with torch.no_grad():
        a = model.activations()
        Rs, Pvs = [], []
        e = torch.empty(0, device=wl.dev)
        for cam in wl.cams_dev[:n]:
            rs = scene.raster_settings(cam, wl.bg, 3, module=dgr)
            R, _c, _d, radii, *_ = dgr._forward_call(rs, a["means3D"], e, a["opacities"], a["scales"], a["rotations"], e, a["shs"])
            Rs.append(R)
            Pvs.append(int((radii > 0).sum()))
    st = {"P": model.P, "P_visible": sum(Pvs) / max(len(Pvs), 1), "R": sum(Rs) / max(len(Rs), 1), "views_sampled": len(Rs)}
    if getattr(model, "_batches", None):
        nr = [n_ for vb_ in model._batches for n_ in vb_.num_rendered]
        st["R_listed"] = sum(nr) / max(len(nr), 1)  # instances the per-step family lists (pruned lists)
    return st


def measure_fit(args, cfg, d, views_global, K, W_, name, scaling, device_targets=True):
    """An extra workload through fit.fit_step: resident and end-to-end views/s, max over ranks."""
    from dge_b200 import fit
    dev = d.dev
    mine = fit.shard_views(views_global, d.rank, d.world)
    wl = Workload(cfg, dev, views_global, mine, device_targets=device_targets)
    model = fit.FitModel(wl.g, dev, lrs=LRS)
    W, H = cfg["W"], cfg["H"]

    def step(host):
        cams = wl.cams_host if host else wl.cams_dev
        return fit.fit_step(model, cams, wl.targets_host if host else wl.targets_stacked,
                            wl.bg, global_batch=views_global, host_inputs=host, image_size=(W, H),
                            next_cameras=cams if d.world > 1 else None)

    restore = fit_restorer(model)
    for _ in range(W_):
        step(False)
    total, per = timed_steps(lambda: step(False), K, d.barrier)
    restore()
    for _ in range(W_):
        step(True).item()
    total2, per2 = timed_steps(lambda: step(True).item(), K, d.barrier)
    restore()
    total, total2 = d.max_([total, total2])
    res, e2e = summarize(total, per, views_global), summarize(total2, per2, views_global)
    stats = sample_view_stats(model, wl, n=2) if wl.cams_dev else {}
    out = {"workload": f"{name}: {cfg['desc']}", "value": res["value"], "unit": "views/s", "scaling": scaling,
           "n_gpus": d.world, "views_per_step": views_global, "views_per_step_per_gpu": len(mine), "steps": K,
           "ms_per_step": res["ms_per_step"], "ms_per_step_median": res["ms_per_step_median"],
           "e2e": {"value": e2e["value"], "ms_per_step": e2e["ms_per_step"], "ms_per_step_median": e2e["ms_per_step_median"],
                   "h2d_bytes_per_step": views_global * (3 * H * W * 4 + 35 * 4), "d2h_bytes_per_step": 4 * d.world},
           "gaussians": cfg["P"], "resolution": [W, H], **stats}
    del model, wl
    torch.cuda.empty_cache()
    return out


def measure_dropin(args, cfg, dev, K, W_):
    """SURVEY.md §8d's boundary: the reference arm's loop (dge_style_step) with ONLY the import swapped —
    GaussianRasterizer.forward + autograd per view behind dge_b200.install(), everything else torch."""
    import dge_b200
    dge_b200.install()
    import diff_gaussian_rasterization as dgr
    V, W, H = cfg["V"], cfg["W"], cfg["H"]
    wl = Workload(cfg, dev, V, range(V))
    model = DGEStyleModel(wl.g, dev)

    def make_settings(cam, bg, deg):
        return dgr.GaussianRasterizationSettings(
            image_height=int(cam.image_height), image_width=int(cam.image_width), tanfovx=math.tan(cam.FoVx * 0.5),
            tanfovy=math.tan(cam.FoVy * 0.5), bg=bg, scale_modifier=1.0, viewmatrix=cam.world_view_transform,
            projmatrix=cam.full_proj_transform, sh_degree=deg, campos=cam.camera_center, prefiltered=False, debug=False)

    def rasterize(rs, means3D, means2D, shs, opacities, scales, rotations):
        return dgr.GaussianRasterizer(raster_settings=rs)(means3D=means3D, means2D=means2D, shs=shs, colors_precomp=None,
                                                          opacities=opacities, scales=scales, rotations=rotations,
                                                          cov3D_precomp=None)

    def step(host):
        return dge_style_step(model, rasterize, make_settings, wl.cams_host if host else wl.cams_dev,
                              wl.targets_host if host else wl.targets_dev, wl.bg, V, host, dev)

    barrier = torch.cuda.synchronize
    restore = dge_restorer(model)
    for _ in range(W_):
        step(False)
    total, per = timed_steps(lambda: step(False), K, barrier)
    restore()
    for _ in range(W_):
        step(True).item()
    total2, per2 = timed_steps(lambda: step(True).item(), K, barrier)
    res, e2e = summarize(total, per, V), summarize(total2, per2, V)
    del model, restore
    torch.cuda.empty_cache()
    boundary = measure_boundary(rasterize, make_settings, wl, dev, K, W_)
    del wl
    torch.cuda.empty_cache()
    return {"workload": "config2 through the drop-in boundary: the reference arm's loop, import swapped",
            "boundary": boundary,
            "value": res["value"], "unit": "views/s", "ms_per_step": res["ms_per_step"],
            "ms_per_step_median": res["ms_per_step_median"], "steps": K,
            "e2e": {"value": e2e["value"], "ms_per_step": e2e["ms_per_step"], "ms_per_step_median": e2e["ms_per_step_median"]}}


def measure_boundary(rasterize, make_settings, wl, dev, K, W_):
    """SURVEY.md §8d's metric at its own boundary: GaussianRasterizer.forward + the autograd backward per view,
    nothing else — activated tensors given as leaves, a fixed random dL/dcolor as the upstream gradient, no
    activations, loss or optimiser. views/s over the workload's views, K passes."""
    g = wl.g
    leaves = [t.to(dev).contiguous().requires_grad_(True) for t in (g.means3D, g.shs, g.opacities, g.scales, g.rotations)]
    m2d = torch.zeros(g.means3D.shape[0], 3, device=dev, requires_grad=True)
    gen = torch.Generator().manual_seed(99)
    H, W = wl.cfg["H"], wl.cfg["W"]
    dL = (torch.randn(3, H, W, generator=gen) / (3.0 * H * W)).to(dev)

    def one_pass():
        for cam in wl.cams_dev:
            for t in leaves + [m2d]:
                t.grad = None  # AccumulateGrad then keeps the returned tensor: no add kernels in the loop
            color, _radii, _depth = rasterize(make_settings(cam, wl.bg, 3), leaves[0], m2d, leaves[1], leaves[2], leaves[3],
                                              leaves[4])
            color.backward(dL)

    for _ in range(W_):
        one_pass()
    total, per = timed_steps(one_pass, K, torch.cuda.synchronize)
    r = summarize(total, per, len(wl.cams_dev))
    return {"workload": "config2 scene, per view: GaussianRasterizer.forward + autograd backward with a fixed dL/dcolor "
                        "(SURVEY.md 8d boundary; no activations, loss or optimiser)",
            "value": r["value"], "unit": "views/s", "value_median": r["value_median"],
            "ms_per_view": r["ms_per_step"] / len(wl.cams_dev), "passes": K, "views_per_pass": len(wl.cams_dev)}


def measure_config3(args, cfg, dev, K, W_, per_view=False, impl="ours"):
    """Mask back-projection (DGE.update_mask, threestudio/systems/DGE.py:101-165): a step = apply_weights of a
    binary disc mask over all V views into weights/cnt, then the selection weights/(cnt+1e-7) > mask_thres.
    `value`: masks resident; `e2e`: masks in pinned host memory copied per view, the selection read back."""
    from dge_b200 import scene
    P, W, H, V = cfg["P"], cfg["W"], cfg["H"], cfg["V"]
    g = scene.make_gaussians(P, seed=cfg["seed"])
    gd = scene.Gaussians(*[t.to(dev) for t in g])
    cams = [scene.camera_to(c, dev) for c in scene.ring_cameras(V, W, H)]
    bg = torch.zeros(3, device=dev)
    mask_host = scene.disc_mask(W, H, radius=160.0).pin_memory()
    mask_dev = mask_host.to(dev)
    weights = torch.zeros(P, 1, device=dev)
    cnt = torch.zeros(P, 1, dtype=torch.int32, device=dev)
    e = torch.empty(0, device=dev)
    if impl == "reference":
        from oracle import ref
        tans = [(math.tan(c.FoVx * 0.5), math.tan(c.FoVy * 0.5)) for c in cams]
        launches = lambda: 0
    else:
        from dge_b200 import fit
        from dge_b200 import _lib as L
        from dge_b200 import diff_gaussian_rasterization as dgr
        lib = L.load()
        launches = lib.dge_launch_count
    masks_host = [mask_host for _ in cams]
    masks_dev = torch.stack([mask_dev for _ in cams])
    masks_stage = torch.empty_like(masks_dev)

    def step(host):
        weights.zero_()
        cnt.zero_()
        if impl == "ours" and not per_view:
            if host:  # pinned host masks -> one staging block, asynchronously
                for i, t in enumerate(masks_host):
                    masks_stage[i].copy_(t, non_blocking=True)
            m = masks_stage if host else masks_dev
            fit.backproject_masks(gd.means3D, gd.opacities, gd.scales, gd.rotations, cams, m, weights, cnt)
            return (weights / (cnt + 1e-7)) > 0.8
        for i, cam in enumerate(cams):
            m = mask_host.to(dev, non_blocking=True) if host else mask_dev
            if impl == "ours":
                rs = scene.raster_settings(cam, bg, 0, module=dgr)
                dgr.GaussianRasterizer(rs).apply_weights(gd.means3D, None, gd.opacities, None, weights, gd.scales,
                                                         gd.rotations, None, cnt, m)
            else:
                ref.apply_weights(bg, gd.means3D, weights, gd.opacities, gd.scales, gd.rotations, 1.0, e,
                                  cam.world_view_transform, cam.full_proj_transform, tans[i][0], tans[i][1], H, W, e, 0,
                                  cam.camera_center, False, m, cnt, False)
        return (weights / (cnt + 1e-7)) > 0.8   # DGE.py:149-152, dge.yaml mask_thres

    barrier = torch.cuda.synchronize
    for _ in range(W_):
        step(False)
    barrier()
    l0 = launches()
    total, per = timed_steps(lambda: step(False), K, barrier)
    n_launch = launches() - l0
    for _ in range(W_):
        step(True)
    n_sel = [0]

    def e2e_step():
        n_sel[0] = int(step(True).sum().item())
    total2, per2 = timed_steps(e2e_step, K, barrier)
    res, e2e = summarize(total, per, V), summarize(total2, per2, V)
    out = {"workload": f"config3: {cfg['desc']}", "value": res["value"], "unit": "views/s", "ms_per_step": res["ms_per_step"],
           "ms_per_step_median": res["ms_per_step_median"], "steps": K, "views_per_step": V,
           "views_per_launch": 1 if (impl != "ours" or per_view) else V, "selected": n_sel[0],
           "gpu_launches": int(n_launch) if impl == "ours" else None,
           "e2e": {"value": e2e["value"], "ms_per_step": e2e["ms_per_step"], "ms_per_step_median": e2e["ms_per_step_median"],
                   "h2d_bytes_per_step": V * W * H * 4, "d2h_bytes_per_step": 8}}
    del gd, weights, cnt, masks_dev, masks_stage
    torch.cuda.empty_cache()
    return out


def main_config3(args, cfg):
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    r = measure_config3(args, cfg, dev, args.steps, max(args.warmup, 3), per_view=args.streams > 0, impl=args.impl)
    out = {"metric": "mask back-projection views/s @1M Gaussians 512^2 (DGE.update_mask: apply_weights over 40 views)",
           "value": r["value"], "unit": "views/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": r["workload"], "gaussians": cfg["P"], "resolution": [cfg["W"], cfg["H"]],
                      "views_per_step": cfg["V"], "mask": "binary disc, radius 160 px", "scene": "randgauss-v1"},
           "detail": r, "e2e": dict(r["e2e"], unit="views/s"), "gpu_launches": r["gpu_launches"]}
    if args.impl == "reference":
        out["impl"] = "reference"
    print(json.dumps(out), flush=True)
    return 0


def run_ours(args, cfg):
    import ctypes as C
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    d = Dist(dev)
    rank, world = d.rank, d.world
    from dge_b200 import fit
    from dge_b200 import _lib as L
    lib = L.load()

    P, W, H, V = cfg["P"], cfg["W"], cfg["H"], cfg["V"]
    # every rank fits its own V views of a ring of V*world cameras (view i -> rank i mod world)
    wl = Workload(cfg, dev, V * world, fit.shard_views(V * world, rank, world), device_targets=cfg["P"] > 1_500_000)
    model = fit.FitModel(wl.g, dev, lrs=LRS, fused_adam=True)
    batched = args.streams == 0
    pipeline = (world > 1 and args.pipeline != 0) or args.pipeline == 1

    def step(host):
        tg = wl.targets_host if host else (wl.targets_stacked if batched else wl.targets_dev)
        cams = wl.cams_host if host else wl.cams_dev
        # several GPUs: the next step's cameras are known (here: the same ones), so its projection / depth sort /
        # binning run under this step's f_rest all-reduce (fit.fit_step: next_cameras)
        return fit.fit_step(model, cams, tg, wl.bg, global_batch=V * world,
                            host_inputs=host, num_streams=max(args.streams, 1), batched=batched, num_chunks=args.chunks,
                            geom_splits=args.geom_splits or None, image_size=(W, H),
                            next_cameras=cams if (pipeline and batched) else None)

    restore = fit_restorer(model)  # every timed region starts from the initial scene

    W_ = max(args.warmup, 3)
    for _ in range(W_):
        step(False)
    d.barrier()

    # ---- diagnostic pass (untimed) on the initial scene: per-stage device time, R and P_v of this rank's views
    restore()
    lib.dge_profile_enable((1 << 7) - 1)
    step(False)
    torch.cuda.synchronize()
    ms, cnt = (C.c_float * 7)(), (C.c_int * 7)()
    L.check(lib.dge_profile_read(ms, cnt), "profile read")
    stage_ms = {n: ms[i] / max(cnt[i], 1) for i, n in enumerate(STAGE_NAMES) if cnt[i]}
    stage_total = {n: ms[i] for i, n in enumerate(STAGE_NAMES) if cnt[i]}
    stage_launches = {n: int(cnt[i]) for i, n in enumerate(STAGE_NAMES) if cnt[i]}
    lib.dge_profile_enable(0)
    restore()
    stats = sample_view_stats(model, wl)
    dominant = max((k for k in stage_total if k in ("preprocess", "render_fwd", "render_bwd", "geom_bwd", "binning")),
                   key=lambda k: stage_total[k])  # largest share of the step
    restore()
    for _ in range(W_):  # (also settles the caching allocator again after the diagnostic allocations)
        step(False)
    # the dominant stage is event-timed live, inside the timed region, on the stream it is launched on
    lib.dge_profile_enable(1 << STAGE_NAMES.index(dominant))
    lib.dge_profile_read((C.c_float * 7)(), (C.c_int * 7)())

    # ---- timed region 1: K steps, inputs resident in HBM
    probe = ClockProbe(lambda p: L.check(lib.dge_clock_probe(p, L.stream_ptr(dev)), "clock probe"), dev)
    launches0 = lib.dge_launch_count()
    total, per = timed_steps(lambda: step(False), args.steps, d.barrier, probe)
    launches = lib.dge_launch_count() - launches0
    mhz = probe.mhz()
    ms, cnt = (C.c_float * 7)(), (C.c_int * 7)()
    L.check(lib.dge_profile_read(ms, cnt), "profile read")
    lib.dge_profile_enable(0)
    i = STAGE_NAMES.index(dominant)
    avg_ms = ms[i] / max(cnt[i], 1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, which = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    T = ((W + 15) // 16) * ((H + 15) // 16)
    # SURVEY.md §8d per-view figure x the views one launch of the stage processes
    units = V if dominant == "geom_bwd" else (-(-V // args.chunks) if batched else 1)
    abytes = units * algorithmic_bytes(dominant, P, stats["P_visible"], stats["R"], W * H, T)
    achieved = abytes / (avg_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(KERNEL_OF_STAGE[dominant])
        traffic = traffic * units if traffic is not None else None  # the capture is of ONE view
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": KERNEL_OF_STAGE[dominant], "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": which,
            "avg_launch_ms": avg_ms, "launches_timed": int(cnt[i]), "algorithmic_bytes_per_launch": abytes,
            "views_per_launch": units,
            "note": "blend kernels are issue bound, not HBM bound (SURVEY.md §8d); see DESIGN.md"}

    # ---- timed region 2: K steps end to end (pinned host inputs copied inside, loss read back), from the
    # same initial state, after its own W warm-up steps (the first host-input step allocates the staging buffers)
    restore()
    for _ in range(W_):
        step(True).item()
    last = [None]

    def e2e_step():
        last[0] = float(step(True).item())
    total2, per2 = timed_steps(e2e_step, args.steps, d.barrier, probe)
    mhz2 = probe.mhz()

    # ---- throttle reasons under the same load (NVML, right after the timed regions)
    def sync_count(n):
        return int(d.max_([float(n)])[0])
    if rank == 0:
        clocks = nvml_clocks_under_load(lambda: step(False), d.barrier, local_rank, total / args.steps, args.steps,
                                        sync_count if world > 1 else None)
        clocks.update(sm_mhz_nvml=clocks.get("sm_mhz"), sm_mhz=mhz if mhz is not None else clocks.get("sm_mhz"),
                      sm_mhz_e2e=mhz2,
                      sampled="sm_mhz / sm_mhz_e2e: measured on the device over the timed regions (per-SM clock64 / "
                              "globaltimer probes at both ends); reasons and sm_mhz_nvml: NVML over identical steps "
                              "run right after them")
    else:
        n = sync_count(0)
        for _ in range(n):
            step(False)
        d.barrier()
        clocks = None
    total, total2 = d.max_([total, total2])
    res, e2e = summarize(total, per, V * world), summarize(total2, per2, V * world)
    del model
    torch.cuda.empty_cache()

    # ---- the other configurations, on the same line (smaller K: they only have to be stable)
    extras = {}
    if args.extras and args.config == "config2":
        Kx, Wx = max(3, min(args.steps, 6)), 3
        try:
            extras["config4_strong"] = measure_fit(args, CONFIGS["config4"], d, 64, Kx, Wx, "config4 (strong: 64 views "
                                                   "split over the ranks)", "strong")
            if world == 1:
                extras["dropin"] = measure_dropin(args, cfg, dev, Kx, Wx)
                extras["boundary"] = extras["dropin"].pop("boundary")
                extras["config3"] = measure_config3(args, CONFIGS["config3"], dev, Kx, Wx)
                extras["config5"] = measure_fit(args, CONFIGS["config5"], d, CONFIGS["config5"]["V"], Kx, Wx,
                                                "config5 (2M point)", "weak")
        except Exception as ex:  # an extra must never cost the headline line
            extras["error"] = f"{type(ex).__name__}: {ex}"
    if rank != 0:
        d.close()
        return 0

    h2d = world * V * (3 * H * W * 4 + (16 + 16 + 3) * 4)  # whole job: every rank copies its own views
    out = {
        "metric": METRIC, "value": res["value"], "unit": "views/s", "n_gpus": world, "steps": args.steps, "warmup": W_,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args.config, cfg, world),
        "workload_stats": stats, "timing": {"resident": res, "e2e": e2e},
        "path": {"impl": "per-step C-ABI family (fit.fit_step)", "parallelism": f"dp{world} (views)",
                 "views_per_launch": -(-V // args.chunks) if batched else 1, "chunks": args.chunks if batched else None,
                 "streams_per_gpu": max(args.streams, 1), "next_step_front_prefetched": bool(pipeline and batched)},
        "clocks": clocks,
        "e2e": {"value": e2e["value"], "unit": "views/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * world,
                "ms_per_step": e2e["ms_per_step"], "loss": last[0]},
        "gpu_launches": int(launches), "roofline": roof, "stages_ms_per_launch": stage_ms,
        "stages_launches_per_step": stage_launches,
    }
    # sort keys/s. One launch of a stage covers `vpl` views: the depth sort orders P
    # (depth bits, id) pairs per view, binning turns them into R_listed (tile, id) instances per view
    # in (tile, depth, id) order — the work the reference does as ONE sort of R 64-bit keys.
    vpl = (-(-V // args.chunks)) if batched else 1
    listed = stats.get("R_listed", stats.get("R", 0.0))
    rates = {}
    if stage_ms.get("depth_sort"):
        rates["depth_sort_pairs_per_s"] = vpl * P / (stage_ms["depth_sort"] * 1e-3)
    if stage_ms.get("binning"):
        rates["binning_instances_per_s"] = vpl * listed / (stage_ms["binning"] * 1e-3)
        if stage_ms.get("depth_sort"):
            rates["reference_equivalent_sort_keys_per_s"] = vpl * stats.get("R", 0.0) / (
                (stage_ms["binning"] + stage_ms["depth_sort"]) * 1e-3)
    out["rates"] = rates
    nccl = d.nccl_summary()
    if nccl:
        out["nccl"] = nccl
    if extras:
        out["extras"] = extras
    if not args.no_cpu_baseline and world == 1 and wl.targets_all is not None:  # rank 0 at N=1 only
        try:
            out["cpu_baseline"] = cpu_baseline(cfg, wl.g, wl.ring, wl.targets_all)
        except Exception as ex:  # the checker is optional for the number, never for the tests
            out["cpu_baseline"] = {"value": None, "unit": "views/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
    print(json.dumps(out), flush=True)
    d.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="config2", choices=list(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", dest="extras", action="store_false",
                    help="only the headline workload (default: dropin / config3 / config4_strong / config5 ride along)")
    ap.add_argument("--gaussians", type=int, default=0, help="override the config's number of Gaussians (config 5 sweep)")
    ap.add_argument("--views", type=int, default=0, help="override the config's views per step per GPU")
    ap.add_argument("--geom-splits", type=int, default=0, help="per-Gaussian backward launches per step (0: "
                    "fit.py's default; > 1 on several GPUs sends finished ranges' f_rest rows early)")
    ap.add_argument("--pipeline", type=int, default=-1, help="-1 (default): prefetch the next step's front half under the "
                    "f_rest all-reduce when there are several GPUs; 0: never; 1: always (also on one GPU)")
    ap.add_argument("--chunks", type=int, default=1, help="batched path: split the step's views into this many "
                    "chunks, each on its own stream")
    ap.add_argument("--streams", type=int, default=0,
                    help="0 (default): all views of a step per launch (batched path); N>0: views one by one, "
                         "round-robin on N CUDA streams")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.gaussians:
        cfg["P"] = args.gaussians
    if args.views:
        cfg["V"] = args.views
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback of the product path)")
    rank = int(os.environ.get("RANK", "0"))
    if args.config == "config3":
        return 0 if rank != 0 else main_config3(args, cfg)
    if args.impl == "reference":
        # the reference is a single-GPU program: rank 0 alone runs it
        return 0 if rank != 0 else run_reference(args, cfg)
    return run_ours(args, cfg)


if __name__ == "__main__":
    sys.exit(main())
