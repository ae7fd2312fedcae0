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
the loss read back every step. `--impl reference` drives the UNMODIFIED reference rasterizer
rebuilt for sm_100a (oracle/_ref; the reference ships no CPU rasterizer) through the same
harness with torch.optim.Adam, on one GPU.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

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


# --------------------------------------------------------------------------- clocks -----
class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region (B200_PROFILING.md) through
    NVML in a background thread: a polling `nvidia-smi -lms` process contends for the driver
    and slows the measured step, NVML queries do not."""
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


# ------------------------------------------------------------------ reference arm -----
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

    return rasterize


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


def main_config3(args, cfg):
    """Mask back-projection (DGE.update_mask, threestudio/systems/DGE.py:101-165): a step = apply_weights of a
    binary disc mask over all V views into weights/cnt, then the selection weights/(cnt+1e-7) > mask_thres.
    `value`: masks resident; `e2e`: masks in pinned host memory copied per view, the selection read back."""
    import numpy as np
    from dge_b200 import scene
    from dge_b200 import _lib as L
    from dge_b200 import diff_gaussian_rasterization as dgr
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    P, W, H, V = cfg["P"], cfg["W"], cfg["H"], cfg["V"]
    g = scene.make_gaussians(P, seed=cfg["seed"])
    gd = scene.Gaussians(*[t.to(dev) for t in g])
    cams = [scene.camera_to(c, dev) for c in scene.ring_cameras(V, W, H)]
    bg = torch.zeros(3, device=dev)
    mask_host = scene.disc_mask(W, H, radius=160.0).pin_memory()
    mask_dev = mask_host.to(dev)
    weights = torch.zeros(P, 1, device=dev)
    cnt = torch.zeros(P, 1, dtype=torch.int32, device=dev)
    lib = L.load()
    e = torch.empty(0, device=dev)
    if args.impl == "reference":
        from oracle import ref
        tans = [(math.tan(c.FoVx * 0.5), math.tan(c.FoVy * 0.5)) for c in cams]

    from dge_b200 import fit
    masks_host = [mask_host for _ in cams]
    masks_dev = torch.stack([mask_dev for _ in cams])
    masks_stage = torch.empty_like(masks_dev)
    per_view = args.streams > 0  # --streams N>0: the per-view API loop; default: one batched call

    def step(host):
        weights.zero_()
        cnt.zero_()
        if args.impl == "ours" and not per_view:
            if host:  # pinned host masks -> one staging block, asynchronously
                for i, t in enumerate(masks_host):
                    masks_stage[i].copy_(t, non_blocking=True)
            m = masks_stage if host else masks_dev
            fit.backproject_masks(gd.means3D, gd.opacities, gd.scales, gd.rotations, cams, m, weights, cnt)
            return (weights / (cnt + 1e-7)) > 0.8
        for i, cam in enumerate(cams):
            m = mask_host.to(dev, non_blocking=True) if host else mask_dev
            if args.impl == "ours":
                rs = scene.raster_settings(cam, bg, 0, module=dgr)
                dgr.GaussianRasterizer(rs).apply_weights(gd.means3D, None, gd.opacities, None, weights, gd.scales,
                                                         gd.rotations, None, cnt, m)
            else:
                ref.apply_weights(bg, gd.means3D, weights, gd.opacities, gd.scales, gd.rotations, 1.0, e,
                                  cam.world_view_transform, cam.full_proj_transform, tans[i][0], tans[i][1], H, W, e, 0,
                                  cam.camera_center, False, m, cnt, False)
        sel = (weights / (cnt + 1e-7)) > 0.8   # DGE.py:149-152, dge.yaml mask_thres
        return sel

    for _ in range(max(args.warmup, 3)):
        step(False)
    torch.cuda.synchronize()
    l0 = lib.dge_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sel = step(False)
    e1.record()
    torch.cuda.synchronize()
    launches = lib.dge_launch_count() - l0
    ms = e0.elapsed_time(e1)
    for _ in range(max(args.warmup, 3)):
        step(True)
    torch.cuda.synchronize()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        n_sel = int(step(True).sum().item())
    e3.record()
    torch.cuda.synchronize()
    ms_e2e = e2.elapsed_time(e3)
    sampler = ClockSampler(0, period=0.05)
    sampler.start()
    for _ in range(args.steps):
        step(False)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    views = V * args.steps
    out = {"metric": "mask back-projection views/s @1M Gaussians 512^2 (DGE.update_mask: apply_weights over 40 views)",
           "value": views / (ms * 1e-3), "unit": "views/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": f"config3: {cfg['desc']}", "gaussians": P, "resolution": [W, H], "views_per_step": V,
                      "mask": "binary disc, radius 160 px", "selected": n_sel, "scene": "randgauss-v1",
                      "views_per_launch": 1 if (args.impl != "ours" or per_view) else V},
           "clocks": clocks,
           "e2e": {"value": views / (ms_e2e * 1e-3), "unit": "views/s", "h2d_bytes_per_step": V * W * H * 4,
                   "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / args.steps},
           "gpu_launches": int(launches) if args.impl == "ours" else None}
    if args.impl == "reference":
        out["impl"] = "reference"
    print(json.dumps(out), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="config2", choices=list(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gaussians", type=int, default=0, help="override the config's number of Gaussians (config 5 sweep)")
    ap.add_argument("--views", type=int, default=0, help="override the config's views per step per GPU")
    ap.add_argument("--geom-splits", type=int, default=0, help="per-Gaussian backward launches per step (0: "
                    "fit.py's default; > 1 on several GPUs sends finished ranges' f_rest rows early)")
    ap.add_argument("--chunks", type=int, default=1, help="batched path: split the step's views into this many "
                    "chunks, each on its own stream")
    ap.add_argument("--dropin", action="store_true",
                    help="ours with DGE's code unchanged: GaussianRasterizer.forward + autograd per view, torch ops "
                         "for the activations and the loss, torch.optim.Adam (the harness of --impl reference)")
    ap.add_argument("--streams", type=int, default=0,
                    help="0 (default): all views of a step per launch (batched path); N>0: views one by one, "
                         "round-robin on N CUDA streams")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.gaussians:
        cfg["P"] = args.gaussians
    if args.views:
        cfg["V"] = args.views
    if args.config == "config3":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        return main_config3(args, cfg)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0  # the reference is a single-GPU program: rank 0 alone runs it
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback of the product path)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    import torch.distributed as dist
    if world > 1 and args.impl == "ours":
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world if args.impl == "ours" else 1

    from dge_b200 import fit, scene
    from dge_b200 import _lib as L
    from dge_b200 import diff_gaussian_rasterization as dgr

    P, W, H, V = cfg["P"], cfg["W"], cfg["H"], cfg["V"]
    g = scene.make_gaussians(P, seed=cfg["seed"])
    # every rank fits its own 20 views of a ring of 20*n_gpus cameras (view i -> rank i mod world)
    ring = scene.ring_cameras(V * n_gpus, W, H)
    mine = fit.shard_views(V * n_gpus, rank if args.impl == "ours" else 0, n_gpus)
    gen = torch.Generator().manual_seed(cfg["seed"] + 17)
    targets_all = [torch.rand(3, H, W, generator=gen) for _ in range(V * n_gpus)]
    cams_host = [scene.Camera(*[t.pin_memory() if isinstance(t, torch.Tensor) else t for t in ring[i]]) for i in mine]
    targets_host = [targets_all[i].pin_memory() for i in mine]
    cams_dev = [scene.camera_to(c, dev) for c in cams_host]
    targets_dev = [t.to(dev) for t in targets_host]
    targets_stacked = torch.stack(targets_dev)  # resident [V,3,H,W] block (what the batched path consumes as is)
    bg = torch.zeros(3, device=dev)

    if args.impl == "ours":
        model = fit.FitModel(g, dev, fused_adam=not args.dropin)
        rasterize, module = fit.default_rasterize, dgr
    else:
        model = fit.FitModel(g, dev, fused_adam=False)
        rasterize, module = make_reference_rasterize(), dgr
    lib = L.load()

    batched = args.impl == "ours" and args.streams == 0 and not args.dropin

    def step(host):
        tg = targets_host if host else (targets_stacked if batched else targets_dev)
        return fit.fit_step(model, cams_host if host else cams_dev, tg, bg,
                            global_batch=V * n_gpus, rasterize=rasterize, settings_module=module, host_inputs=host,
                            num_streams=max(args.streams, 1) if args.impl == "ours" else 1, batched=batched,
                            num_chunks=args.chunks, geom_splits=args.geom_splits or None,
                            direct=False if args.dropin else None)

    def barrier():
        if world > 1 and args.impl == "ours":
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(False)
    barrier()

    # ---- diagnostic pass (untimed): per-stage device time, R and P_v of this rank's views
    stage_ms, stats = {}, {}
    if args.impl == "ours":
        import ctypes as C
        lib.dge_profile_enable((1 << 7) - 1)
        step(False)
        torch.cuda.synchronize()
        ms, cnt = (C.c_float * 7)(), (C.c_int * 7)()
        L.check(lib.dge_profile_read(ms, cnt), "profile read")
        stage_ms = {n: ms[i] / max(cnt[i], 1) for i, n in enumerate(STAGE_NAMES) if cnt[i]}
        stage_total = {n: ms[i] for i, n in enumerate(STAGE_NAMES) if cnt[i]}
        stage_launches = {n: int(cnt[i]) for i, n in enumerate(STAGE_NAMES) if cnt[i]}
        lib.dge_profile_enable(0)
        with torch.no_grad():
            a = model.activations()
            Rs, Pvs = [], []
            for cam in cams_dev[:4]:
                rs = scene.raster_settings(cam, bg, 3, module=dgr)
                e = torch.empty(0, device=dev)
                R, _c, _d, radii, *_ = dgr._forward_call(rs, a["means3D"], e, a["opacities"], a["scales"], a["rotations"], e, a["shs"])
                Rs.append(R)
                Pvs.append(int((radii > 0).sum()))
        stats = {"P": P, "P_visible": sum(Pvs) / len(Pvs), "R": sum(Rs) / len(Rs), "views_sampled": len(Rs)}
        if batched and getattr(model, "_batches", None):
            # instances actually listed per view by the per-step family (lists pruned to the tiles a
            # Gaussian can reach with alpha >= 1/255); R above is the reference's count
            nr = [n for vb_ in model._batches for n in vb_.num_rendered]
            stats["R_listed"] = sum(nr) / max(len(nr), 1)
        dominant = max((k for k in stage_total if k in ("preprocess", "render_fwd", "render_bwd", "geom_bwd", "binning")),
                       key=lambda k: stage_total[k])  # largest share of the step
        lib.dge_profile_enable(1 << STAGE_NAMES.index(dominant))
        lib.dge_profile_read((C.c_float * 7)(), (C.c_int * 7)())
    for _ in range(2):  # settle the caching allocator again after the diagnostic allocations
        step(False)
    if args.impl == "ours":
        lib.dge_profile_read((C.c_float * 7)(), (C.c_int * 7)())
    barrier()

    # ---- timed region 1: K steps, inputs resident in HBM
    launches0 = lib.dge_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(False)
    e1.record()
    barrier()
    ms_resident = e0.elapsed_time(e1)
    launches = lib.dge_launch_count() - launches0
    roof = None
    if args.impl == "ours":
        import ctypes as C
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
        # the ALU floor next to the byte roofline: warp instructions of the launch (ncu,
        # smsp__inst_executed.sum of the same kernel on this workload) over its live duration, against the
        # issue rate of the part: SMs x 4 schedulers x SM clock
        issue = None
        try:
            nvtx = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["warp_instructions"]
            wi = nvtx.get(KERNEL_OF_STAGE[dominant])
            if wi is not None and args.config == "config2":
                sms = torch.cuda.get_device_properties(dev).multi_processor_count
                peak_issue = sms * 4 * 1.965e9
                issue = {"warp_instructions_per_launch": wi * units, "achieved": wi * units / (avg_ms * 1e-3),
                         "peak": peak_issue, "unit": "warp-instr/s", "frac": wi * units / (avg_ms * 1e-3) / peak_issue,
                         "source": "instruction count from profiles/ncu_full_r1g.md (same kernel, same workload), "
                                   "duration live; peak = SMs x 4 schedulers x 1.965 GHz"}
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": KERNEL_OF_STAGE[dominant], "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": which,
                "avg_launch_ms": avg_ms, "launches_timed": int(cnt[i]), "algorithmic_bytes_per_launch": abytes,
                "views_per_launch": units,
                "issue": issue,
                "note": "blend kernels are issue/atomic bound, not HBM bound (SURVEY.md §8d); see DESIGN.md"}

    # ---- timed region 2: K steps end to end (pinned host inputs copied inside, loss read back),
    # after its own W warm-up steps (the first host-input step allocates the staging buffers, the copy
    # stream and the camera records of the host cameras)
    for _ in range(max(args.warmup, 3)):
        step(True).item()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last_loss = None
    for _ in range(args.steps):
        last_loss = float(step(True).item())
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)

    # ---- clocks under the same load: K more identical steps with NVML sampling in a side thread.
    # On this driver every NVML clock query stalls the measured process for ~25 ms (a 100 ms
    # nvidia-smi / NVML poll made a 13 ms step read 64 ms), so the sampling runs right AFTER the
    # two timed regions instead of inside them; the load, clocks and power state are the same.
    sampler = ClockSampler(local_rank, period=0.05)
    if rank == 0:
        sampler.start()
    # at least ~0.6 s of load so that the median rests on a dozen samples (same count on every rank:
    # the steps hold a collective)
    n_clock_steps = max(args.steps, min(200, int(math.ceil(600.0 / max(ms_resident / args.steps, 0.05)))))
    if world > 1 and args.impl == "ours":
        tn = torch.tensor([n_clock_steps], device=dev, dtype=torch.int64)
        dist.all_reduce(tn, op=dist.ReduceOp.MAX)
        n_clock_steps = int(tn[0])
    for _ in range(n_clock_steps):
        step(False)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["sampled"] = "NVML, during K identical steps run right after the timed regions (see bench.py)"

    t = torch.tensor([ms_resident, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1 and args.impl == "ours":
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_resident, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1 and args.impl == "ours":
            dist.destroy_process_group()
        return 0

    views = V * n_gpus * args.steps
    value = views / (ms_resident * 1e-3)
    e2e_value = views / (ms_e2e * 1e-3)
    h2d = n_gpus * V * (3 * H * W * 4 + (16 + 16 + 3) * 4)  # whole job: every rank copies its own views
    out = {
        "metric": "fwd+bwd views/s @1M Gaussians 512^2 (DGE 3D-fit step incl. Adam)", "value": value, "unit": "views/s",
        "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_resident / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {cfg['desc']}", "gaussians": P, "resolution": [W, H],
                   "views_per_step_per_gpu": V, "global_batch": V * n_gpus, "sh_degree": 3, "parallelism": f"dp{n_gpus} (views)",
                   "views_per_launch": -(-V // args.chunks) if batched else 1, "chunks": args.chunks if batched else None,
                   "streams_per_gpu": max(args.streams, 1) if args.impl == "ours" else 1,
                   "path": "autograd per view (drop-in import swap, DGE unchanged)" if args.dropin else
                           ("autograd per view" if args.impl != "ours" else "per-step C-ABI family (fit.fit_step)"),
                   "l2": "per-view working set (inputs 236 MB + scratch) exceeds the 126 MB L2; no explicit flush",
                   "scene": "randgauss-v1", **stats},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "views/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * n_gpus,
                "ms_per_step": ms_e2e / args.steps, "loss": last_loss},
        "gpu_launches": int(launches),
    }
    if args.impl == "ours":
        out["roofline"] = roof
        out["stages_ms_per_launch"] = stage_ms
        out["stages_launches_per_step"] = stage_launches
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
        if not args.no_cpu_baseline and n_gpus == 1:  # rank 0 at N=1 only (torchrun pins OMP_NUM_THREADS=1)
            try:
                out["cpu_baseline"] = cpu_baseline(cfg, g, ring, targets_all)
            except Exception as ex:  # the checker is optional for the number, never for the tests
                out["cpu_baseline"] = {"value": None, "unit": "views/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
    else:
        out["impl"] = "reference"
        out["gpu_launches"] = None
        out["cpu_baseline"] = {"value": value, "unit": "views/s", "cores": os.cpu_count(), "kind": "reference",
                               "sample": "the reference ships no CPU rasterizer; this arm runs its UNMODIFIED CUDA code rebuilt "
                                         "for sm_100a (oracle/_ref) on one B200 of the same box, same harness, torch.optim.Adam"}
    print(json.dumps(out), flush=True)
    if world > 1 and args.impl == "ours":
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
