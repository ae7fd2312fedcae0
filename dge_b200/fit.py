"""View-sharded 3D-fit step.

One process per GPU, every rank holds a full Gaussian replica. A step renders this rank's
share of the batch of edited views (forward + backward through the rasterizer), reduces the
parameter gradients and the densification statistics over the ranks with ONE NCCL
all-reduce (SUM) over a flat fp32 buffer plus one small MAX for the radii, and applies Adam
on every replica. What DGE does per view and that does not depend on the view is hoisted:

  * activations (exp / sigmoid / normalize / cat of gaussiansplatting/scene/gaussian_model.py:
    221-258) run once per step; the views back-propagate into detached leaves and one
    backward through the activations follows (DGE re-runs them for every render()).
  * the screen-space gradient tap `means2D` (gaussian_renderer/__init__.py:60-69) is one
    tensor shared by the step's views, so its .grad IS the per-step sum DGE builds in
    on_before_optimizer_step (threestudio/systems/DGE.py:269-276).
  * the six parameter tensors, their .grad and the screen-space gradient are views into flat
    buffers: autograd accumulates straight into the buffer NCCL reduces in place, and the
    fused Adam (csrc/apply_weights.cu: fused_adam_kernel) sweeps contiguous slices.
"""
import math
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib as L
from . import diff_gaussian_rasterization as dgr
from . import scene

# name, floats per Gaussian, shape tail — gaussian_model.py:341-372 (param groups, same order)
GROUPS = [("xyz", 3, (3,)), ("f_dc", 3, (1, 3)), ("f_rest", 45, (15, 3)), ("opacity", 1, (1,)),
          ("scaling", 3, (3,)), ("rotation", 4, (4,))]
FLOATS_PER_GAUSSIAN = sum(g[1] for g in GROUPS)  # 59
# Order of the groups INSIDE the flat parameter / gradient / Adam buffers: the features first, then everything a
# projection needs, so that "geometry gradients + screen-space gradient + loss" is ONE contiguous all-reduce (the
# part the next step's front half waits for) and "f_dc + f_rest" another (three quarters of the bytes, which that
# front half runs under).
LAYOUT = ("f_dc", "f_rest", "xyz", "opacity", "scaling", "rotation")
GEOMETRY_GROUPS = ("xyz", "opacity", "scaling", "rotation")
# gaussiansplatting/arguments/__init__.py:72-81 (OptimizationParams: position_lr_init, feature_lr, opacity_lr,
# scaling_lr, rotation_lr; f_rest at feature_lr / 20, gaussian_model.py:344); configs/dge.yaml scales them by 1
DEFAULT_LRS = {"xyz": 0.00016, "f_dc": 0.0125, "f_rest": 0.0125 / 20.0, "opacity": 0.05, "scaling": 0.005,
               "rotation": 0.001}
MASKED_GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling")  # gaussian_model.py:848 (no rotation)


def inverse_sigmoid(x):
    return torch.log(x / (1 - x))


def _build_rotation(r):
    """gaussiansplatting/utils/general_utils.py:78-100 (rotation matrices of un-normalised quaternions)."""
    norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    R = torch.zeros((q.size(0), 3, 3), device=r.device)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - r * z)
    R[:, 0, 2] = 2 * (x * z + r * y)
    R[:, 1, 0] = 2 * (x * y + r * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - r * x)
    R[:, 2, 0] = 2 * (x * z - r * y)
    R[:, 2, 1] = 2 * (y * z + r * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


class FitModel:
    """Flat-buffer replica of GaussianModel's optimisable state."""

    def __init__(self, gaussians: scene.Gaussians, device, lrs=None, fused_adam=True, sh_degree=3):
        self.device, self.sh_degree = device, sh_degree
        self.lrs = dict(DEFAULT_LRS if lrs is None else lrs)
        self.fused_adam = fused_adam
        self.step_count = 0
        raw = {
            "xyz": gaussians.means3D, "f_dc": gaussians.shs[:, :1, :], "f_rest": gaussians.shs[:, 1:, :],
            "opacity": inverse_sigmoid(gaussians.opacities), "scaling": torch.log(gaussians.scales),
            "rotation": gaussians.rotations,
        }
        self._allocate(raw)
        P = self.P
        # densification statistics (gaussian_model.py:338-339, 811-815; DGE.py:277-284)
        self.xyz_gradient_accum = torch.zeros(P, 1, device=device)
        self.denom = torch.zeros(P, 1, device=device)
        self.max_radii2D = torch.zeros(P, dtype=torch.int32, device=device)
        self.grad_mask: Optional[torch.Tensor] = None  # uint8 [P], local editing

    @staticmethod
    def _layout(P):
        """Offsets of the groups inside the flat buffers for P Gaussians: (total floats, {name: slice}). Every
        group starts on a 16-byte boundary (float4 accesses in the kernels: rotation rows, vectorised Adam),
        whatever P is; order as LAYOUT."""
        pad4 = lambda x: (x + 3) // 4 * 4
        off, slices = 0, {}
        for name, k, _ in sorted(GROUPS, key=lambda g: LAYOUT.index(g[0])):
            slices[name] = slice(off, off + k * P)
            off += pad4(k * P)
        return off, slices

    def _allocate(self, raw, exp_avg=None, exp_avg_sq=None):
        """(Re)builds the flat parameter / gradient / Adam-state buffers for the raw parameter tensors
        `raw` (name -> [P, ...]); exp_avg / exp_avg_sq: optional per-group Adam state to carry over."""
        device = self.device
        P = raw["xyz"].shape[0]
        n, slices = self._layout(P)
        flat, m, v = (torch.zeros(n, dtype=torch.float32, device=device) for _ in range(3))
        for name, k, tail in GROUPS:
            sl = slices[name]
            flat[sl].view(P, *tail).copy_(raw[name].detach().to(device).reshape(P, *tail))
            if exp_avg is not None:
                m[sl].copy_(exp_avg[name].reshape(-1))
                v[sl].copy_(exp_avg_sq[name].reshape(-1))
        self._bind(P, flat, m, v, carry_state=exp_avg is not None)

    def _bind(self, P, flat, exp_avg, exp_avg_sq, carry_state=False):
        """Makes `flat` / `exp_avg` / `exp_avg_sq` (laid out by _layout(P)) the model's state: parameter leaves
        and their gradients as views, a fresh gradient buffer, caches of the old buffers dropped."""
        device = self.device
        pad4 = lambda x: (x + 3) // 4 * 4
        n, self.slices = self._layout(P)
        self.P, self.flat, self.exp_avg, self.exp_avg_sq = P, flat, exp_avg, exp_avg_sq
        # gradient buffer: 59P parameter grads, the 3P screen-space gradient sum, and (last 4 floats) the step's
        # loss, which rides in the same all-reduce
        self.flat_grad = torch.zeros(n + pad4(3 * P) + 4, dtype=torch.float32, device=device)
        self.params = {}
        for name, k, tail in GROUPS:
            sl = self.slices[name]
            p = self.flat[sl].view(P, *tail)
            p.requires_grad_(True)  # a leaf: `flat` itself never requires grad
            p.grad = self.flat_grad[sl].view(P, *tail)
            self.params[name] = p
        self.means2D = torch.zeros(P, 3, dtype=torch.float32, device=device, requires_grad=True)
        self.means2D_slice = slice(n, n + 3 * P)
        self.means2D.grad = self.flat_grad[self.means2D_slice].view(P, 3)
        self.loss_slot = self.flat_grad[n + pad4(3 * P):n + pad4(3 * P) + 1]
        # [early_slice]: the geometry groups, what the next step's projection waits for; [stats_slice]: screen-space
        # gradient + loss (statistics and the returned loss only); [late_slice]: the features
        self.late_slice = slice(self.slices["f_dc"].start, self.slices["f_rest"].stop)
        self.early_slice = slice(self.slices["xyz"].start, n)
        self.stats_slice = slice(n, n + pad4(3 * P) + 4)
        # per-P caches of the fit step (lanes, batches, activations) belong to the old buffers
        self._geom_version = getattr(self, "_geom_version", 0) + 1
        for attr in ("_lane_key", "_batch_key", "_acts", "_lanes", "_batches", "_acc", "_flags", "_lane_acc",
                     "_lane_flags", "_min_chunks", "_front"):
            if hasattr(self, attr):
                delattr(self, attr)
        if not self.fused_adam:
            # the reference's optimiser, verbatim (gaussian_model.py:374)
            groups = [{"params": [self.params[nm]], "lr": self.lrs[nm], "name": nm} for nm, _, _ in GROUPS]
            self.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
            if carry_state:
                for nm, _, _ in GROUPS:
                    sl, pp = self.slices[nm], self.params[nm]
                    self.optimizer.state[pp] = {"step": torch.tensor(float(self.step_count)),
                                                "exp_avg": self.exp_avg[sl].view_as(pp),
                                                "exp_avg_sq": self.exp_avg_sq[sl].view_as(pp)}

    @classmethod
    def from_raw(cls, raw, device, **kw):
        """A model from the six RAW parameter tensors by group name (e.g. ply.load_ply)."""
        P = raw["xyz"].shape[0]
        shs = torch.cat((raw["f_dc"], raw["f_rest"]), dim=1)
        model = cls(scene.Gaussians(raw["xyz"], torch.exp(raw["scaling"]), raw["rotation"], torch.sigmoid(raw["opacity"]),
                                    shs), device, **kw)
        model._allocate({k: v.reshape(P, *[g[2] for g in GROUPS if g[0] == k][0]) for k, v in raw.items()})
        return model

    def save_ply(self, path):
        """GaussianModel.save_ply's file (gaussian_model.py:410-445): see dge_b200/ply.py."""
        from . import ply
        ply.save_ply(self.params, path)

    def state_dict(self):
        """Everything a resumed fit needs (the reference checkpoints through Lightning + save_ply,
        threestudio/systems/DGE.py:497; a PLY alone loses the optimiser): raw parameters, Adam moments and
        step, learning rates and the xyz schedule, densification statistics, edit mask. CPU tensors."""
        cpu = lambda t: t.detach().to("cpu").clone()
        moments = {nm: tuple(cpu(m) for m in self.adam_state(nm)) for nm, _, _ in GROUPS}
        return {
            "params": {nm: cpu(p) for nm, p in self.params.items()},
            "exp_avg": {nm: m[0] for nm, m in moments.items()}, "exp_avg_sq": {nm: m[1] for nm, m in moments.items()},
            "step_count": self.step_count, "lrs": dict(self.lrs), "xyz_schedule": getattr(self, "_xyz_schedule", None),
            "sh_degree": self.sh_degree, "xyz_gradient_accum": cpu(self.xyz_gradient_accum), "denom": cpu(self.denom),
            "max_radii2D": cpu(self.max_radii2D), "grad_mask": None if self.grad_mask is None else cpu(self.grad_mask),
        }

    def load_state_dict(self, sd):
        """Inverse of state_dict(); the number of Gaussians may differ from the current one (densification)."""
        self.step_count, self.lrs, self.sh_degree = int(sd["step_count"]), dict(sd["lrs"]), int(sd["sh_degree"])
        if sd.get("xyz_schedule") is not None:
            self._xyz_schedule = tuple(sd["xyz_schedule"])
        self._allocate(sd["params"], sd["exp_avg"], sd["exp_avg_sq"])
        dev = self.device
        self.xyz_gradient_accum = sd["xyz_gradient_accum"].to(dev).clone()
        self.denom = sd["denom"].to(dev).clone()
        self.max_radii2D = sd["max_radii2D"].to(dev).clone()
        self.set_grad_mask(sd.get("grad_mask"))

    def parameters_changed(self):
        """To be called after editing geometry parameters in place from outside (anything but adam_step,
        densify_and_prune and load_state_dict, which do it themselves): a front half prefetched for the next
        step (fit_step's next_cameras) belongs to the old parameters and must not be used."""
        self._geom_version = getattr(self, "_geom_version", 0) + 1

    def adam_state(self, name):
        """(exp_avg, exp_avg_sq) of a parameter group, shaped like the parameter."""
        if not self.fused_adam and self.params[name] in self.optimizer.state:
            st = self.optimizer.state[self.params[name]]
            return st["exp_avg"], st["exp_avg_sq"]
        sl, pp = self.slices[name], self.params[name]
        return self.exp_avg[sl].view_as(pp), self.exp_avg_sq[sl].view_as(pp)

    # -- gaussian_model.py:543-807
    @torch.no_grad()
    def densify_and_prune(self, max_grad, max_densify_percent, min_opacity, extent, max_screen_size,
                          percent_dense=0.01, N=2, generator=None, normal_samples=None, device_kernels=None):
        """GaussianModel.densify_and_prune (gaussiansplatting/scene/gaussian_model.py:770-797) on the
        flat-buffer replica: clone small Gaussians with a large accumulated screen-space gradient
        (:728-768), split large ones into N samples (:675-726), prune transparent / oversized ones
        (:786-796), carrying the Adam state the way cat_tensors_to_optimizer / _prune_optimizer do
        (:543-640: zeros for new rows, rows dropped with their Gaussians) and resetting the statistics as
        densification_postfix does (:664-666). Only Gaussians inside the edit mask (set_grad_mask)
        are touched, as in the reference (:774, :795). The flat buffers are rebuilt for the new count.
        Multi-GPU replicas stay identical when every rank passes a generator seeded alike (the
        reference draws from the global CUDA generator, :685-687). normal_samples: the N(0,1)-scaled
        draw to use instead (tests). Returns (P_before, P_after_clone, P_after_split, P_after_prune).
        On a CUDA device with the fused optimiser (device_kernels, default) the decisions, the compaction, the
        clone / split appends and the Adam-state surgery are TWO kernels over the flat buffers
        (csrc/densify.cu: dge_densify_select / dge_densify_gather) instead of the boolean-mask gathers and
        torch.cat re-allocations below; same rows in the same order (tests/test_fit_gpu.py)."""
        dev = self.device
        P0 = self.P
        mask = torch.ones(P0, dtype=torch.bool, device=dev) if self.grad_mask is None else self.grad_mask.bool()
        grads = self.xyz_gradient_accum / self.denom
        grads[grads.isnan()] = 0.0
        grads[~mask] = 0.0
        if max_densify_percent < 1:
            valid_percent = len(grads.nonzero()) * max_densify_percent / grads.shape[0]
            thresold_value = torch.quantile(grads, 1 - valid_percent)
            grads[grads < thresold_value] = 0.0
        if device_kernels is None:
            device_kernels = dev.type == "cuda" and self.fused_adam
        if device_kernels:
            return self._densify_on_device(grads, max_grad, min_opacity, extent, max_screen_size, percent_dense, N,
                                           generator, normal_samples)
        raw = {k: v.detach() for k, v in self.params.items()}
        m = {k: self.adam_state(k)[0].detach().clone() for k in raw}
        v = {k: self.adam_state(k)[1].detach().clone() for k in raw}

        def extend(new):  # cat_tensors_to_optimizer (:609-640)
            for k in raw:
                raw[k] = torch.cat((raw[k], new[k]), dim=0)
                m[k] = torch.cat((m[k], torch.zeros_like(new[k])), dim=0)
                v[k] = torch.cat((v[k], torch.zeros_like(new[k])), dim=0)

        def keep(valid):  # _prune_optimizer (:568-587)
            for k in raw:
                raw[k], m[k], v[k] = raw[k][valid], m[k][valid], v[k][valid]

        # ---- clone (:728-768)
        sel = torch.norm(grads, dim=-1) >= max_grad
        sel = torch.logical_and(sel, torch.max(torch.exp(raw["scaling"]), dim=1).values <= percent_dense * extent)
        extend({k: raw[k][sel] for k in raw})
        mask = torch.cat([mask, mask[sel]], dim=0)
        P1 = raw["xyz"].shape[0]
        # ---- split (:675-726)
        padded_grad = torch.zeros(P1, device=dev)
        padded_grad[:grads.shape[0]] = grads.squeeze()
        sel = padded_grad >= max_grad
        scaling = torch.exp(raw["scaling"])
        sel = torch.logical_and(sel, torch.max(scaling, dim=1).values > percent_dense * extent)
        stds = scaling[sel].repeat(N, 1)
        if normal_samples is not None:
            samples = normal_samples.to(dev)
        else:
            samples = torch.normal(mean=torch.zeros((stds.size(0), 3), device=dev), std=stds, generator=generator)
        rots = _build_rotation(raw["rotation"][sel]).repeat(N, 1, 1)
        new = {
            "xyz": torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + raw["xyz"][sel].repeat(N, 1),
            "scaling": torch.log(scaling[sel].repeat(N, 1) / (0.8 * N)),
            "rotation": raw["rotation"][sel].repeat(N, 1),
            "f_dc": raw["f_dc"][sel].repeat(N, 1, 1),
            "f_rest": raw["f_rest"][sel].repeat(N, 1, 1),
            "opacity": raw["opacity"][sel].repeat(N, 1),
        }
        extend(new)
        mask = torch.cat([mask] + [mask[sel]] * N, dim=0)
        P2_all = raw["xyz"].shape[0]
        prune_filter = torch.cat((sel, torch.zeros(N * int(sel.sum()), device=dev, dtype=torch.bool)))
        keep(~prune_filter)
        mask = mask[~prune_filter]
        P2 = raw["xyz"].shape[0]
        # densification_postfix (:664-666) reset the statistics for the new count
        max_radii2D = torch.zeros(P2, dtype=torch.int32, device=dev)
        # ---- prune (:786-796)
        prune_mask = (torch.sigmoid(raw["opacity"]) < min_opacity).squeeze()
        if max_screen_size:
            big_points_vs = max_radii2D > max_screen_size
            big_points_ws = torch.exp(raw["scaling"]).max(dim=1).values > 0.1 * extent
            prune_mask = torch.logical_or(torch.logical_or(prune_mask, big_points_vs), big_points_ws)
        prune_mask = torch.logical_and(prune_mask, mask)
        keep(~prune_mask)
        mask = mask[~prune_mask]
        P3 = raw["xyz"].shape[0]
        had_mask = self.grad_mask is not None
        self._allocate(raw, m, v)
        self.xyz_gradient_accum = torch.zeros(P3, 1, device=dev)
        self.denom = torch.zeros(P3, 1, device=dev)
        self.max_radii2D = torch.zeros(P3, dtype=torch.int32, device=dev)
        self.grad_mask = mask.to(torch.uint8).contiguous() if had_mask else None
        del P2_all
        return P0, P1, P2, P3

    @torch.no_grad()
    def _densify_on_device(self, grads, max_grad, min_opacity, extent, max_screen_size, percent_dense, N, generator,
                           normal_samples):
        """densify_and_prune's data movement as two kernels (csrc/densify.cu). `grads`: the masked / thresholded
        accumulated gradients [P,1] the torch path computes (gaussian_model.py:770-781)."""
        lib, dev, P0 = L.load(), self.device, self.P
        st = L.stream_ptr(dev)
        g = grads.reshape(-1).to(torch.float32).contiguous()
        keep = torch.empty(4, P0, dtype=torch.int32, device=dev)
        sel = torch.empty(P0, dtype=torch.uint8, device=dev)
        mask = self.grad_mask
        prune_size = float(0.1 * extent) if max_screen_size else -1.0
        L.check(lib.dge_densify_select(P0, g.data_ptr(), self.params["scaling"].data_ptr(), self.params["opacity"].data_ptr(),
                                       None if mask is None else mask.data_ptr(), float(max_grad),
                                       float(percent_dense * extent), float(min_opacity), prune_size, int(N),
                                       keep.data_ptr(), sel.data_ptr(), st), "densify select")
        incl = torch.cumsum(keep, dim=1, dtype=torch.int32)
        scan = (incl - keep).contiguous()
        K_orig, K_clone, K_child, K_split = (int(x) for x in incl[:, -1].tolist())
        n_clone = int((sel & 1).sum())
        if normal_samples is not None:
            samples = normal_samples.to(dev, torch.float32).contiguous()
        else:  # the draw of densify_and_split (:685-687): N(0, scale) per split-selected Gaussian and copy
            stds = torch.exp(self.params["scaling"].detach())[(sel & 2) != 0].repeat(N, 1)
            samples = torch.normal(mean=torch.zeros((stds.size(0), 3), device=dev), std=stds, generator=generator)
        assert samples.shape[0] == N * K_split
        P3 = K_orig + K_clone + N * K_child
        n, slices = self._layout(P3)
        new = [torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(3)]
        names = [nm for nm, _, _ in GROUPS]
        widths = (L.C.c_int * 6)(*[k for _, k, _ in GROUPS])
        old = [self.flat, self.exp_avg, self.exp_avg_sq]
        src = (L.C.c_void_p * 18)(*[old[a][self.slices[nm]].data_ptr() for a in range(3) for nm in names])
        dst = (L.C.c_void_p * 18)(*[new[a][slices[nm]].data_ptr() if P3 else 0 for a in range(3) for nm in names])
        had_mask = mask is not None
        mask_out = torch.empty(P3, dtype=torch.uint8, device=dev) if had_mask else None
        if P3:
            L.check(lib.dge_densify_gather(P0, int(N), keep.data_ptr(), scan.data_ptr(), K_orig, K_clone, K_child, K_split,
                                           src, dst, widths, names.index("xyz"), names.index("scaling"),
                                           names.index("rotation"), samples.data_ptr() if K_split else None,
                                           None if mask is None else mask.data_ptr(),
                                           None if mask_out is None else mask_out.data_ptr(), st), "densify gather")
        self._bind(P3, new[0], new[1], new[2])
        self.xyz_gradient_accum = torch.zeros(P3, 1, device=dev)
        self.denom = torch.zeros(P3, 1, device=dev)
        self.max_radii2D = torch.zeros(P3, dtype=torch.int32, device=dev)
        self.grad_mask = mask_out if had_mask else None
        P1 = P0 + n_clone
        return P0, P1, P1 + (N - 1) * K_split, P3

    # -- gaussian_model.py:221-258
    def activations(self):
        p = self.params
        return {
            "means3D": p["xyz"],
            "shs": torch.cat((p["f_dc"], p["f_rest"]), dim=1),
            "opacities": torch.sigmoid(p["opacity"]),
            "scales": torch.exp(p["scaling"]),
            "rotations": torch.nn.functional.normalize(p["rotation"]),
        }

    def activations_fused(self, which: str = "all"):
        """The same activations from ONE kernel into persistent buffers (no autograd graph): their
        backward is applied by dge_fit_backward_geom_raw's epilogue. which = "geometry"
        (opacity / scaling / rotation only) or "features" (shs = cat(f_dc, f_rest) only) refreshes one half: the
        pipelined multi-GPU step projects the next step's Gaussians before its features have been stepped."""
        if getattr(self, "_acts", None) is None:
            f32 = dict(dtype=torch.float32, device=self.device)
            self._acts = {"shs": torch.empty(self.P, 16, 3, **f32), "opacities": torch.empty(self.P, 1, **f32),
                          "scales": torch.empty(self.P, 3, **f32), "rotations": torch.empty(self.P, 4, **f32)}
        a, p = self._acts, self.params
        geo, feat = which in ("all", "geometry"), which in ("all", "features")
        L.check(L.load().dge_fit_activate(self.P, p["f_dc"].data_ptr() if feat else None,
                                          p["f_rest"].data_ptr() if feat else None,
                                          p["opacity"].data_ptr() if geo else None, p["scaling"].data_ptr(),
                                          p["rotation"].data_ptr(), a["shs"].data_ptr(), a["opacities"].data_ptr(),
                                          a["scales"].data_ptr(), a["rotations"].data_ptr(), L.stream_ptr(self.device)),
                "activate")
        return {"means3D": p["xyz"].detach(), **a}

    def training_setup(self, position_lr_max_steps, position_lr_init=0.00016, position_lr_final=0.000016,
                       position_lr_delay_mult=0.01, spatial_lr_scale=1.0, feature_lr=0.0125, opacity_lr=0.05,
                       scaling_lr=0.005, rotation_lr=0.001):
        """The learning rates of GaussianModel.training_setup (gaussian_model.py:341-380) with the defaults of
        OptimizationParams (gaussiansplatting/arguments/__init__.py:72-81; DGE passes trainer.max_steps and the
        camera extent as spatial_lr_scale, DGE.py:503-515): xyz scaled by the scene extent and scheduled by
        update_learning_rate, f_rest at feature_lr / 20. Adam state is kept."""
        self.lrs = {"xyz": position_lr_init * spatial_lr_scale, "f_dc": feature_lr, "f_rest": feature_lr / 20.0,
                    "opacity": opacity_lr, "scaling": scaling_lr, "rotation": rotation_lr}
        self._xyz_schedule = (position_lr_init * spatial_lr_scale, position_lr_final * spatial_lr_scale,
                              position_lr_delay_mult, position_lr_max_steps)
        self._sync_optimizer_lrs()

    def update_learning_rate(self, iteration: int) -> float:
        """GaussianModel.update_learning_rate (gaussian_model.py:382-388, called every step from
        DGE.training_step, DGE.py:621): log-linear interpolation of the xyz rate from lr_init at step 0 to
        lr_final at max_steps (utils/general_utils.py:29-62 with lr_delay_steps = 0, as training_setup
        calls it); the other groups keep their rates. Returns the new xyz rate."""
        lr_init, lr_final, _delay_mult, max_steps = getattr(self, "_xyz_schedule", (self.lrs["xyz"],) * 2 + (1.0, 1))
        if iteration < 0 or (lr_init == 0.0 and lr_final == 0.0):
            lr = 0.0
        else:
            t = min(max(iteration / max_steps, 0.0), 1.0)
            lr = math.exp(math.log(lr_init) * (1.0 - t) + math.log(lr_final) * t)
        self.lrs["xyz"] = lr
        self._sync_optimizer_lrs()
        return lr

    def _sync_optimizer_lrs(self):
        if not self.fused_adam:  # the fused path reads self.lrs at every step
            for group in self.optimizer.param_groups:
                group["lr"] = self.lrs[group["name"]]

    def set_grad_mask(self, mask: Optional[torch.Tensor]):
        self.grad_mask = None if mask is None else mask.to(self.device).to(torch.uint8).contiguous()

    def zero_grad(self):
        self.flat_grad.zero_()

    def adam_step(self, only=None, skip=(), advance=True, rows=None):
        """One Adam step over the parameter groups (`only` / `skip`: a subset of them, so that the
        groups whose all-reduce has landed can be stepped first; advance=False: same step number;
        rows=(r0, r1): only these Gaussians of the selected groups, r0 a multiple of 4; fused path only)."""
        if advance:
            self.step_count += 1
        stepped = [nm for nm, _, _ in GROUPS if nm not in skip and (only is None or nm in only)]
        if any(nm in GEOMETRY_GROUPS for nm in stepped):
            self._geom_version = getattr(self, "_geom_version", 0) + 1  # a prefetched front half is stale now
        if not self.fused_adam:
            if self.grad_mask is not None:  # the reference's hooks (gaussian_model.py:837-856)
                m = self.grad_mask.to(torch.float32)
                for nm in MASKED_GROUPS:
                    g = self.params[nm].grad
                    g.mul_(m.view(-1, *([1] * (g.ndim - 1))))
            self.optimizer.step()
            return
        lib = L.load()
        st = L.stream_ptr(self.device)
        for name, k, _ in GROUPS:
            if name in skip or (only is not None and name not in only):
                continue
            sl = self.slices[name]
            mask = self.grad_mask if (self.grad_mask is not None and name in MASKED_GROUPS) else None
            r0, r1 = rows if rows is not None else (0, self.P)
            if r1 <= r0:
                continue
            off = 4 * k * r0  # bytes; r0 % 4 == 0 keeps the float4 path aligned
            L.check(lib.dge_fused_adam(self.flat[sl].data_ptr() + off, self.flat_grad[sl].data_ptr() + off,
                                       self.exp_avg[sl].data_ptr() + off, self.exp_avg_sq[sl].data_ptr() + off,
                                       k * (r1 - r0), self.lrs[name], 0.9, 0.999, 1e-15, self.step_count,
                                       None if mask is None else mask.data_ptr() + r0, k, st),
                    "fused adam")


CAM_FLOATS = 40  # view[16] | proj[16] | campos[3] | tan_fovx | tan_fovy | pad[3]  (include/dge_b200.h "fit step")


_CAM_RECORDS = {}


def camera_record(cam: scene.Camera) -> torch.Tensor:
    """pack_camera, memoised per camera object (cameras are reused from step to step; packing one costs
    three device-to-host copies when its matrices live on the GPU). The cache holds a reference to the
    camera's tensors — so their storage cannot be freed and handed to ANOTHER camera while the entry
    lives (a key made of data pointers alone went stale exactly that way) — and checks their version
    counters, so in-place updates are seen."""
    wv, fp, cc = cam.world_view_transform, cam.full_proj_transform, cam.camera_center
    key = (id(wv), id(fp), id(cc))
    ver = (wv._version, fp._version, cc._version, cam.FoVx, cam.FoVy)
    hit = _CAM_RECORDS.get(key)
    if hit is not None and hit[0] == ver:
        return hit[2]
    if len(_CAM_RECORDS) > 4096:
        _CAM_RECORDS.clear()
    rec = pack_camera(cam)
    _CAM_RECORDS[key] = (ver, (wv, fp, cc), rec)
    return rec


def pack_camera(cam: scene.Camera) -> torch.Tensor:
    """The camera as the pinned 40-float record the fit entry points take."""
    rec = torch.zeros(CAM_FLOATS, dtype=torch.float32)
    rec[0:16] = cam.world_view_transform.detach().cpu().reshape(-1)
    rec[16:32] = cam.full_proj_transform.detach().cpu().reshape(-1)
    rec[32:35] = cam.camera_center.detach().cpu().reshape(-1)
    rec[35], rec[36] = math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5)
    return rec.pin_memory() if torch.cuda.is_available() else rec


class ViewLane:
    """Everything one CUDA stream needs to push views through the C-ABI without touching the
    allocator, autograd or Python object construction per view: fixed output / scratch buffers
    (geom and image blobs have a fixed size for a given P, W, H; the binning blob only grows) and
    allocator callbacks created once."""

    def __init__(self, model: "FitModel", W: int, H: int, stream):
        lib = L.load()
        dev, P = model.device, model.P
        self.stream, self.W, self.H, self.P = stream, W, H, P
        f32 = dict(dtype=torch.float32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        self.color = torch.empty(3, H, W, **f32)
        self.depth = torch.empty(1, H, W, **f32)
        self.dL = torch.empty(3, H, W, **f32)
        self.radii = torch.empty(P, dtype=torch.int32, device=dev)
        self.radii_max = torch.zeros(P, dtype=torch.int32, device=dev)
        self.loss = torch.zeros((), **f32)
        self.geom = torch.empty(lib.dge_geom_bytes(P), **u8)
        self.img = torch.empty(lib.dge_image_bytes(W, H), **u8)
        self.binning = torch.empty(0, **u8)
        self.target_buf = torch.empty(3, H, W, **f32)  # staged target image (host_inputs)

        def fixed(t):
            ptr = t.data_ptr()
            return L.ALLOC_FN(lambda _ctx, nbytes: ptr if nbytes <= t.numel() else 0)

        def growing(_ctx, nbytes):
            if nbytes > self.binning.numel():
                with torch.cuda.stream(self.stream):
                    self.binning = torch.empty(int(nbytes * 1.25) + 256, **u8)
            return self.binning.data_ptr()

        self.cb_geom, self.cb_img = fixed(self.geom), fixed(self.img)
        self.cb_binning = L.ALLOC_FN(growing)
        self.stream_ptr = L.C.c_void_p(stream.cuda_stream)

    def reset(self):
        self.loss.zero_()
        self.radii_max.zero_()


def _background_is_black(model, bg) -> int:
    """1 when the background is (0,0,0): the blend backward then runs without the background term of
    dL/dalpha (dge_fit_*_backward_blend's promise). The answer is cached per background TENSOR OBJECT and
    version counter — never per data pointer: the caching allocator hands a freed tensor's address to the
    next one, and an in-place update keeps it."""
    hit = getattr(model, "_bg_seen", None)
    if hit is None or hit[0] is not bg or hit[1] != bg._version:
        model._bg_seen = (bg, bg._version, int(not bool(bg.detach().cpu().any())))
    return model._bg_seen[2]


def _direct_views(model: FitModel, acts, cameras, targets, bg, scale, host_inputs, num_streams):
    """The step's views through the C-ABI, round-robin over `num_streams` lanes (no autograd):
    per view forward -> fused L1 loss+gradient -> blend backward into that view's acc rows; then ONE
    batched per-Gaussian backward over all views. Returns (grads w.r.t. the activated tensors and the
    screen-space tap, loss, max radii). Leaves the step's gradients in model.flat_grad."""
    lib = L.load()
    dev = model.device
    H, W = cameras[0].image_height, cameras[0].image_width
    V = len(cameras)
    if V > 64:
        raise ValueError("a step holds at most 64 views per rank (dge_fit_backward_geom_raw): shard the batch "
                         "over more ranks or take two steps")
    S = max(1, min(num_streams, V))
    P = model.P
    main = torch.cuda.current_stream(dev)
    key = (W, H, S, V)
    if getattr(model, "_lane_key", None) != key:
        model._lanes = [ViewLane(model, W, H, torch.cuda.Stream(dev)) for _ in range(S)]
        f32 = dict(dtype=torch.float32, device=dev)
        model._lane_acc = torch.empty(V, P, 12, **f32)          # blend-stage sums of every view of the step
        model._lane_flags = torch.empty(V, P, dtype=torch.uint8, device=dev)  # visible / SH-clamp bits
        model._lane_cams = torch.empty(V, CAM_FLOATS, **f32)
        model._lane_key = key
    lanes, acc, flags, cams_dev = model._lanes, model._lane_acc, model._lane_flags, model._lane_cams
    a = {k: v.detach() for k, v in acts.items()}
    ptrs = {k: v.data_ptr() for k, v in a.items()}
    M = a["shs"].shape[1]
    for ln in lanes:
        ln.stream.wait_stream(main)
        with torch.cuda.stream(ln.stream):
            ln.reset()
    bgp = bg.data_ptr()
    bg_black = _background_is_black(model, bg)
    n_img = 3 * H * W
    for i, (cam, target) in enumerate(zip(cameras, targets)):
        ln = lanes[i % S]
        rec = camera_record(cam)
        tfx, tfy = float(rec[35]), float(rec[36])
        with torch.cuda.stream(ln.stream):
            cams_dev[i].copy_(rec, non_blocking=True)
            if host_inputs:  # pinned host -> this lane's staging buffer, asynchronously on its stream
                ln.target_buf.copy_(target, non_blocking=True)
                target = ln.target_buf
            cam_ptr, acc_ptr = cams_dev[i].data_ptr(), acc[i].data_ptr()
            R = lib.dge_fit_forward(
                ln.cb_geom, ln.cb_binning, ln.cb_img, None, P, model.sh_degree, M, bgp, W, H, ptrs["means3D"],
                ptrs["shs"], ptrs["opacities"], ptrs["scales"], 1.0, ptrs["rotations"], cam_ptr, tfx,
                tfy, ln.color.data_ptr(), ln.depth.data_ptr(), ln.radii.data_ptr(), acc_ptr, flags[i].data_ptr(),
                ln.stream_ptr)
            L.check(R, "fit forward")
            L.check(lib.dge_l1_loss_grad(ln.color.data_ptr(), target.data_ptr(), n_img, scale, ln.dL.data_ptr(),
                                         ln.loss.data_ptr(), ln.stream_ptr), "l1 loss")
            L.check(lib.dge_fit_backward_blend(P, R, bgp, bg_black, W, H, ln.geom.data_ptr(), ln.binning.data_ptr(),
                                               ln.img.data_ptr(), ln.dL.data_ptr(), acc_ptr, ln.stream_ptr),
                    "fit backward blend")
            torch.maximum(ln.radii_max, ln.radii, out=ln.radii_max)
    for ln in lanes:
        main.wait_stream(ln.stream)
    # ONE per-Gaussian backward for all views, activations' backward in its epilogue: gradients
    # w.r.t. the raw parameters go straight into the flat buffer the collective reduces (every
    # row is written, so the buffer needs no zero fill either)
    gp = {k: v.grad.data_ptr() for k, v in model.params.items()}
    L.check(lib.dge_fit_backward_geom_raw(
        P, model.sh_degree, V, cams_dev.data_ptr(), W, H, 1.0, acc.data_ptr(), P * 12, flags.data_ptr(), P,
        ptrs["means3D"], ptrs["shs"],
        ptrs["opacities"], ptrs["scales"], ptrs["rotations"], model.params["rotation"].data_ptr(), gp["xyz"],
        model.means2D.grad.data_ptr(), gp["f_dc"], gp["f_rest"], gp["opacity"], gp["scaling"], gp["rotation"],
        L.stream_ptr(dev)), "fit backward geom")
    loss, radii_max = lanes[0].loss.clone(), lanes[0].radii_max
    for ln in lanes[1:]:
        loss += ln.loss
        radii_max = torch.maximum(radii_max, ln.radii_max)
    return loss, radii_max


class ViewBatch:
    """Buffers for pushing a CHUNK of the step's views through the batched C-ABI entry points
    (dge_fit_views_forward / _backward_blend): one launch per stage with the view as a grid dimension.
    Geometry / image blobs hold V views at a uniform stride, the binning arena only grows. Each chunk
    has its own stream; `acc` and `cams` are this chunk's rows of the step-wide buffers."""

    def __init__(self, model: "FitModel", W: int, H: int, V: int, acc, flags, cams, cams_host):
        lib = L.load()
        dev, P = model.device, model.P
        self.W, self.H, self.V, self.P = W, H, V, P
        f32 = dict(dtype=torch.float32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        self.stream = torch.cuda.Stream(dev)
        self.color = torch.empty(V, 3, H, W, **f32)
        self.depth = torch.empty(V, 1, H, W, **f32)
        self.dL = torch.empty(V, 3, H, W, **f32)
        self.targets = torch.empty(V, 3, H, W, **f32)       # staging for host / listed targets
        self.radii_max = torch.zeros(P, dtype=torch.int32, device=dev)
        self.loss = torch.zeros((), **f32)
        self.acc, self.flags, self.cams, self.cams_host = acc, flags, cams, cams_host
        self.geom = torch.empty((lib.dge_geom_bytes(P) + 255) // 256 * 256 * V, **u8)
        self.img = torch.empty((lib.dge_image_bytes(W, H) + 255) // 256 * 256 * V, **u8)
        self.binning = torch.empty(0, **u8)
        self.num_rendered = (L.C.c_int * V)()
        self.R = 0
        self.copy_stream = torch.cuda.Stream(dev)
        self.copied = torch.cuda.Event()
        self.stream_ptr = L.C.c_void_p(self.stream.cuda_stream)

        def fixed(t):
            ptr = t.data_ptr()
            return L.ALLOC_FN(lambda _ctx, nbytes: ptr if nbytes <= t.numel() else 0)

        def growing(_ctx, nbytes):
            if nbytes > self.binning.numel():
                with torch.cuda.stream(self.stream):
                    self.binning = torch.empty(int(nbytes * 1.25) + 256, **u8)
            return self.binning.data_ptr()

        self.cb_geom, self.cb_img = fixed(self.geom), fixed(self.img)
        self.cb_binning = L.ALLOC_FN(growing)


def _batched_front(model: FitModel, acts, cameras, num_chunks=1, prune_lists=True, with_colour=True):
    """Front half of a step through the batched C-ABI: ONE launch per stage for all views of a chunk — preprocess
    that reads the Gaussians once for all its cameras, segmented depth sort, binning (dge_fit_views_front). With
    num_chunks > 1 the views are split into that many chunks, each on its own stream, so one chunk's
    bandwidth-bound stages overlap another's issue-bound blends. with_colour=False leaves the SH colours out
    (dge_fit_views_colour fills them in later: _batched_colour). Records in model._front what was projected."""
    lib = L.load()
    dev = model.device
    H, W = cameras[0].image_height, cameras[0].image_width
    V, P = len(cameras), model.P
    if V > 64:
        raise ValueError("a step holds at most 64 views per rank (dge_fit_backward_geom_raw)")
    C = max(1, min(num_chunks, V))
    key = (W, H, V, C)
    if getattr(model, "_batch_key", None) != key:
        f32 = dict(dtype=torch.float32, device=dev)
        model._acc = torch.empty(V, P, 12, **f32)          # blend-stage sums of every view of the step
        model._flags = torch.empty(V, P, dtype=torch.uint8, device=dev)  # visible / SH-clamp bits per (view, Gaussian)
        model._cams = torch.empty(V, CAM_FLOATS, **f32)
        model._cams_host = torch.empty(V, CAM_FLOATS, dtype=torch.float32).pin_memory()
        bounds = [V * c // C for c in range(C + 1)]
        model._chunk_bounds = bounds
        model._batches = [ViewBatch(model, W, H, bounds[c + 1] - bounds[c], model._acc[bounds[c]:bounds[c + 1]],
                                    model._flags[bounds[c]:bounds[c + 1]], model._cams[bounds[c]:bounds[c + 1]],
                                    model._cams_host[bounds[c]:bounds[c + 1]])
                          for c in range(C)]
        model._batch_key = key
        model._cams_key = None
    batches = model._batches
    main = torch.cuda.current_stream(dev)
    a = {k: v.detach() for k, v in acts.items()}
    ptrs = {k: v.data_ptr() for k, v in a.items()}
    M = a["shs"].shape[1]
    # cameras: one pinned [V,40] block, one H2D copy
    recs = [camera_record(cam) for cam in cameras]
    # records are memoised per camera: same objects <=> same cameras. The key HOLDS the records (an id()
    # of a freed record can be reused by another camera's)
    cams_key = getattr(model, "_cams_key", None)
    if cams_key is None or len(cams_key) != len(recs) or any(x is not y for x, y in zip(cams_key, recs)):
        cams_key = tuple(recs)
        for i, rec in enumerate(recs):
            model._cams_host[i].copy_(rec)
        model._cams.copy_(model._cams_host, non_blocking=True)
        model._cams_key = cams_key
    for vb in batches:
        vb.stream.wait_stream(main)
    # every chunk's front half (each call waits once for its instance counts)
    for vb in batches:
        with torch.cuda.stream(vb.stream):
            vb.R = L.check(lib.dge_fit_views_front(
                vb.cb_geom, vb.cb_binning, vb.cb_img, None, P, model.sh_degree, M, vb.V, W, H, ptrs["means3D"],
                ptrs["shs"] if with_colour else None, ptrs["opacities"], ptrs["scales"], 1.0, ptrs["rotations"],
                vb.cams.data_ptr(), vb.radii_max.data_ptr(), vb.flags.data_ptr(), P, vb.num_rendered, int(prune_lists),
                vb.stream_ptr), "fit views front")
    model._front = dict(cams=tuple(recs), key=(key, bool(prune_lists)), geom_version=model._geom_version,
                        coloured=with_colour, W=W, H=H, V=V, main=main)


def _batched_colour(model: FitModel, acts):
    """The SH colours of a front half that ran without them (after the features have been stepped)."""
    lib = L.load()
    fr, P = model._front, model.P
    for vb in model._batches:
        with torch.cuda.stream(vb.stream):
            vb.stream.wait_stream(fr["main"])  # the feature activations were refreshed on the caller's stream
            L.check(lib.dge_fit_views_colour(P, model.sh_degree, acts["shs"].shape[1], vb.V, acts["means3D"].data_ptr(),
                                             acts["shs"].data_ptr(), vb.cams.data_ptr(), vb.geom.data_ptr(),
                                             vb.flags.data_ptr(), P, vb.stream_ptr), "fit views colour")
    fr["coloured"] = True


def _batched_blend(model: FitModel, acts, bg, extra=None):
    """The forward blend of the front half in model._front (all views of a chunk per launch). `extra` [P] (DGE:
    gaussian.mask.float()): blended as a fourth channel into vb.sem, the image of DGE.forward's second,
    mask-colour render of every view (DGE.py:198-204). Leaves images / depth / scratch in model._batches and what
    the backward half needs in model._fw; the chunk streams are left running."""
    lib = L.load()
    fr, dev, P = model._front, model.device, model.P
    assert fr["coloured"], "the front half has no colours yet (_batched_colour)"
    W, H, V = fr["W"], fr["H"], fr["V"]
    a = {k: v.detach() for k, v in acts.items()}
    ptrs = {k: v.data_ptr() for k, v in a.items()}
    ex = None if extra is None else extra.detach().to(dev, torch.float32).reshape(-1).contiguous()
    bgp = bg.data_ptr()
    acc_stride = P * 12
    for vb in model._batches:
        with torch.cuda.stream(vb.stream):
            if ex is not None and getattr(vb, "sem", None) is None:
                vb.sem = torch.empty(vb.V, 3, H, W, dtype=torch.float32, device=dev)
            L.check(lib.dge_fit_views_blend(
                P, vb.V, vb.R, bgp, W, H, vb.geom.data_ptr(), vb.binning.data_ptr(), vb.img.data_ptr(),
                vb.color.data_ptr(), vb.depth.data_ptr(), vb.acc.data_ptr(), acc_stride,
                None if ex is None else ex.data_ptr(), None if ex is None else vb.sem.data_ptr(), vb.stream_ptr),
                "fit views blend")
    model._fw = dict(acts=a, ptrs=ptrs, bg=bg, bg_black=_background_is_black(model, bg), W=W, H=H, V=V, main=fr["main"])


def _front_is_valid(model: FitModel, cameras, num_chunks, prune_lists) -> bool:
    """Does model._front hold the front half of exactly this step (same cameras, chunking and pruning, geometry
    parameters untouched since)? True after the previous step prefetched it (fit_step(next_cameras=...))."""
    fr = getattr(model, "_front", None)
    if fr is None or len(cameras) == 0 or len(fr["cams"]) != len(cameras):
        return False
    V = len(cameras)
    key = ((cameras[0].image_width, cameras[0].image_height, V, max(1, min(num_chunks, V))), bool(prune_lists))
    return (fr["key"] == key and fr["geom_version"] == model._geom_version and
            all(camera_record(c) is r for c, r in zip(cameras, fr["cams"])))


def _batched_forward(model: FitModel, acts, cameras, bg, num_chunks=1, prune_lists=True, extra=None):
    """Forward half of a step: front half (with colours) + forward blend."""
    _batched_front(model, acts, cameras, num_chunks, prune_lists, with_colour=True)
    _batched_blend(model, acts, bg, extra)


def _batched_backward(model: FitModel, dL=None, geom_splits=1, on_range_done=None):
    """Backward half: the blend backward of every chunk (dL: [V,3,H,W] upstream gradient of the images, or None
    when the chunks' vb.dL already hold it, e.g. from the fused L1), then ONE batched per-Gaussian backward over
    all views with the activations' backward in its epilogue. The per-Gaussian backward may be issued as
    geom_splits launches over consecutive Gaussian ranges; on_range_done(first, count) is called after each
    (multi-GPU: the all-reduce of a finished range's gradients then overlaps the next range's kernel).
    Leaves the step's raw-parameter gradients in model.flat_grad; returns the max of the radii."""
    lib = L.load()
    fw, dev, P = model._fw, model.device, model.P
    W, H, V, main, ptrs = fw["W"], fw["H"], fw["V"], fw["main"], fw["ptrs"]
    batches, bounds = model._batches, model._chunk_bounds
    acc_stride = P * 12
    bgp = fw["bg"].data_ptr()
    for c, vb in enumerate(batches):
        with torch.cuda.stream(vb.stream):
            if dL is not None:
                vb.stream.wait_stream(main)  # dL was produced on the caller's stream
                g = dL[bounds[c]:bounds[c + 1]]
            else:
                g = vb.dL
            L.check(lib.dge_fit_views_backward_blend(P, vb.V, vb.R, bgp, fw["bg_black"], W, H, vb.geom.data_ptr(),
                                                     vb.binning.data_ptr(), vb.img.data_ptr(), g.data_ptr(),
                                                     vb.acc.data_ptr(), acc_stride, vb.stream_ptr),
                    "fit views backward blend")
    for vb in batches:
        main.wait_stream(vb.stream)
    gp = {k: v.grad.data_ptr() for k, v in model.params.items()}
    # ranges start at multiples of 128 Gaussians: every per-Gaussian array stays 16-byte aligned
    splits = max(1, min(geom_splits, P // 128))
    cuts = [(P * k // splits) // 128 * 128 for k in range(splits)] + [P]
    rot_raw, m2d = model.params["rotation"].data_ptr(), model.means2D.grad.data_ptr()
    with torch.cuda.stream(main):
        for first, stop in zip(cuts[:-1], cuts[1:]):
            f = 4 * first  # bytes per float column
            L.check(lib.dge_fit_backward_geom_raw(
                stop - first, model.sh_degree, V, model._cams.data_ptr(), W, H, 1.0, model._acc.data_ptr() + 12 * f,
                acc_stride, model._flags.data_ptr() + first, P, ptrs["means3D"] + 3 * f, ptrs["shs"] + 48 * f,
                ptrs["opacities"] + f, ptrs["scales"] + 3 * f, ptrs["rotations"] + 4 * f, rot_raw + 4 * f,
                gp["xyz"] + 3 * f, m2d + 3 * f, gp["f_dc"] + 3 * f, gp["f_rest"] + 45 * f, gp["opacity"] + f,
                gp["scaling"] + 3 * f, gp["rotation"] + 4 * f, L.C.c_void_p(main.cuda_stream)), "fit backward geom")
            if on_range_done is not None:
                on_range_done(first, stop - first)
    radii_max = batches[0].radii_max
    for vb in batches[1:]:
        radii_max = torch.maximum(radii_max, vb.radii_max)
    return radii_max


def _batched_views(model: FitModel, cameras, targets, bg, scale, host_inputs, num_chunks=1, prune_lists=True,
                   geom_splits=1, on_range_done=None):
    """The step's views through the batched C-ABI (front half — or the one the previous step prefetched, completed
    with its colours —, forward blend, the fused L1 loss + gradient against `targets`, _batched_backward).
    Returns (loss, max radii); leaves the step's raw-parameter gradients in model.flat_grad."""
    lib = L.load()
    if _front_is_valid(model, cameras, num_chunks, prune_lists) and not model._front["coloured"]:
        model._front["main"] = torch.cuda.current_stream(model.device)
        acts = model.activations_fused("features")  # geometry activations date from the prefetch and are still valid
        _batched_colour(model, acts)
    else:
        acts = model.activations_fused()
        _batched_front(model, acts, cameras, num_chunks, prune_lists, with_colour=True)
    _batched_blend(model, acts, bg)
    fw = model._fw
    W, H, main = fw["W"], fw["H"], fw["main"]
    batches, bounds = model._batches, model._chunk_bounds
    resident = isinstance(targets, torch.Tensor) and targets.is_cuda
    n_img = 3 * H * W
    for c, vb in enumerate(batches):
        lo, hi = bounds[c], bounds[c + 1]
        # targets: a resident [V,3,H,W] tensor is used as is; host tensors are copied on a side stream,
        # issued AFTER the forward has been queued so that the GPU is already busy while the host
        # walks through the V copy calls (they then overlap sorts / binning / the forward blend);
        # device lists are gathered once
        if resident:
            vb.tgt = targets[lo:hi]
        elif host_inputs:
            vb.copy_stream.wait_stream(main)  # the previous step's loss kernel has read vb.targets
            with torch.cuda.stream(vb.copy_stream):
                for i in range(lo, hi):
                    vb.targets[i - lo].copy_(targets[i], non_blocking=True)
                vb.copied.record(vb.copy_stream)
            vb.tgt = vb.targets
        else:
            with torch.cuda.stream(vb.stream):
                torch.stack(list(targets[lo:hi]), out=vb.targets)
            vb.tgt = vb.targets
    for vb in batches:
        with torch.cuda.stream(vb.stream):
            if host_inputs and not resident:
                vb.stream.wait_event(vb.copied)
            vb.loss.zero_()
            L.check(lib.dge_l1_loss_grad(vb.color.data_ptr(), vb.tgt.data_ptr(), vb.V * n_img, scale, vb.dL.data_ptr(),
                                         vb.loss.data_ptr(), vb.stream_ptr), "l1 loss")
    radii_max = _batched_backward(model, None, geom_splits, on_range_done)
    loss = batches[0].loss.clone()
    for vb in batches[1:]:
        loss += vb.loss
    return loss, radii_max


def render_views(means3D, shs, opacities, scales, rotations, cameras, bg, extra=None, sh_degree=3,
                 scale_modifier=1.0, prune_lists=True):
    """Forward-only rendering of many views at once (DGE.render_all_view / the camera loop of DGE.forward
    without the backward): gaussian_renderer.render() for every camera, ALL views per launch
    (dge_fit_views_forward). `extra` [P]: a per-Gaussian scalar (DGE: gaussian.mask.float()) blended as an
    additional channel; the third return value is then the image the reference gets from a SECOND
    render(..., override_color=extra repeated 3x) of each view (DGE.py:198-204), bit for bit.
    Returns (color [V,3,H,W], depth [V,1,H,W], extra image [V,3,H,W] or None, max radii [P])."""
    lib = L.load()
    dev = means3D.device
    P = means3D.shape[0]
    H, W = cameras[0].image_height, cameras[0].image_width
    means3D, shs, opacities, scales, rotations = (t.detach().to(torch.float32).contiguous() for t in
                                                  (means3D, shs, opacities, scales, rotations))
    bg = bg.detach().to(dev, torch.float32).contiguous()
    f32 = dict(dtype=torch.float32, device=dev)
    Vt = len(cameras)
    color, depth = torch.empty(Vt, 3, H, W, **f32), torch.empty(Vt, 1, H, W, **f32)
    sem = torch.empty(Vt, 3, H, W, **f32) if extra is not None else None
    ex = None if extra is None else extra.detach().to(dev, torch.float32).reshape(-1).contiguous()
    radii_max = torch.zeros(P, dtype=torch.int32, device=dev)
    for lo in range(0, Vt, 64):
        cams = cameras[lo:lo + 64]
        V = len(cams)
        recs = torch.stack([camera_record(c) for c in cams]).to(dev, non_blocking=True)
        arena = _scratch_arena(dev, ("render", P, V, W, H))
        rm = torch.empty(P, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.dge_fit_views_forward(
                arena.cbs[0], arena.cbs[1], arena.cbs[2], None, P, sh_degree, shs.shape[1], V, bg.data_ptr(), W, H,
                means3D.data_ptr(), shs.data_ptr(), opacities.data_ptr(), scales.data_ptr(), float(scale_modifier),
                rotations.data_ptr(), recs.data_ptr(), color[lo:lo + V].data_ptr(), depth[lo:lo + V].data_ptr(),
                rm.data_ptr(), None, 0, None, 0, None, None if ex is None else ex.data_ptr(),
                None if sem is None else sem[lo:lo + V].data_ptr(), int(prune_lists), L.stream_ptr(dev)),
                "fit views forward (render)")
        radii_max = torch.maximum(radii_max, rm)
    return color, depth, sem, radii_max


def backproject_masks(means3D, opacities, scales, rotations, cameras, masks, weights, cnt, scale_modifier=1.0,
                      prune_lists=True):
    """DGE.update_mask's loop (threestudio/systems/DGE.py:112-147: one GaussianModel.apply_weights per
    camera, gaussian_model.py:817-832) as ONE call per 64 views: `masks[v]` ([CH,H,W], the 2-D
    segmentation mask of view v) is back-projected onto the Gaussians it is blended from; `weights`
    [P,CH] f32 and `cnt` [P(,1)] i32 are accumulated in place, exactly as the per-view calls do (binary
    masks give identical results bit for bit). Returns the per-view instance counts."""
    lib = L.load()
    dev = means3D.device
    P = means3D.shape[0]
    H, W = cameras[0].image_height, cameras[0].image_width
    if not (weights.is_contiguous() and cnt.is_contiguous()) or weights.dtype != torch.float32 or cnt.dtype != torch.int32:
        raise RuntimeError("backproject_masks: weights must be contiguous float32 and cnt contiguous int32")
    masks = masks if isinstance(masks, torch.Tensor) else torch.stack(list(masks))
    masks = masks.to(dev, torch.float32).contiguous()
    CH = masks.shape[1]
    means3D, opacities, scales, rotations = (t.detach().to(torch.float32).contiguous() for t in
                                             (means3D, opacities, scales, rotations))
    counts = []
    for lo in range(0, len(cameras), 64):
        cams = cameras[lo:lo + 64]
        V = len(cams)
        recs = torch.stack([camera_record(c) for c in cams]).to(dev, non_blocking=True)
        arena = _scratch_arena(dev, (P, V, W, H))
        nr = (L.C.c_int * V)()
        with torch.cuda.device(dev):
            L.check(lib.dge_fit_views_apply_weights(
                arena.cbs[0], arena.cbs[1], arena.cbs[2], None, P, V, W, H, means3D.data_ptr(), opacities.data_ptr(),
                scales.data_ptr(), float(scale_modifier), rotations.data_ptr(), recs.data_ptr(),
                masks[lo:lo + V].data_ptr(), CH, weights.data_ptr(), cnt.data_ptr(), nr, int(prune_lists),
                L.stream_ptr(dev)), "fit views apply_weights")
        counts += list(nr)
    return counts


class _GrowingArena:
    """Three scratch blobs that are kept between calls and only ever grow (update_mask is called with
    the same sizes every time; 6 GB through the caching allocator per call is not free)."""

    def __init__(self, device):
        self.device = device
        self.bufs = [torch.empty(0, dtype=torch.uint8, device=device) for _ in range(3)]
        self.cbs = [L.ALLOC_FN(self._make(i)) for i in range(3)]

    def _make(self, i):
        def alloc(_ctx, nbytes):
            if nbytes > self.bufs[i].numel():
                self.bufs[i] = torch.empty(0, dtype=torch.uint8, device=self.device)  # drop before growing
                self.bufs[i] = torch.empty(int(nbytes * 1.1) + 256, dtype=torch.uint8, device=self.device)
            return self.bufs[i].data_ptr()
        return alloc


_ARENAS = {}


def _scratch_arena(device, key):
    k = (str(device),) + tuple(key)
    if k not in _ARENAS:
        if len(_ARENAS) >= 2:
            _ARENAS.clear()
        _ARENAS[k] = _GrowingArena(device)
    return _ARENAS[k]


def shard_views(num_views: int, rank: int, world: int) -> List[int]:
    """View i of the step's batch goes to rank i mod world."""
    return list(range(rank, num_views, world))


class ViewSampler:
    """The batch of edited views of a step, drawn as GSLoadIterableDataset does (threestudio/data/gs_load.py:
    213-222 constructor, 256-272 collate, 286-292 update_cameras): `max_view_num` of the scene's cameras are
    sampled once with seed 0, a step takes `batch_size` of them at random WITHOUT replacement from a stack
    that is refilled when it runs empty. The reference seeds Python's global `random`; here the same
    generator is private to the sampler, so every rank that builds it with the same arguments draws the same
    batches whatever else uses `random` — the precondition of sharding a step's views."""

    def __init__(self, total_view_num: int, max_view_num: int, batch_size: int, seed: int = 0):
        import random
        self._random = random
        self.total_view_num, self.max_view_num, self.batch_size = total_view_num, max_view_num, batch_size
        self.update_cameras(seed)

    def update_cameras(self, random_seed: int = 0):
        self._rng = self._random.Random(random_seed)
        self.n2n_view_index = self._rng.sample(range(0, self.total_view_num),
                                               min(self.total_view_num, self.max_view_num))
        self.view_index_stack = self.n2n_view_index.copy()

    def next_batch(self) -> List[int]:
        """View indices of the next step (`batch["index"]` of the reference's collate)."""
        out = []
        for _ in range(self.batch_size):
            if not self.view_index_stack:
                self.view_index_stack = self.n2n_view_index.copy()
            view_index = self._rng.choice(self.view_index_stack)
            self.view_index_stack.remove(view_index)
            out.append(view_index)
        return out

    def next_shard(self, rank: int, world: int):
        """(whole batch, this rank's share of it): view i of the batch goes to rank i mod world."""
        batch = self.next_batch()
        return batch, [batch[i] for i in shard_views(len(batch), rank, world)]


class _ImagesWithBackward(torch.autograd.Function):
    """The step's rendered images as a differentiable tensor: whatever torch loss sits on top (L1, LPIPS, ...),
    its dL/dimages arrives here and goes through the batched backward into model.flat_grad."""

    @staticmethod
    def forward(ctx, hook, images, adapter):
        ctx.adapter = adapter
        return images.clone()

    @staticmethod
    def backward(ctx, g):
        ctx.adapter._backward(g.contiguous())
        return torch.zeros_like(ctx.adapter._hook), None, None


class DGEFitAdapter:
    """DGE's training step mapped onto the per-step family: what
    DGE.forward (threestudio/systems/DGE.py:170-239), loss.backward() and DGE.on_before_optimizer_step (:266-296)
    do around gaussian_renderer.render(), for this rank's share of the step's cameras, with ALL views per launch.

        out = adapter.forward(cameras, bg, mask=gaussian_mask)   # images, depth, semantic masks, radii
        loss = any_torch_loss(out["comp_rgb"], ...)               # e.g. DGE.py:637-683: 10 L1 + 10 LPIPS
        loss.backward()                                           # -> batched backward, raw-parameter gradients
        adapter.on_before_optimizer_step()                        # all-reduce over the ranks + densification stats
        adapter.optimizer_step()                                  # fused Adam (FitModel.adam_step)

    The loss must be the GLOBAL batch's loss restricted to this rank's views (e.g. an L1 mean over the global
    batch: sum over this rank's pixels / global pixel count), so that the all-reduced gradient is the step's.
    The second, mask-colour render DGE makes of every view (DGE.py:198-204) rides as a fourth blended channel of
    the same launches; `semantic` / `masks` are what DGE.forward derives from it."""

    def __init__(self, model: FitModel, process_group=None, prune_lists: bool = True, num_chunks: int = 1):
        self.model, self.group, self.prune_lists, self.num_chunks = model, process_group, prune_lists, num_chunks
        self._hook = torch.zeros((), device=model.device, requires_grad=True)
        self._radii = None

    def forward(self, cameras, bg, mask: Optional[torch.Tensor] = None):
        model = self.model
        if len(cameras) == 0:
            raise ValueError("DGEFitAdapter.forward needs at least one camera on every rank")
        _batched_forward(model, model.activations_fused(), cameras, bg, self.num_chunks, self.prune_lists,
                         extra=None if mask is None else mask.reshape(-1).float())
        batches, main = model._batches, model._fw["main"]
        for vb in batches:
            main.wait_stream(vb.stream)
        cat = lambda ts: ts[0] if len(ts) == 1 else torch.cat(ts)
        images, depth = cat([vb.color for vb in batches]), cat([vb.depth for vb in batches])
        radii = batches[0].radii_max
        for vb in batches[1:]:
            radii = torch.maximum(radii, vb.radii_max)
        self._radii = radii
        comp = _ImagesWithBackward.apply(self._hook, images, self)
        depths = depth.permute(0, 2, 3, 1)
        out = {"comp_rgb": comp.permute(0, 2, 3, 1), "depth": depths, "opacity": depths / (depths.max() + 1e-5),
               "radii": radii, "visibility_filter": radii > 0}
        if mask is not None:
            sem = cat([vb.sem for vb in batches])               # [V,3,H,W]: render(..., override_color=mask x3)
            semantic_map = torch.norm(sem, dim=1) > 0.8          # DGE.py:205-206
            viz = images.detach().clone().permute(0, 2, 3, 1)    # DGE.py:207-216
            viz[semantic_map] = 0.40 * viz[semantic_map] + 0.60 * torch.tensor([1.0, 0.0, 0.0], device=viz.device)
            out["semantic"], out["masks"], out["semantic_render"] = viz.permute(0, 3, 1, 2), semantic_map, sem
        return out

    __call__ = forward

    def _backward(self, dL_dimages):
        early = None
        self._radii = _batched_backward(self.model, dL_dimages, 1, early)

    def on_before_optimizer_step(self, update_stats: bool = True):
        """DGE.py:266-284 (+ the collective of SURVEY.md §8e): the gradients of all ranks' views summed, the
        radii maxed, xyz_gradient_accum / denom / max_radii2D updated. Densification itself
        (FitModel.densify_and_prune, DGE.py:286-296) stays with the caller's schedule."""
        zero = torch.zeros((), device=self.model.device)
        _finish_step(self.model, zero, self._radii, self.group, update_stats, None, adam=False)

    def optimizer_step(self):
        self.model.adam_step()


def default_rasterize(rs, means3D, means2D, shs, opacities, scales, rotations):
    return dgr.GaussianRasterizer(rs)(means3D=means3D, means2D=means2D, shs=shs, colors_precomp=None,
                                      opacities=opacities, scales=scales, rotations=rotations, cov3D_precomp=None)


def fit_step(model: FitModel, cameras: Sequence[scene.Camera], targets: Sequence[torch.Tensor], bg: torch.Tensor,
             global_batch: int, rasterize: Callable = default_rasterize, settings_module=dgr,
             process_group=None, lambda_l1: float = 10.0, host_inputs: bool = False, update_stats: bool = True,
             num_streams: int = 1, direct: Optional[bool] = None, batched: Optional[bool] = None,
             num_chunks: int = 1, prune_lists: bool = True, geom_splits: Optional[int] = None,
             image_size: Optional[Sequence[int]] = None, next_cameras: Optional[Sequence[scene.Camera]] = None):
    """One optimisation step over this rank's views. `cameras`/`targets` are this rank's share;
    `global_batch` the number of views in the whole step (L1 is a mean over the global batch,
    DGE.py:672). With host_inputs the cameras/targets live in pinned host memory and are copied
    inside the step. Returns the step's loss as a 0-d device tensor (already globally reduced).

    Views are independent until the optimiser step, so with num_streams > 1 they are issued
    round-robin on that many CUDA streams: one view's latency-bound stages (sorts, the tail of
    the blend over the densest tiles) overlap another view's kernels. Each stream accumulates
    into its own gradient leaves; the partial sums are added on the main stream afterwards.

    direct (default: on CUDA with the stock rasterizer) drives the C-ABI without autograd. batched
    (default) pushes ALL views of the step through each stage in one launch (_batched_views: the view
    is a grid dimension, the Gaussians are read once per step, one host wait per step); batched=False
    issues the views one by one round-robin on num_streams CUDA streams (_direct_views). The autograd path below is the reference-shaped one (per-view tensors,
    torch ops for the loss, AccumulateGrad) and is what `rasterize` overrides go through.

    next_cameras (batched path): this rank's cameras of the NEXT step (fit.ViewSampler knows them). On several
    GPUs the step then ends by projecting, depth-sorting and binning the next step's views — none of which needs
    the SH coefficients — while the all-reduce of the f_rest gradient, three quarters of the step's bytes, is on
    the wire; the next call finds that front half (same cameras), fills in the colours from the freshly stepped
    features and goes straight to the blend. Results are those of the unpipelined step (same kernels, same
    inputs); a front half that does not match the next call is simply recomputed."""
    dev = model.device
    if len(cameras) == 0:
        # this rank's share of the batch is empty (fewer views than ranks): it contributes zero gradients,
        # zero loss and zero radii, and still joins every collective of the step
        model.flat_grad.zero_()
        return _finish_step(model, torch.zeros((), device=dev), torch.zeros(model.P, dtype=torch.int32, device=dev),
                            process_group, update_stats)
    H, W = cameras[0].image_height, cameras[0].image_width
    if image_size is not None and (int(image_size[0]), int(image_size[1])) != (W, H):
        raise ValueError(f"image_size {tuple(image_size)} does not match the cameras' {W}x{H}")
    scale = lambda_l1 / float(global_batch * 3 * H * W)
    if direct is None:
        direct = dev.type == "cuda" and rasterize is default_rasterize and model.sh_degree == 3
    if direct:
        if batched is None:
            batched = len(cameras) <= 64
        if batched:
            # a chunk's instance lists share one arena addressed with 30-bit positions: when the views of a
            # chunk hold more than 2^30 instances (6 M Gaussians at 1080p: ~60 M per view), split further
            chunks = max(num_chunks, getattr(model, "_min_chunks", 1))
            world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
            splits = geom_splits if geom_splits is not None else (GEOM_SPLITS_MULTI_GPU if world > 1 else 1)
            early = _EarlyRestReduce(model, process_group) if (world > 1 and splits > 1) else None
            while True:
                try:
                    loss, radii_max = _batched_views(model, cameras, targets, bg, scale, host_inputs, chunks,
                                                     prune_lists, splits, early)
                    break
                except RuntimeError as ex:
                    if "2^30" not in str(ex) or chunks >= len(cameras):
                        raise
                    torch.cuda.synchronize(dev)
                    chunks = min(len(cameras), chunks * 2)
                    model._min_chunks = chunks
        else:
            early = None
            loss, radii_max = _direct_views(model, model.activations_fused(), cameras, targets, bg, scale, host_inputs,
                                            num_streams)
        prefetch = None
        if batched and next_cameras is not None and len(next_cameras) > 0:
            def prefetch():
                _batched_front(model, model.activations_fused("geometry"), next_cameras, chunks, prune_lists,
                               with_colour=False)
        return _finish_step(model, loss, radii_max, process_group, update_stats, early, prefetch=prefetch)
    model.zero_grad()
    acts_graph = model.activations()
    S = max(1, min(num_streams, len(cameras)))
    main = torch.cuda.current_stream(dev) if dev.type == "cuda" else None
    if S > 1:
        if len(getattr(model, "_streams", [])) < S:
            model._streams = [torch.cuda.Stream(dev) for _ in range(S)]
        streams = model._streams[:S]
        for st in streams:
            st.wait_stream(main)
    else:
        streams = [None]

    lanes = []  # per stream: detached activation leaves, means2D tap, loss and radii accumulators
    for st in streams:
        with torch.cuda.stream(st) if st is not None else _null():
            acts = {k: v.detach().requires_grad_(True) for k, v in acts_graph.items()}
            m2d = model.means2D if S == 1 else torch.zeros_like(model.means2D, requires_grad=True)
            lanes.append(dict(acts=acts, m2d=m2d, loss=torch.zeros((), device=dev),
                              radii=torch.zeros(model.P, dtype=torch.int32, device=dev)))
    for i, (cam, target) in enumerate(zip(cameras, targets)):
        lane, st = lanes[i % S], streams[i % S]
        with torch.cuda.stream(st) if st is not None else _null():
            if host_inputs:
                cam = scene.camera_to(cam, dev, non_blocking=True)
                target = target.to(dev, non_blocking=True)
            rs = scene.raster_settings(cam, bg, model.sh_degree, module=settings_module)
            a = lane["acts"]
            color, radii, _depth = rasterize(rs, a["means3D"], lane["m2d"], a["shs"], a["opacities"], a["scales"],
                                             a["rotations"])
            lv = (color - target).abs().sum() * scale
            lv.backward()
            lane["loss"] += lv.detach()
            torch.maximum(lane["radii"], radii, out=lane["radii"])
    if S > 1:
        for st in streams:
            main.wait_stream(st)
    # fold the per-stream partial sums (main stream)
    acts = lanes[0]["acts"]
    loss, radii_max = lanes[0]["loss"], lanes[0]["radii"]
    for lane in lanes[1:]:
        for k in acts:
            if lane["acts"][k].grad is not None:
                acts[k].grad.add_(lane["acts"][k].grad)
        loss = loss + lane["loss"]
        radii_max = torch.maximum(radii_max, lane["radii"])
    if S > 1:
        g2 = model.means2D.grad
        for lane in lanes:
            if lane["m2d"].grad is not None:
                g2.add_(lane["m2d"].grad)
    # one backward through the activations for the whole step
    through = [k for k in acts_graph if acts_graph[k].grad_fn is not None and acts[k].grad is not None]
    if through:
        torch.autograd.backward([acts_graph[k] for k in through], [acts[k].grad for k in through])
    # xyz has no activation: its detached leaf's grad goes straight into the flat buffer
    if acts["means3D"].grad is not None:
        model.params["xyz"].grad.add_(acts["means3D"].grad)
    return _finish_step(model, loss, radii_max, process_group, update_stats)


# Measured on 2 B200s (bench.py --gpus 2 --geom-splits 1/2/4/8: 8.36 / 8.43 / 8.37 / 8.43 ms per step): sending
# three quarters of f_rest under the per-Gaussian backward buys nothing — that kernel is bandwidth-bound and
# short (0.46 ms), NCCL's copy kernels take their share of it back — so one launch stays the default.
GEOM_SPLITS_MULTI_GPU = 1
# all-reduce pieces of the f_rest gradient, each followed by its own Adam launch. Measured on 2 B200s: four pieces
# take 0.55 ms on the wire against 0.38 ms for one (smaller messages, and Adam competing for HBM), which eats the
# 0.14 ms of Adam they hide: 8.38 vs 8.36 ms per step. One piece.
REST_PIECES = 1


class _EarlyRestReduce:
    """Multi-GPU: the per-Gaussian backward runs as a few launches over consecutive Gaussian ranges; as
    soon as a range is queued, the all-reduce of its f_rest gradient rows (45 of the 62 floats per
    Gaussian, contiguous in the group-major flat buffer) is issued on NCCL's stream and runs under the
    next range's kernel. The LAST range is left to _finish_step, which sends the small groups first."""

    def __init__(self, model, process_group):
        self.model, self.group, self.works, self.done_rows = model, process_group, [], 0

    def __call__(self, first, count):
        if first == 0:
            self.works, self.done_rows = [], 0
        if first + count >= self.model.P:  # last range: reduced after the small groups
            return
        g = self.model.params["f_rest"].grad.view(self.model.P, -1)
        self.works.append(dist.all_reduce(g[first:first + count], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.done_rows = first + count


def _finish_step(model, loss, radii_max, process_group, update_stats, early=None, adam=True, prefetch=None):
    """The collective, the densification statistics and (adam=True) the optimiser step. With adam=False the
    gradients are left reduced in model.flat_grad for a separate FitModel.adam_step() (DGEFitAdapter).
    prefetch(): launches the next step's geometry front half; called once the geometry groups have been stepped,
    i.e. while the f_rest gradient is still being reduced on several GPUs (fit_step's next_cameras)."""
    world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
    pending, geometry_stepped = None, False
    if world > 1:
        # THE collective: 59P parameter grads + 3P screen-space grads, SUM — issued as
        # contiguous pieces of the flat buffer so that the optimiser can start on the small parameter
        # groups (and the statistics on the screen-space gradient) while the 45P floats of f_rest, three
        # quarters of the bytes, are still on the wire; MAX over the radii, SUM over the loss.
        a0 = model.slices["f_rest"].start
        sent = early.done_rows if early is not None else 0  # f_rest rows already on the wire
        ar = lambda t, op=dist.ReduceOp.SUM: dist.all_reduce(t, op=op, group=process_group, async_op=True)
        # what the next step's projection waits for goes first, as ONE call: the geometry groups' gradients are
        # contiguous in the flat buffer (LAYOUT). Then what only the statistics and the caller need — the radii (MAX)
        # and, in one call, the screen-space gradient + the loss —, then the features: f_dc, and f_rest in
        # REST_PIECES row ranges (Adam on a piece runs under the next piece's transfer)
        model.loss_slot.copy_(loss.reshape(1))
        loss = model.loss_slot[0]
        if prefetch is not None:
            radii_max = radii_max.clone()  # the prefetched front half reuses the chunks' radii buffers
        works = [ar(model.flat_grad[model.early_slice])]
        radii_work = ar(radii_max, dist.ReduceOp.MAX)
        stats_work = ar(model.flat_grad[model.stats_slice])
        dc = model.slices["f_dc"]
        late_dc = ar(model.flat_grad[dc.start:dc.stop])
        P = model.P
        cuts = [sent] + [max(sent, (P * k // REST_PIECES) // 4 * 4) for k in range(1, REST_PIECES)] + [P]
        rest = [(r0, r1, ar(model.flat_grad[a0 + 45 * r0:a0 + 45 * r1])) for r0, r1 in zip(cuts[:-1], cuts[1:]) if r1 > r0]
        if not model.fused_adam:
            works += [late_dc] + [w for _, _, w in rest]
            rest, late_dc = [], None
        pending = rest
        for w in works:
            w.wait()
        if prefetch is not None and adam and model.fused_adam:
            # the geometry groups are final once stepped: project / sort / bin the next step's views now, under
            # the features' all-reduce (the statistics below do not touch what the front half reads)
            model.adam_step(only=GEOMETRY_GROUPS)
            prefetch()
            prefetch, geometry_stepped = None, True
        radii_work.wait()
        stats_work.wait()
        loss = loss.clone()
    fused_stats = (update_stats and model.device.type == "cuda" and model.fused_adam
                   and radii_max.dtype == torch.int32 and model.max_radii2D.dtype == torch.int32
                   and model.xyz_gradient_accum.dtype == model.denom.dtype == torch.float32
                   and all(t.is_cuda and t.is_contiguous() for t in (radii_max, model.max_radii2D,
                                                                     model.xyz_gradient_accum, model.denom)))
    if fused_stats:
        # DGE.py:266-284, gaussian_model.py:811-815 in one pass (the torch statements below are ten launches)
        L.check(L.load().dge_fit_update_stats(
            model.P, radii_max.data_ptr(), model.means2D.grad.data_ptr(), model.max_radii2D.data_ptr(),
            model.xyz_gradient_accum.data_ptr(), model.denom.data_ptr(), L.stream_ptr(model.device)), "update stats")
    elif update_stats:
        with torch.no_grad():  # DGE.py:266-284, gaussian_model.py:811-815
            vis = radii_max > 0
            model.max_radii2D = torch.where(vis, torch.maximum(model.max_radii2D, radii_max), model.max_radii2D)
            gnorm = model.means2D.grad[:, :2].norm(dim=-1, keepdim=True)
            model.xyz_gradient_accum += torch.where(vis[:, None], gnorm, torch.zeros_like(gnorm))
            model.denom += vis[:, None].to(model.denom.dtype)
    if not adam:
        for w in ([w for _, _, w in pending] if pending else []) + (early.works if early is not None else []) + (
                [late_dc] if (world > 1 and late_dc is not None) else []):
            w.wait()
        return loss
    if pending is not None and model.fused_adam:
        if not geometry_stepped:
            model.adam_step(only=GEOMETRY_GROUPS)
        late_dc.wait()
        model.adam_step(only=("f_dc",), advance=False)
        if early is not None and early.done_rows:
            for w in early.works:
                w.wait()
            model.adam_step(only=("f_rest",), advance=False, rows=(0, early.done_rows))
        for r0, r1, w in pending:
            w.wait()
            model.adam_step(only=("f_rest",), advance=False, rows=(r0, r1))
    else:
        for w in (early.works if early is not None else []):
            w.wait()
        if model.fused_adam and prefetch is not None:
            # one process: the same order of work as on several (geometry groups, front half, features)
            model.adam_step(only=GEOMETRY_GROUPS)
            prefetch()
            model.adam_step(only=("f_dc", "f_rest"), advance=False)
        else:
            model.adam_step()
    return loss


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
