"""Generates tests/golden/densify_*.npz by running the REFERENCE's own densification code
(gaussiansplatting/scene/gaussian_model.py: densify_and_clone :728-768, densify_and_split :675-726,
prune_points :589-607 and the optimizer surgery :543-640, orchestrated as densify_and_prune :770-807) on
the CPU of this container. Test infrastructure only; needs /root/reference, so it runs HERE, and only the
fixtures travel.

How the reference is made to run without a GPU and without its package's heavy imports:
  * gaussian_model.py is loaded straight from its file; the modules it imports but does not need
    for densification (plyfile, simple_knn, kornia-based graphics_utils, the renderer, knn) are stubs;
    gaussiansplatting/utils/general_utils.py (build_rotation, inverse_sigmoid) is the real file.
  * its hard-coded device="cuda" arguments are redirected to the CPU by wrapping the torch factory
    functions for the duration of the run.
  * torch.normal is wrapped to RECORD the samples densify_and_split draws, so that the fixture holds
    them and the implementation under test can be fed the same draw.
The anchor / grad-mask hook bookkeeping at the end of densify_and_prune (:802-806) is not part of the
hot path's state and is left out: the script calls the reference's methods in the order of :771-797.

  python oracle/make_densify_golden.py     # writes tests/golden/densify_{a,b}.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/gaussiansplatting"
FACTORIES = ["zeros", "ones", "empty", "tensor", "full", "arange", "zeros_like", "ones_like"]


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def load_reference_model_class():
    _stub("gaussiansplatting"), _stub("gaussiansplatting.utils"), _stub("gaussiansplatting.scene")
    _load("gaussiansplatting.utils.general_utils", REF + "/utils/general_utils.py")
    _stub("gaussiansplatting.utils.system_utils", mkdir_p=lambda p: None)
    _stub("plyfile", PlyData=object, PlyElement=object)
    _stub("gaussiansplatting.utils.sh_utils", RGB2SH=lambda x: x)
    _stub("simple_knn"), _stub("simple_knn._C", distCUDA2=None)
    _stub("gaussiansplatting.utils.graphics_utils", BasicPointCloud=object)
    _stub("gaussiansplatting.gaussian_renderer", camera2rasterizer=None)
    _stub("gaussiansplatting.knn", K_nearest_neighbors=None)
    return _load("gaussiansplatting.scene.gaussian_model", REF + "/scene/gaussian_model.py").GaussianModel


class cpu_for_cuda:
    """device="cuda" -> "cpu" in torch's factory functions; records torch.normal draws."""

    def __enter__(self):
        self.saved = {n: getattr(torch, n) for n in FACTORIES + ["normal"]}
        self.normals = []

        def redirect(fn):
            def wrapped(*a, **kw):
                if str(kw.get("device", "")).startswith("cuda"):
                    kw["device"] = "cpu"
                return fn(*a, **kw)
            return wrapped

        for n in FACTORIES:
            setattr(torch, n, redirect(self.saved[n]))

        def normal(*a, **kw):
            out = self.saved["normal"](*a, **kw)
            self.normals.append(out.clone())
            return out

        torch.normal = normal
        return self

    def __exit__(self, *exc):
        for n, f in self.saved.items():
            setattr(torch, n, f)
        return False


GROUP_NAMES = ["xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation"]


def make_case(path, P, seed, max_grad, max_densify_percent, min_opacity, extent, max_screen_size, mask_fraction):
    GaussianModel = load_reference_model_class()
    g = torch.Generator().manual_seed(seed)
    rnd = lambda *s: torch.randn(*s, generator=g)
    with cpu_for_cuda() as ctx:
        gm = GaussianModel(3, 0.0, 0.1, 2.0)
        raw = {
            "xyz": rnd(P, 3),
            "f_dc": rnd(P, 1, 3) * 0.5,
            "f_rest": rnd(P, 15, 3) * 0.1,
            # opacities on both sides of min_opacity; scales on both sides of percent_dense * extent and 0.1 * extent
            "opacity": rnd(P, 1) * 3.0 - 1.0,
            "scaling": torch.log(torch.exp(rnd(P, 3) * 1.2 + np.log(0.01 * extent))),
            "rotation": rnd(P, 4),
        }
        gm._xyz, gm._features_dc, gm._features_rest = (torch.nn.Parameter(raw[k].clone()) for k in ("xyz", "f_dc", "f_rest"))
        gm._opacity, gm._scaling, gm._rotation = (torch.nn.Parameter(raw[k].clone()) for k in ("opacity", "scaling", "rotation"))
        gm.percent_dense = 0.01
        params = [gm._xyz, gm._features_dc, gm._features_rest, gm._opacity, gm._scaling, gm._rotation]
        groups = [{"params": [p], "lr": 1e-3, "name": n} for p, n in zip(params, GROUP_NAMES)]
        gm.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)   # gaussian_model.py:374
        for p in params:                                             # two steps so that exp_avg / exp_avg_sq are populated
            p.grad = torch.randn(p.shape, generator=g) * 1e-2
        gm.optimizer.step()
        for p in params:
            p.grad = torch.randn(p.shape, generator=g) * 1e-2
        gm.optimizer.step()
        before = {n: p.detach().clone().numpy() for n, p in zip(GROUP_NAMES, params)}
        state = {n: (gm.optimizer.state[p]["exp_avg"].clone().numpy(), gm.optimizer.state[p]["exp_avg_sq"].clone().numpy(),
                     int(gm.optimizer.state[p]["step"])) for n, p in zip(GROUP_NAMES, params)}
        gm.xyz_gradient_accum = torch.rand(P, 1, generator=g) * 4 * max_grad * 20
        gm.denom = torch.randint(0, 40, (P, 1), generator=g).float()    # zeros -> NaN grads -> 0 (:773)
        gm.max_radii2D = torch.randint(0, 12, (P,), generator=g).float()
        gm.mask = torch.rand(P, generator=g) < mask_fraction
        gm._generation = torch.zeros(P, dtype=torch.int64)
        stats_in = dict(xyz_gradient_accum=gm.xyz_gradient_accum.clone().numpy(), denom=gm.denom.clone().numpy(),
                        max_radii2D=gm.max_radii2D.clone().numpy(), mask=gm.mask.clone().numpy())

        # ---- densify_and_prune, gaussian_model.py:771-797 (the reference's own methods), called under
        # torch.no_grad() as DGE.on_before_optimizer_step does (DGE.py:266-267)
        torch.set_grad_enabled(False)
        grads = gm.xyz_gradient_accum / gm.denom
        grads[grads.isnan()] = 0.0
        grads[~gm.mask] = 0.0
        if max_densify_percent < 1:
            valid_percent = len(grads.nonzero()) * max_densify_percent / grads.shape[0]
            thresold_value = torch.quantile(grads, 1 - valid_percent)
            grads[grads < thresold_value] = 0.0
        n0 = gm.get_xyz.shape[0]
        gm.densify_and_clone(grads, max_grad, extent)
        n1 = gm.get_xyz.shape[0]
        gm.densify_and_split(grads, max_grad, extent)
        n2 = gm.get_xyz.shape[0]
        prune_mask = (gm.get_opacity < min_opacity).squeeze()
        if max_screen_size:
            big_points_vs = gm.max_radii2D > max_screen_size
            big_points_ws = gm.get_scaling.max(dim=1).values > 0.1 * extent
            prune_mask = torch.logical_or(torch.logical_or(prune_mask, big_points_vs), big_points_ws)
        prune_mask = torch.logical_and(prune_mask, gm.mask)
        gm.prune_points(prune_mask)
        n3 = gm.get_xyz.shape[0]

        after_params = [gm._xyz, gm._features_dc, gm._features_rest, gm._opacity, gm._scaling, gm._rotation]
        out = {}
        for n, p in zip(GROUP_NAMES, after_params):
            out["out_" + n] = p.detach().numpy()
            st = gm.optimizer.state[gm.optimizer.param_groups[GROUP_NAMES.index(n)]["params"][0]]
            out["out_m_" + n], out["out_v_" + n] = st["exp_avg"].detach().numpy(), st["exp_avg_sq"].detach().numpy()
        out.update(out_mask=gm.mask.numpy(), out_generation=gm._generation.numpy(),
                   out_xyz_gradient_accum=gm.xyz_gradient_accum.numpy(), out_denom=gm.denom.numpy(),
                   out_max_radii2D=gm.max_radii2D.numpy())
        samples = ctx.normals[0].detach().numpy() if ctx.normals else np.zeros((0, 3), np.float32)
        torch.set_grad_enabled(True)
    inp = {"in_" + n: before[n] for n in GROUP_NAMES}
    for n in GROUP_NAMES:
        inp["in_m_" + n], inp["in_v_" + n] = state[n][0], state[n][1]
    inp.update({"in_" + k: v for k, v in stats_in.items()})
    np.savez_compressed(path, counts=np.array([n0, n1, n2, n3]), normal_samples=samples,
                        hyper=np.array([max_grad, max_densify_percent, min_opacity, extent, max_screen_size, 0.01], np.float64),
                        adam_step=np.array([state["xyz"][2]]), **inp, **out)
    print(path, "P", n0, "-> clone", n1, "-> split", n2, "-> prune", n3, "split samples", samples.shape)


if __name__ == "__main__":
    gold = os.path.join(ROOT, "tests", "golden")
    # DGE defaults: max_densify_percent 0.01, min_opacity 0.005, max_screen_size 5 (DGE.py:39-54, :290-296)
    make_case(os.path.join(gold, "densify_a.npz"), 1500, 11, 0.0002, 0.01, 0.005, 4.0, 5, 0.7)
    # no quantile cut, everything editable, no screen-size pruning
    make_case(os.path.join(gold, "densify_b.npz"), 1000, 12, 0.0002, 1.0, 0.05, 2.0, 0, 1.0)
