"""The reference's OWN callers, unchanged, on the drop-in's Python surface (CPU; needs /root/reference, so it runs
in the build container and skips on the GPU box, where the reference tree does not exist).

gaussiansplatting/gaussian_renderer/__init__.py (render, camera2rasterizer) and GaussianModel
(gaussiansplatting/scene/gaussian_model.py: the activations and apply_weights, :221-258, :817-832) are loaded
straight from the reference tree, with `diff_gaussian_rasterization` resolving to dge_b200's module
(dge_b200.install()). The native library cannot run here (no GPU), so the three places where the binding calls
into it are replaced by recording fakes that check what the C-ABI would be handed (tensor shapes, dtypes,
contiguity, scalar arguments) and return tensors of the right shapes; everything else — the keyword call of
render(), the positional call of apply_weights, GaussianRasterizationSettings' fields, the autograd Function's
argument order and the gradients it hands back to DGE's tensors — is the real code on both sides.
What the numbers are is the GPU suite's business (tests/test_parity*_gpu.py, test_fit_gpu.py)."""
import importlib.util
import math
import os
import sys
import types

import pytest
import torch

REF = "/root/reference/gaussiansplatting"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")


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


class _cpu_for_cuda:
    """device="cuda" -> "cpu" in torch's factory functions and Tensor.to (the reference hard-codes "cuda")."""
    NAMES = ["zeros", "ones", "empty", "tensor", "full", "zeros_like", "ones_like"]

    def __enter__(self):
        self.saved = {n: getattr(torch, n) for n in self.NAMES}
        self.saved_to = torch.Tensor.to

        def redirect(fn):
            def wrapped(*a, **kw):
                if str(kw.get("device", "")).startswith("cuda"):
                    kw["device"] = "cpu"
                return fn(*a, **kw)
            return wrapped
        for n in self.NAMES:
            setattr(torch, n, redirect(self.saved[n]))
        saved_to = self.saved_to

        def to(t, *a, **kw):
            a = tuple("cpu" if isinstance(x, str) and x.startswith("cuda") else x for x in a)
            return saved_to(t, *a, **kw)
        torch.Tensor.to = to
        return self

    def __exit__(self, *exc):
        for n, f in self.saved.items():
            setattr(torch, n, f)
        torch.Tensor.to = self.saved_to
        return False


@pytest.fixture()
def reference_modules():
    import dge_b200
    saved = dict(sys.modules)
    dgr = dge_b200.install()
    _stub("gaussiansplatting"), _stub("gaussiansplatting.utils"), _stub("gaussiansplatting.scene")
    _load("gaussiansplatting.utils.general_utils", REF + "/utils/general_utils.py")
    _load("gaussiansplatting.utils.sh_utils", REF + "/utils/sh_utils.py")
    _stub("gaussiansplatting.utils.system_utils", mkdir_p=lambda p: None)
    _stub("plyfile", PlyData=object, PlyElement=object)
    _stub("simple_knn"), _stub("simple_knn._C", distCUDA2=None)
    _stub("gaussiansplatting.utils.graphics_utils", BasicPointCloud=object)
    _stub("gaussiansplatting.knn", K_nearest_neighbors=None)
    renderer = _load("gaussiansplatting.gaussian_renderer", REF + "/gaussian_renderer/__init__.py")
    model = _load("gaussiansplatting.scene.gaussian_model", REF + "/scene/gaussian_model.py")
    yield dgr, renderer, model.GaussianModel
    for k in list(sys.modules):
        if k not in saved:
            del sys.modules[k]


def _camera(W, H):
    from dge_b200 import scene
    return scene.ring_cameras(3, W, H)[1]


def test_reference_render_and_apply_weights_run_on_the_drop_in(reference_modules, monkeypatch):
    dgr, renderer, GaussianModel = reference_modules
    assert renderer.GaussianRasterizer is dgr.GaussianRasterizer            # the import swap took
    assert renderer.GaussianRasterizationSettings is dgr.GaussianRasterizationSettings
    P, W, H = 300, 48, 32
    calls = []

    def fake_forward(rs, means3D, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, sh):
        for t in (means3D, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, sh, rs.bg, rs.viewmatrix,
                  rs.projmatrix, rs.campos):
            assert t.dtype == torch.float32
        assert means3D.shape == (P, 3) and opacities.shape == (P, 1) and scales.shape == (P, 3) and rotations.shape == (P, 4)
        assert (sh.numel() == 0) != (colors_precomp.numel() == 0)      # exactly one of SHs / precomputed colours
        assert sh.numel() == 0 or sh.shape == (P, 16, 3)
        assert colors_precomp.numel() == 0 or colors_precomp.shape == (P, 3)
        assert cov3Ds_precomp.numel() == 0
        assert (int(rs.image_height), int(rs.image_width)) == (H, W) and rs.prefiltered is False and rs.debug is False
        assert rs.viewmatrix.shape == (4, 4) and rs.projmatrix.shape == (4, 4) and rs.campos.shape == (3,)
        calls.append(("forward", int(rs.sh_degree), sh.numel() != 0))
        e = torch.empty(0, dtype=torch.uint8)
        return 7, torch.full((3, H, W), 0.25), torch.full((1, H, W), 2.0), torch.ones(P, dtype=torch.int32), e, e, e

    def fake_backward(rs, means3D, radii, colors_precomp, scales, rotations, cov3Ds_precomp, grad_out_color, sh, geom,
                      num_rendered, binning, img, need):
        assert grad_out_color.shape == (3, H, W) and num_rendered == 7 and radii.dtype == torch.int32
        calls.append(("backward", need["colors"], need["cov3D"]))
        M = sh.shape[1] if sh.numel() else 0
        # (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)
        return (torch.full((P, 3), 1.0), torch.full((P, 3), 2.0) if need["colors"] else None, torch.full((P, 1), 3.0),
                torch.full((P, 3), 4.0), None, torch.full((P, M, 3), 5.0), torch.full((P, 3), 6.0), torch.full((P, 4), 7.0))

    class FakeLib:
        def dge_apply_weights(self, *a):
            calls.append(("apply_weights", a))
            return 5

    monkeypatch.setattr(dgr, "_forward_call", fake_forward)
    monkeypatch.setattr(dgr, "_backward_call", fake_backward)
    monkeypatch.setattr(dgr.L, "load", lambda: FakeLib())
    monkeypatch.setattr(dgr.L, "stream_ptr", lambda dev=None: None)
    monkeypatch.setattr(dgr, "_empty", lambda: torch.Tensor([]).to(torch.float32))
    monkeypatch.setattr(torch.cuda, "device", lambda d: _NullCtx())

    with _cpu_for_cuda():
        pc = GaussianModel(3, 0.0, 0.05, 1.0)                         # DGE.configure: sh_degree 3 (DGE.py:85-92)
        gen = torch.Generator().manual_seed(1)
        nn = torch.nn
        pc._xyz = nn.Parameter(torch.randn(P, 3, generator=gen))
        pc._features_dc = nn.Parameter(torch.randn(P, 1, 3, generator=gen))
        pc._features_rest = nn.Parameter(torch.randn(P, 15, 3, generator=gen))
        pc._opacity = nn.Parameter(torch.randn(P, 1, generator=gen))
        pc._scaling = nn.Parameter(torch.randn(P, 3, generator=gen))
        pc._rotation = nn.Parameter(torch.randn(P, 4, generator=gen))
        pc.active_sh_degree = 3
        pc.mask = torch.rand(P, generator=gen) < 0.5
        cam = _camera(W, H)
        pipe = types.SimpleNamespace(compute_cov3D_python=False, convert_SHs_python=False)
        bg = torch.zeros(3)
        # DGE.forward's two renders of a view (threestudio/systems/DGE.py:181, 198-204)
        pkg = renderer.render(cam, pc, pipe, bg)
        sem = renderer.render(cam, pc, pipe, bg, override_color=pc.mask[..., None].float().repeat(1, 3))["render"]
        assert set(pkg) == {"render", "viewspace_points", "visibility_filter", "radii", "depth_3dgs"}
        assert pkg["render"].shape == (3, H, W) and pkg["depth_3dgs"].shape == (1, H, W) and sem.shape == (3, H, W)
        assert pkg["radii"].dtype == torch.int32 and pkg["visibility_filter"].dtype == torch.bool
        (pkg["render"].sum() + 0.0 * sem.sum()).backward()
        # the gradients the autograd Function returned reached DGE's tensors in the reference's argument order
        assert torch.equal(pkg["viewspace_points"].grad, torch.full((P, 3), 1.0))    # means2D (DGE.py:269-276)
        assert torch.equal(pc._xyz.grad, torch.full((P, 3), 4.0) * 2)                 # means3D, both renders
        assert pc._features_dc.grad is not None and torch.equal(pc._features_dc.grad, torch.full((P, 1, 3), 5.0))
        assert torch.equal(pc._features_rest.grad, torch.full((P, 15, 3), 5.0))
        assert pc._opacity.grad is not None and pc._scaling.grad is not None and pc._rotation.grad is not None
        # GaussianModel.apply_weights (gaussian_model.py:817-832): positional call, in-place accumulators
        weights = torch.zeros(P, 1)
        cnt = torch.zeros(P, 1, dtype=torch.int32)
        image_weights = torch.ones(1, H, W)
        pc.apply_weights(cam, weights, cnt, image_weights)
    kinds = [c[0] for c in calls]
    assert kinds == ["forward", "forward", "backward", "backward", "apply_weights"]
    assert calls[0] == ("forward", 3, True) and calls[1] == ("forward", 3, False)
    assert {calls[2][1:], calls[3][1:]} == {(False, False), (True, False)}
    a = calls[4][1]
    # dge_apply_weights(alloc x3, ctx, P, D, M, bg, W, H, means3D, shs, weights, opacities, scales, scale_modifier, ...)
    assert a[4] == P and a[5] == 0 and a[6] == 0 and a[8] == W and a[9] == H      # camera2rasterizer: sh_degree 0
    assert a[12] == weights.data_ptr() and a[26] == cnt.data_ptr() and a[27] == 1   # in place; 1 channel
    assert a[11] is None and a[17] is None                                          # no SHs, no precomputed cov3D


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
