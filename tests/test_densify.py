"""FitModel.densify_and_prune against fixtures produced by the REFERENCE's own
GaussianModel.densify_and_clone / densify_and_split / prune_points
(gaussiansplatting/scene/gaussian_model.py:543-807, run on the CPU by oracle/make_densify_golden.py, with the
draw of torch.normal recorded). Everything is index / copy / elementwise work in the same torch ops, so
parameters, Adam state, masks and statistics must be IDENTICAL, fused-Adam layout or torch.optim layout."""
import glob
import os

import numpy as np
import pytest
import torch

from dge_b200 import fit, scene

FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "densify_*.npz")))
NAMES = [g[0] for g in fit.GROUPS]


def _model(z, fused):
    P = z["in_xyz"].shape[0]
    dummy = scene.make_gaussians(P, seed=1)
    model = fit.FitModel(dummy, torch.device("cpu"), fused_adam=fused)
    raw = {n: torch.from_numpy(z["in_" + n]) for n in NAMES}
    m = {n: torch.from_numpy(z["in_m_" + n]) for n in NAMES}
    v = {n: torch.from_numpy(z["in_v_" + n]) for n in NAMES}
    model.step_count = int(z["adam_step"][0])
    model._allocate(raw, m, v)
    model.xyz_gradient_accum = torch.from_numpy(z["in_xyz_gradient_accum"]).clone()
    model.denom = torch.from_numpy(z["in_denom"]).clone()
    model.max_radii2D = torch.from_numpy(z["in_max_radii2D"]).to(torch.int32)
    model.set_grad_mask(torch.from_numpy(z["in_mask"]))
    return model


def test_fixtures_present():
    assert len(FILES) >= 2, "run oracle/make_densify_golden.py (needs /root/reference)"


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_densify_and_prune_matches_reference(path, fused):
    z = np.load(path)
    max_grad, pct, min_opacity, extent, max_screen, percent_dense = (float(x) for x in z["hyper"])
    model = _model(z, fused)
    counts = model.densify_and_prune(max_grad, pct, min_opacity, extent, int(max_screen), percent_dense=percent_dense,
                                     normal_samples=torch.from_numpy(z["normal_samples"]))
    assert list(counts) == [int(c) for c in z["counts"]]
    assert model.P == int(z["counts"][3])
    for n in NAMES:
        assert np.array_equal(model.params[n].detach().numpy(), z["out_" + n]), n
        m, v = model.adam_state(n)
        assert np.array_equal(m.detach().numpy(), z["out_m_" + n]), n
        assert np.array_equal(v.detach().numpy(), z["out_v_" + n]), n
    assert np.array_equal(model.grad_mask.bool().numpy(), z["out_mask"])
    assert np.array_equal(model.xyz_gradient_accum.numpy(), z["out_xyz_gradient_accum"])
    assert np.array_equal(model.denom.numpy(), z["out_denom"])
    assert np.array_equal(model.max_radii2D.numpy().astype(np.float32), z["out_max_radii2D"])
    # the rebuilt model is a working FitModel: leaves alias the flat buffers, gradients alias flat_grad
    for n, sl in model.slices.items():
        assert model.params[n].data_ptr() == model.flat[sl].data_ptr()
        assert model.params[n].grad.data_ptr() == model.flat_grad[sl].data_ptr()
    model.flat_grad.normal_(generator=torch.Generator().manual_seed(0))
    before = model.flat.clone()
    if fused:
        return  # the fused Adam kernel needs the GPU; the torch path below exercises the carried-over state
    model.adam_step()
    assert (model.flat != before).any()


def test_replicas_stay_identical_with_seeded_generators():
    """Two replicas (ranks) that densify with generators seeded alike end up bit-identical — what the
    multi-GPU fit needs instead of a broadcast."""
    z = np.load(FILES[0])
    max_grad, pct, min_opacity, extent, max_screen, percent_dense = (float(x) for x in z["hyper"])
    a, b = _model(z, True), _model(z, True)
    ca = a.densify_and_prune(max_grad, pct, min_opacity, extent, int(max_screen), generator=torch.Generator().manual_seed(7))
    cb = b.densify_and_prune(max_grad, pct, min_opacity, extent, int(max_screen), generator=torch.Generator().manual_seed(7))
    assert ca == cb and torch.equal(a.flat, b.flat) and torch.equal(a.exp_avg, b.exp_avg)
