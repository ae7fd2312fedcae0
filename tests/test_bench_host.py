"""Host-side arithmetic of bench.py (no GPU): the algorithmic-byte figures behind `roofline.achieved`
must add up to SURVEY.md §8d's per-view totals, and the synthetic workload helpers must be deterministic."""
import importlib.util
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_algorithmic_bytes_add_up_to_the_survey_totals():
    b = _bench()
    P, Pv, R, N, T, n_pass = 1_000_000, 853_000, 3_840_000, 512 * 512, 1024, 6
    stage = lambda s: b.algorithmic_bytes(s, P, Pv, R, N, T, n_pass)
    # B_f = 36 P + 303 P_v + (72 + 24 n_pass) R + 24 N_pix + 24 T
    B_f = 36 * P + 303 * Pv + (72 + 24 * n_pass) * R + 24 * N + 24 * T
    assert stage("preprocess") + stage("binning") + stage("render_fwd") == B_f
    # B_b = 8 P + 659 P_v + 284 (P - P_v) + 40 R + 20 N_pix + 8 T
    B_b = 8 * P + 659 * Pv + 284 * (P - Pv) + 40 * R + 20 * N + 8 * T
    assert stage("render_bwd") + stage("geom_bwd") == B_b
    # the survey's worked example (P = P_v = 1 M, R = 4 M): 1.21 GB forward, 0.83 GB backward per view
    ex = lambda s: b.algorithmic_bytes(s, 1_000_000, 1_000_000, 4_000_000, N, T, 6)
    assert abs(ex("preprocess") + ex("binning") + ex("render_fwd") - 1.21e9) < 0.01e9
    assert abs(ex("render_bwd") + ex("geom_bwd") - 0.83e9) < 0.01e9


def test_configs_name_the_baseline_workloads():
    b = _bench()
    c2 = b.CONFIGS["config2"]
    assert (c2["P"], c2["W"], c2["H"], c2["V"]) == (1_000_000, 512, 512, 20)  # BASELINE.json configs[1]
    assert (b.CONFIGS["config4"]["W"], b.CONFIGS["config4"]["H"]) == (1264, 832)
    assert (b.CONFIGS["config5"]["W"], b.CONFIGS["config5"]["H"]) == (1920, 1080)
    assert b.STAGE_NAMES.index("render_bwd") == 4  # include/dge_b200.h: stage ids of dge_profile_enable


def test_synthetic_scene_is_seeded():
    from dge_b200 import scene
    a, c = scene.make_gaussians(1000, seed=7), scene.make_gaussians(1000, seed=7)
    for x, y in zip(a, c):
        assert torch.equal(x, y)
    cams = scene.ring_cameras(4, 64, 48)
    assert len(cams) == 4 and cams[0].image_width == 64 and cams[0].image_height == 48
