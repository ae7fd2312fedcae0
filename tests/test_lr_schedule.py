"""FitModel's learning-rate schedule against the reference's own get_expon_lr_func
(gaussiansplatting/utils/general_utils.py:29-62), values generated here from the reference file
(tests/golden/lr_schedule.json, made by the __main__ block below) so the test runs without /root/reference."""
import json
import os

import pytest
import torch

from dge_b200 import fit, scene

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "lr_schedule.json")


def _model(fused):
    g = scene.make_gaussians(64, seed=3)
    return fit.FitModel(g, torch.device("cpu"), fused_adam=fused)


def test_schedule_matches_reference_function():
    ref = json.load(open(GOLDEN))
    for case in ref["cases"]:
        m = _model(True)
        m.training_setup(case["max_steps"], position_lr_init=case["lr_init"], position_lr_final=case["lr_final"],
                         spatial_lr_scale=case["scale"])
        for step, want in zip(case["steps"], case["lrs"]):
            got = m.update_learning_rate(step)
            assert got == pytest.approx(want, rel=1e-12, abs=0.0), (case, step)
            assert m.lrs["xyz"] == got


def test_training_setup_rates_and_torch_optimizer_groups():
    m = _model(False)
    m.training_setup(1500, spatial_lr_scale=5.0)
    lrs = {g["name"]: g["lr"] for g in m.optimizer.param_groups}
    # gaussian_model.py:341-372 with OptimizationParams' defaults
    assert lrs == pytest.approx({"xyz": 0.00016 * 5.0, "f_dc": 0.0125, "f_rest": 0.0125 / 20.0, "opacity": 0.05,
                                 "scaling": 0.005, "rotation": 0.001})
    m.update_learning_rate(1500)
    lrs = {g["name"]: g["lr"] for g in m.optimizer.param_groups}
    assert lrs["xyz"] == pytest.approx(0.000016 * 5.0) and lrs["f_dc"] == 0.0125


if __name__ == "__main__":  # regenerate the fixture from the reference's file (this container only)
    import importlib.util
    import sys
    import types
    for name in ("numpy",):
        __import__(name)
    spec = importlib.util.spec_from_file_location(
        "ref_general_utils", "/root/reference/gaussiansplatting/utils/general_utils.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cases = []
    for lr_init, lr_final, scale, max_steps in [(0.00016, 0.000016, 1.0, 1500), (0.00016, 0.000016, 4.37, 3000),
                                                (0.0005, 0.00016, 0.5, 100)]:
        f = mod.get_expon_lr_func(lr_init=lr_init * scale, lr_final=lr_final * scale, lr_delay_mult=0.01,
                                  max_steps=max_steps)
        steps = [-1, 0, 1, 7, max_steps // 3, max_steps // 2, max_steps - 1, max_steps, max_steps + 50]
        cases.append(dict(lr_init=lr_init, lr_final=lr_final, scale=scale, max_steps=max_steps, steps=steps,
                          lrs=[float(f(s)) for s in steps]))
    json.dump({"source": "gaussiansplatting/utils/general_utils.py:get_expon_lr_func", "cases": cases},
              open(GOLDEN, "w"), indent=1)
    print("wrote", GOLDEN)
