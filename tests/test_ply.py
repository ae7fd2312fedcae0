"""PLY files in the reference's on-disk layout (dge_b200/ply.py; gaussian_model.py:396-445, :447-540)."""
import os

import numpy as np
import pytest
import torch

from dge_b200 import fit, ply, scene


def test_header_and_layout(tmp_path):
    g = scene.make_gaussians(37, seed=3)
    model = fit.FitModel(g, torch.device("cpu"))
    path = str(tmp_path / "pc" / "point_cloud.ply")
    model.save_ply(path)
    data = open(path, "rb").read()
    end = data.index(b"end_header\n") + 11
    lines = data[:end].decode().splitlines()
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 37"]
    props = [ln.split()[2] for ln in lines if ln.startswith("property")]
    assert all(ln.split()[1] == "float" for ln in lines if ln.startswith("property"))
    assert props == (["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(45)]
                     + ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"])   # :396-408
    rows = np.frombuffer(data, "<f4", offset=end).reshape(37, 62)
    assert len(data) == end + 37 * 62 * 4
    p = {k: v.detach().numpy() for k, v in model.params.items()}
    assert np.array_equal(rows[:, 0:3], p["xyz"]) and not rows[:, 3:6].any()          # zero normals (:414)
    # channel-major SH: f_rest_j = coefficient (j % 15) of channel (j // 15)   (:423-430)
    assert np.array_equal(rows[:, 9 + 0 * 15 + 4], p["f_rest"][:, 4, 0])
    assert np.array_equal(rows[:, 9 + 2 * 15 + 7], p["f_rest"][:, 7, 2])
    assert np.array_equal(rows[:, 6:9], p["f_dc"][:, 0, :])
    assert np.array_equal(rows[:, 54], p["opacity"][:, 0])                               # raw (pre-sigmoid) opacity
    assert np.array_equal(rows[:, 55:58], p["scaling"]) and np.array_equal(rows[:, 58:62], p["rotation"])


def test_round_trip(tmp_path):
    g = scene.make_gaussians(101, seed=4)
    model = fit.FitModel(g, torch.device("cpu"))
    path = str(tmp_path / "a.ply")
    model.save_ply(path)
    raw = ply.load_ply(path)
    again = fit.FitModel.from_raw(raw, torch.device("cpu"))
    assert again.P == model.P
    for k in model.params:
        assert torch.equal(again.params[k], model.params[k]), k


REF = "/root/reference/gaussiansplatting"


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")
def test_files_interchange_with_the_reference_save_and_load(tmp_path):
    """The reference's OWN GaussianModel.save_ply / load_ply (gaussian_model.py:410-445, :455-551), loaded unchanged
    from the reference tree with `plyfile` replaced by tests/mini_plyfile.py (the package is not in this image):
    a file the reference writes loads here to the reference model's tensors, and a file written here loads in the
    reference to ours — bit for bit, features in the [P, coeffs, 3] layout both sides use in memory."""
    import sys
    import types
    from tests import mini_plyfile
    from tests.test_reference_callers import _cpu_for_cuda, _load, _stub
    saved = dict(sys.modules)
    try:
        _stub("gaussiansplatting"), _stub("gaussiansplatting.utils"), _stub("gaussiansplatting.scene")
        _load("gaussiansplatting.utils.general_utils", REF + "/utils/general_utils.py")
        _load("gaussiansplatting.utils.system_utils", REF + "/utils/system_utils.py")
        _load("gaussiansplatting.utils.sh_utils", REF + "/utils/sh_utils.py")
        sys.modules["plyfile"] = mini_plyfile
        _stub("simple_knn"), _stub("simple_knn._C", distCUDA2=None)
        _stub("gaussiansplatting.utils.graphics_utils", BasicPointCloud=object)
        _stub("gaussiansplatting.gaussian_renderer", camera2rasterizer=None)
        _stub("gaussiansplatting.knn", K_nearest_neighbors=None)
        GaussianModel = _load("gaussiansplatting.scene.gaussian_model", REF + "/scene/gaussian_model.py").GaussianModel
        g = scene.make_gaussians(53, seed=9)
        ours = fit.FitModel(g, torch.device("cpu"))
        with _cpu_for_cuda():
            # ours -> file -> the reference's load_ply
            path_a = str(tmp_path / "ours" / "point_cloud.ply")
            ours.save_ply(path_a)
            ref = GaussianModel(3, 0.0, 0.05, 1.0)
            ref.load_ply(path_a)
            assert ref.max_sh_degree == 3 and ref.active_sh_degree == 3
            pairs = [("xyz", ref._xyz), ("f_dc", ref._features_dc), ("f_rest", ref._features_rest), ("opacity", ref._opacity),
                     ("scaling", ref._scaling), ("rotation", ref._rotation)]
            for name, t in pairs:
                assert torch.equal(t.detach(), ours.params[name].detach()), name
            # the reference's save_ply -> file -> ours
            path_b = str(tmp_path / "ref" / "point_cloud.ply")
            ref.save_ply(path_b)
        raw = ply.load_ply(path_b)
        for name, t in pairs:
            assert torch.equal(raw[name].reshape(t.shape), t.detach()), name
        assert open(path_a, "rb").read() == open(path_b, "rb").read()   # and the two files are the same bytes
    finally:
        for k in list(sys.modules):
            if k not in saved:
                del sys.modules[k]
