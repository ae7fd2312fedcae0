"""PLY files in the reference's on-disk layout (dge_b200/ply.py; gaussian_model.py:396-445, :447-540)."""
import numpy as np
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
