"""Gaussian point-cloud PLY files in the reference's on-disk layout.

GaussianModel.save_ply / load_ply (gaussiansplatting/scene/gaussian_model.py:396-445, :447-540) go through
the third-party `plyfile` package; the format they produce is a binary little-endian PLY with ONE element
`vertex` whose float32 properties are, in this order,
    x y z  nx ny nz  f_dc_0..2  f_rest_0..44  opacity  scale_0..2  rot_0..3
with RAW (pre-activation) opacity / scale, zero normals, and the SH features stored CHANNEL-major
(`_features_dc/_features_rest [P, coeffs, 3]` are transposed to `[P, 3, coeffs]` and flattened, :413-428;
loading reshapes to `[P, 3, coeffs]` and transposes back, :484-487, :516-521). This module writes and
reads exactly that with numpy — no plyfile, nothing from oracle/.
"""
import os
from typing import Dict

import numpy as np
import torch


def attribute_names(n_dc: int = 3, n_rest: int = 45):
    # construct_list_of_attributes, gaussian_model.py:396-408
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(n_dc)]
    names += [f"f_rest_{i}" for i in range(n_rest)]
    names += ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]
    return names


def save_ply(raw: Dict[str, torch.Tensor], path: str) -> None:
    """raw: the six RAW parameter tensors by group name (FitModel.params): xyz [P,3], f_dc [P,1,3],
    f_rest [P,K,3], opacity [P,1], scaling [P,3], rotation [P,4]."""
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    t = {k: v.detach().to(torch.float32).cpu() for k, v in raw.items()}
    P = t["xyz"].shape[0]
    f_dc = t["f_dc"].transpose(1, 2).flatten(start_dim=1).contiguous().numpy()
    f_rest = t["f_rest"].transpose(1, 2).flatten(start_dim=1).contiguous().numpy()
    xyz = t["xyz"].numpy()
    rows = np.concatenate((xyz, np.zeros_like(xyz), f_dc, f_rest, t["opacity"].reshape(P, 1).numpy(),
                           t["scaling"].numpy(), t["rotation"].numpy()), axis=1).astype("<f4")
    names = attribute_names(f_dc.shape[1], f_rest.shape[1])
    assert rows.shape[1] == len(names)
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {P}\n"
    header += "".join(f"property float {n}\n" for n in names) + "end_header\n"
    with open(path, "wb") as fh:
        fh.write(header.encode("ascii"))
        fh.write(np.ascontiguousarray(rows).tobytes())


def load_ply(path: str) -> Dict[str, torch.Tensor]:
    """Returns the six RAW parameter tensors (see save_ply); properties are looked up by NAME, the f_rest /
    scale / rot families sorted by their numeric suffix as the reference does (gaussian_model.py:470-500)."""
    with open(path, "rb") as fh:
        data = fh.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    lines = data[:end].decode("ascii").splitlines()
    if lines[0] != "ply" or "format binary_little_endian 1.0" not in lines:
        raise ValueError("load_ply: expected a binary little-endian PLY")
    P, props, in_vertex = 0, [], False
    for ln in lines:
        tok = ln.split()
        if tok[:1] == ["element"]:
            in_vertex = tok[1] == "vertex"
            if in_vertex:
                P = int(tok[2])
        elif tok[:1] == ["property"] and in_vertex:
            if tok[1] not in ("float", "float32"):
                raise ValueError(f"load_ply: property {tok[2]} is {tok[1]}, expected float")
            props.append(tok[2])
    rows = np.frombuffer(data, dtype="<f4", count=P * len(props), offset=end).reshape(P, len(props))
    col = {n: i for i, n in enumerate(props)}

    def family(prefix):
        names = sorted((n for n in props if n.startswith(prefix)), key=lambda x: int(x.split("_")[-1]))
        return rows[:, [col[n] for n in names]]

    f_dc = family("f_dc_")
    f_rest = family("f_rest_")
    out = {
        "xyz": rows[:, [col["x"], col["y"], col["z"]]],
        "f_dc": f_dc.reshape(P, 3, -1).transpose(0, 2, 1),
        "f_rest": f_rest.reshape(P, 3, -1).transpose(0, 2, 1),
        "opacity": rows[:, [col["opacity"]]],
        "scaling": family("scale_"),
        "rotation": family("rot_"),
    }
    return {k: torch.from_numpy(np.ascontiguousarray(v)).float() for k, v in out.items()}
