"""TEST INFRASTRUCTURE — numpy front end of the CPU restatement (oracle/splat_oracle.c).

`forward()` runs the reference's stage sequence K1..K6 on the CPU and returns every
intermediate under the reference's names; `backward()` runs K7..K9; `apply_weights()` runs
K13/K2-K5/K14. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this module. Pinned against fixtures produced by the reference itself
(tests/golden/, oracle/make_golden.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


class OView(C.Structure):
    _fields_ = [("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("W", C.c_int), ("H", C.c_int),
                ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("scale_modifier", C.c_float),
                ("view", C.c_void_p), ("proj", C.c_void_p), ("campos", C.c_void_p), ("bg", C.c_void_p)]


def build():
    subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_scan.restype = C.c_uint32
        _lib.oracle_higher_msb.restype = C.c_uint32
    return _lib


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _view(P, D, M, W, H, tanfovx, tanfovy, scale_modifier, view, proj, campos, bg):
    keep = [_f32(view).reshape(-1), _f32(proj).reshape(-1), _f32(campos).reshape(-1), _f32(bg).reshape(-1)]
    v = OView(P, D, M, W, H, np.float32(tanfovx), np.float32(tanfovy), np.float32(scale_modifier),
              _p(keep[0]), _p(keep[1]), _p(keep[2]), _p(keep[3]))
    v._keep = keep
    return v


def forward(means3D, opacities, view, proj, campos, bg, W, H, tanfovx, tanfovy, shs=None, colors_precomp=None,
            scales=None, rotations=None, cov3D_precomp=None, sh_degree=3, scale_modifier=1.0, render=True):
    """CudaRasterizer::Rasterizer::forward (DGR/cuda_rasterizer/rasterizer_impl.cu:179-285) on the CPU."""
    lib = load()
    means3D, opacities = _f32(means3D), _f32(opacities).reshape(-1)
    shs, colors_precomp = _f32(shs), _f32(colors_precomp)
    scales, rotations, cov3D_precomp = _f32(scales), _f32(rotations), _f32(cov3D_precomp)
    P = means3D.shape[0]
    M = 0 if shs is None else shs.shape[1]
    v = _view(P, sh_degree, M, W, H, tanfovx, tanfovy, scale_modifier, view, proj, campos, bg)
    o = {
        "radii": np.zeros(P, np.int32), "means2D": np.zeros((P, 2), np.float32), "depths": np.zeros(P, np.float32),
        "cov3D": np.zeros((P, 6), np.float32), "conic_opacity": np.zeros((P, 4), np.float32),
        "rgb": np.zeros((P, 3), np.float32), "clamped": np.zeros((P, 3), np.uint8),
        "tiles_touched": np.zeros(P, np.uint32), "rect": np.zeros((P, 4), np.int32),
    }
    lib.oracle_preprocess(C.byref(v), _p(means3D), _p(scales), _p(rotations), _p(opacities), _p(shs),
                          _p(cov3D_precomp), _p(colors_precomp), _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]),
                          _p(o["cov3D"]), _p(o["conic_opacity"]), _p(o["rgb"]), _p(o["clamped"]),
                          _p(o["tiles_touched"]), _p(o["rect"]))
    if cov3D_precomp is not None:
        o["cov3D"] = cov3D_precomp.copy()
    o["point_offsets"] = np.zeros(P, np.uint32)
    R = int(lib.oracle_scan(P, _p(o["tiles_touched"]), _p(o["point_offsets"])))
    o["num_rendered"] = R
    gx, gy = (W + 15) // 16, (H + 15) // 16
    T = gx * gy
    keys, vals = np.zeros(max(R, 1), np.uint64), np.zeros(max(R, 1), np.uint32)
    lib.oracle_duplicate(P, W, H, _p(o["means2D"]), _p(o["depths"]), _p(o["point_offsets"]), _p(o["radii"]),
                         _p(keys), _p(vals))
    o["keys_unsorted"], o["point_list_unsorted"] = keys[:R].copy(), vals[:R].copy()
    bit = int(lib.oracle_higher_msb(C.c_uint32(T)))
    kt, vt = np.zeros_like(keys), np.zeros_like(vals)
    lib.oracle_sort_pairs(C.c_uint32(R), 32 + bit, _p(keys), _p(vals), _p(kt), _p(vt))
    o["keys"], o["point_list"] = keys[:R], vals[:R]
    o["ranges"] = np.zeros((T, 2), np.uint32)
    lib.oracle_tile_ranges(C.c_uint32(R), _p(keys), T, _p(o["ranges"]))
    if render:
        colors = colors_precomp if colors_precomp is not None else o["rgb"]
        o["out_color"] = np.zeros((3, H, W), np.float32)
        o["out_depth"] = np.zeros((1, H, W), np.float32)
        o["final_T"] = np.zeros((H, W), np.float32)
        o["n_contrib"] = np.zeros((H, W), np.uint32)
        lib.oracle_render_forward(W, H, _p(o["ranges"]), _p(vals), _p(o["means2D"]), _p(colors), _p(o["depths"]),
                                  _p(o["conic_opacity"]), _p(_f32(bg).reshape(-1)), _p(o["out_color"]),
                                  _p(o["out_depth"]), _p(o["final_T"]), _p(o["n_contrib"]))
    vis = o["radii"] > 0
    for k in ("depths", "clamped", "means2D", "cov3D", "conic_opacity", "rgb"):
        o[k][~vis] = 0
    return o


def backward(fw, dL_dpix, means3D, view, proj, campos, bg, W, H, tanfovx, tanfovy, shs=None, colors_precomp=None,
             scales=None, rotations=None, cov3D_precomp=None, sh_degree=3, scale_modifier=1.0):
    """CudaRasterizer::Rasterizer::backward (DGR/cuda_rasterizer/rasterizer_impl.cu:289-341) on the CPU,
    given `fw` = the dict returned by forward(). Returns the eight tensors of
    RasterizeGaussiansBackwardCUDA (DGR/rasterize_points.cu:155-156) plus dL_dconic."""
    lib = load()
    means3D, shs, colors_precomp = _f32(means3D), _f32(shs), _f32(colors_precomp)
    scales, rotations, cov3D_precomp = _f32(scales), _f32(rotations), _f32(cov3D_precomp)
    dL_dpix = _f32(dL_dpix)
    P = means3D.shape[0]
    M = 0 if shs is None else shs.shape[1]
    v = _view(P, sh_degree, M, W, H, tanfovx, tanfovy, scale_modifier, view, proj, campos, bg)
    colors = colors_precomp if colors_precomp is not None else fw["rgb"]
    d_mean2D, d_conic = np.zeros((P, 2), np.float64), np.zeros((P, 3), np.float64)
    d_opacity, d_color = np.zeros(P, np.float64), np.zeros((P, 3), np.float64)
    plist = np.ascontiguousarray(fw["point_list"]) if fw["num_rendered"] else np.zeros(1, np.uint32)
    lib.oracle_render_backward(W, H, _p(fw["ranges"]), _p(plist), _p(_f32(bg).reshape(-1)), _p(fw["means2D"]),
                               _p(fw["conic_opacity"]), _p(colors), _p(fw["final_T"]), _p(fw["n_contrib"]),
                               _p(dL_dpix), _p(d_mean2D), _p(d_conic), _p(d_opacity), _p(d_color))
    m2 = d_mean2D.astype(np.float32)
    con = d_conic.astype(np.float32)
    col = d_color.astype(np.float32)
    out = {
        "dL_dmeans3D": np.zeros((P, 3), np.float32), "dL_dcov3D": np.zeros((P, 6), np.float32),
        "dL_dsh": np.zeros((P, M, 3), np.float32), "dL_dscales": np.zeros((P, 3), np.float32),
        "dL_drotations": np.zeros((P, 4), np.float32),
    }
    cov3D = cov3D_precomp if cov3D_precomp is not None else fw["cov3D"]
    lib.oracle_geom_backward(C.byref(v), _p(means3D), _p(fw["radii"]), _p(shs), _p(fw["clamped"]), _p(scales),
                             _p(rotations), _p(np.ascontiguousarray(cov3D)), int(cov3D_precomp is not None), _p(m2),
                             _p(con), _p(col), _p(out["dL_dmeans3D"]), _p(out["dL_dcov3D"]),
                             _p(out["dL_dsh"]) if M else None, _p(out["dL_dscales"]), _p(out["dL_drotations"]))
    out["dL_dmeans2D"] = np.concatenate([m2, np.zeros((P, 1), np.float32)], axis=1)
    out["dL_dcolors"] = col
    out["dL_dopacity"] = d_opacity.astype(np.float32).reshape(P, 1)
    out["dL_dconic"] = con
    return out


def apply_weights(means3D, opacities, view, proj, campos, W, H, tanfovx, tanfovy, image_weights, weights, cnt,
                  scales=None, rotations=None, cov3D_precomp=None, scale_modifier=1.0):
    """Rasterizer::apply_weights (DGR/cuda_rasterizer/rasterizer_impl.cu:343-447) on the CPU.
    Returns updated (weights float64 [P,CH], cnt int64 [P]); inputs are not modified."""
    lib = load()
    image_weights = _f32(image_weights)
    CH = image_weights.shape[0]
    fw = forward(means3D, opacities, view, proj, campos, np.zeros(3, np.float32), W, H, tanfovx, tanfovy,
                 colors_precomp=np.zeros((np.asarray(means3D).shape[0], 3), np.float32), scales=scales,
                 rotations=rotations, cov3D_precomp=cov3D_precomp, sh_degree=0, scale_modifier=scale_modifier,
                 render=False)
    w = np.ascontiguousarray(np.asarray(weights, dtype=np.float64).reshape(-1, CH)).copy()
    c = np.ascontiguousarray(np.asarray(cnt, dtype=np.int64).reshape(-1)).copy()
    plist = np.ascontiguousarray(fw["point_list"]) if fw["num_rendered"] else np.zeros(1, np.uint32)
    lib.oracle_apply_weights_render(W, H, CH, _p(fw["ranges"]), _p(plist), _p(fw["means2D"]),
                                    _p(fw["conic_opacity"]), _p(image_weights), _p(w), _p(c))
    return w, c
