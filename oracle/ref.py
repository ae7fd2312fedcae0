"""TEST INFRASTRUCTURE — ctypes driver of oracle/_ref/libref_rast.so.

Runs the UNMODIFIED reference rasterizer (rebuilt for sm_100a by oracle/Makefile from
/root/reference) the way its torch glue does (DGR/rasterize_points.cu): same allocations
(torch.full outputs, resizable byte buffers, nine zero-filled gradient tensors), same
argument order. Used by tests/ (live parity on the GPU box), oracle/make_golden.py and
bench.py's reference arm only; nothing under dge_b200/ imports this module.
The reference launches on the legacy default stream; callers must run it with torch's
default stream current.
"""
import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_rast.so")
ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_size_t)
_lib = None
_p, _i, _f = C.c_void_p, C.c_int, C.c_float


def available():
    return os.path.exists(LIB_PATH)


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB_PATH)
        lib.ref_last_error.restype = C.c_char_p
        lib.ref_forward.restype = _i
        lib.ref_forward.argtypes = [ALLOC_FN, ALLOC_FN, ALLOC_FN, _i, _i, _i, _p, _i, _i, _p, _p, _p, _p, _p, _f,
                                    _p, _p, _p, _p, _p, _f, _f, _i, _p, _p, _p, _i]
        lib.ref_backward.restype = _i
        lib.ref_backward.argtypes = [_i, _i, _i, _i, _p, _i, _i, _p, _p, _p, _p, _f, _p, _p, _p, _p, _p, _f, _f,
                                     _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i]
        lib.ref_apply_weights.restype = _i
        lib.ref_apply_weights.argtypes = [ALLOC_FN, ALLOC_FN, ALLOC_FN, _i, _i, _i, _p, _i, _i, _p, _p, _p, _p, _p,
                                          _f, _p, _p, _p, _p, _p, _f, _f, _i, _p, _p, _p, _i, _i]
        lib.ref_clock_probe.restype = None
        lib.ref_clock_probe.argtypes = [_p]
        lib.ref_mark_visible.restype = None
        lib.ref_mark_visible.argtypes = [_i, _p, _p, _p, _p]
        for n in ("ref_geom_pointers", "ref_binning_pointers", "ref_img_pointers"):
            getattr(lib, n).restype = None
            getattr(lib, n).argtypes = [_p, C.c_size_t, C.POINTER(_p)]
        _lib = lib
    return _lib


def _ptr(t):
    # the reference's glue passes data_ptr() of empty tensors, which is NULL
    return None if t is None or t.numel() == 0 else t.data_ptr()


class _Arena:
    def __init__(self, device):
        self.bufs = [torch.empty(0, dtype=torch.uint8, device=device) for _ in range(3)]
        self.cbs = [ALLOC_FN(self._make(i, device)) for i in range(3)]

    def _make(self, i, device):
        def alloc(n):
            self.bufs[i] = torch.empty(int(n), dtype=torch.uint8, device=device)  # resize_ of an empty tensor
            return self.bufs[i].data_ptr()
        return alloc


def rasterize_gaussians(bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, tan_fovx, tan_fovy, H, W, sh, degree, campos, prefiltered, debug):
    """RasterizeGaussiansCUDA (DGR/rasterize_points.cu:35-95)."""
    lib = load()
    dev = means3D.device
    P = means3D.shape[0]
    out_color = torch.full((3, H, W), 0.0, dtype=torch.float32, device=dev)
    out_depth = torch.full((1, H, W), 0.0, dtype=torch.float32, device=dev)
    radii = torch.full((P,), 0, dtype=torch.int32, device=dev)
    arena = _Arena(dev)
    rendered = 0
    if P != 0:
        M = sh.shape[1] if sh.numel() != 0 else 0
        rendered = lib.ref_forward(arena.cbs[0], arena.cbs[1], arena.cbs[2], P, degree, M, _ptr(bg), W, H,
                                   _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(opacity), _ptr(scales),
                                   scale_modifier, _ptr(rotations), _ptr(cov3D_precomp), _ptr(viewmatrix),
                                   _ptr(projmatrix), _ptr(campos), tan_fovx, tan_fovy, int(prefiltered),
                                   _ptr(out_color), _ptr(out_depth), _ptr(radii), int(debug))
        if rendered < 0:
            raise RuntimeError(lib.ref_last_error().decode())
    return rendered, out_color, out_depth, radii, arena.bufs[0], arena.bufs[1], arena.bufs[2]


def rasterize_gaussians_backward(bg, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp,
                                 viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, sh, degree, campos,
                                 geom, R, binning, img, debug):
    """RasterizeGaussiansBackwardCUDA (DGR/rasterize_points.cu:97-157)."""
    lib = load()
    dev = means3D.device
    P = means3D.shape[0]
    H, W = dL_dout_color.shape[1], dL_dout_color.shape[2]
    M = sh.shape[1] if sh.numel() != 0 else 0
    z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
    dL_dmeans3D, dL_dmeans2D, dL_dcolors, dL_dconic = z(P, 3), z(P, 3), z(P, 3), z(P, 2, 2)
    dL_dopacity, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations = z(P, 1), z(P, 6), z(P, M, 3), z(P, 3), z(P, 4)
    if P != 0:
        rc = lib.ref_backward(P, degree, M, R, _ptr(bg), W, H, _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(scales),
                              scale_modifier, _ptr(rotations), _ptr(cov3D_precomp), _ptr(viewmatrix),
                              _ptr(projmatrix), _ptr(campos), tan_fovx, tan_fovy, _ptr(radii), _ptr(geom),
                              _ptr(binning), _ptr(img), _ptr(dL_dout_color.contiguous()), _ptr(dL_dmeans2D),
                              _ptr(dL_dconic), _ptr(dL_dopacity), _ptr(dL_dcolors), _ptr(dL_dmeans3D),
                              _ptr(dL_dcov3D), _ptr(dL_dsh), _ptr(dL_dscales), _ptr(dL_drotations), int(debug))
        if rc < 0:
            raise RuntimeError(lib.ref_last_error().decode())
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations, dL_dconic


def apply_weights(bg, means3D, weights, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                  projmatrix, tan_fovx, tan_fovy, H, W, sh, degree, campos, prefiltered, image_weights, cnt, debug):
    """applyWeightsGaussiansCUDA (DGR/rasterize_points.cu:177-234)."""
    lib = load()
    dev = means3D.device
    P = means3D.shape[0]
    num_channels = image_weights.shape[0]
    radii = torch.full((P,), 0, dtype=torch.int32, device=dev)
    arena = _Arena(dev)
    if P != 0:
        M = sh.shape[1] if sh.numel() != 0 else 0
        rc = lib.ref_apply_weights(arena.cbs[0], arena.cbs[1], arena.cbs[2], P, degree, M, _ptr(bg), W, H,
                                   _ptr(means3D), _ptr(sh), _ptr(weights), _ptr(opacity), _ptr(scales),
                                   scale_modifier, _ptr(rotations), _ptr(cov3D_precomp), _ptr(viewmatrix),
                                   _ptr(projmatrix), _ptr(campos), tan_fovx, tan_fovy, int(prefiltered),
                                   _ptr(image_weights.contiguous()), _ptr(radii), _ptr(cnt), num_channels,
                                   int(debug))
        if rc < 0:
            raise RuntimeError(lib.ref_last_error().decode())


def mark_visible(means3D, viewmatrix, projmatrix):
    lib = load()
    P = means3D.shape[0]
    present = torch.full((P,), False, dtype=torch.bool, device=means3D.device)
    if P != 0:
        lib.ref_mark_visible(P, _ptr(means3D), _ptr(viewmatrix), _ptr(projmatrix), _ptr(present))
    return present


def _view(base_tensor, ptr, dtype, count):
    """Tensor view of `count` elements of `dtype` at device address `ptr` inside base_tensor."""
    off = ptr - base_tensor.data_ptr()
    nbytes = count * torch.empty(0, dtype=dtype).element_size()
    return base_tensor[off:off + nbytes].view(dtype)


def intermediates(P, R, H, W, geom, binning, img, radii):
    """Every intermediate of one reference forward, as CPU numpy arrays
    (layout from the reference's own fromChunk, DGR/cuda_rasterizer/rasterizer_impl.cu:135-175)."""
    lib = load()
    out = {}
    gp = (_p * 10)()
    lib.ref_geom_pointers(geom.data_ptr(), P, gp)
    out["depths"] = _view(geom, gp[0], torch.float32, P)
    out["clamped"] = _view(geom, gp[1], torch.uint8, 3 * P).view(P, 3)
    out["means2D"] = _view(geom, gp[3], torch.float32, 2 * P).view(P, 2)
    out["cov3D"] = _view(geom, gp[4], torch.float32, 6 * P).view(P, 6)
    out["conic_opacity"] = _view(geom, gp[5], torch.float32, 4 * P).view(P, 4)
    out["rgb"] = _view(geom, gp[6], torch.float32, 3 * P).view(P, 3)
    out["tiles_touched"] = _view(geom, gp[7], torch.int32, P)
    out["point_offsets"] = _view(geom, gp[9], torch.int32, P)
    if R > 0:
        bp = (_p * 4)()
        lib.ref_binning_pointers(binning.data_ptr(), R, bp)
        out["point_list"] = _view(binning, bp[0], torch.int32, R)
        out["point_list_unsorted"] = _view(binning, bp[1], torch.int32, R)
        out["keys"] = _view(binning, bp[2], torch.int64, R)
        out["keys_unsorted"] = _view(binning, bp[3], torch.int64, R)
    ip = (_p * 3)()
    N = H * W
    T = ((W + 15) // 16) * ((H + 15) // 16)
    lib.ref_img_pointers(img.data_ptr(), N, ip)
    out["final_T"] = _view(img, ip[0], torch.float32, N).view(H, W)
    out["n_contrib"] = _view(img, ip[1], torch.int32, N).view(H, W)
    out["ranges"] = _view(img, ip[2], torch.int32, 2 * T).view(T, 2)
    out["radii"] = radii
    res = {k: v.detach().cpu().numpy().copy() for k, v in out.items()}
    # entries of culled Gaussians are uninitialised memory in the reference: blank them
    vis = res["radii"] > 0
    for k in ("depths", "clamped", "means2D", "cov3D", "conic_opacity", "rgb"):
        res[k][~vis] = 0
    return res
