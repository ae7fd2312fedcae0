"""Drop-in `diff_gaussian_rasterization` for DGE, backed by libdge_b200.so.

Mirrors the reference binding
(gaussiansplatting/submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py,
DGR/... below): same names, argument order, defaults and error behaviour, so that
gaussiansplatting/gaussian_renderer/__init__.py:render() and
GaussianModel.apply_weights (gaussiansplatting/scene/gaussian_model.py:817-832) run unchanged.
`_C` (a torch C++ extension in the reference) is replaced by ctypes calls into the C-ABI of
include/dge_b200.h; there is no fallback when the library is missing.
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from .. import _lib as L


def cpu_deep_copy_tuple(input_tuple):
    # DGR/diff_gaussian_rasterization/__init__.py:18-23
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    # DGR/diff_gaussian_rasterization/__init__.py:26-47
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


def _f32c(t):
    """What the reference's glue does to every tensor: .contiguous() (rasterize_points.cu:79-93)."""
    if t.dtype != torch.float32:
        raise RuntimeError("expected scalar type Float")  # data<float>() in the reference
    return t.contiguous()


def _check_means(means3D):
    # DGR/rasterize_points.cu:46-48
    if means3D.ndim != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")


def _forward_call(rs, means3D, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, sh):
    """_C.rasterize_gaussians (DGR/rasterize_points.cu:35-95): returns
    (num_rendered, color, depth, radii, geomBuffer, binningBuffer, imgBuffer)."""
    lib = L.load()
    _check_means(means3D)
    dev = means3D.device
    if dev.type != "cuda":
        raise RuntimeError("dge_b200: tensors must live on a CUDA device (no CPU fallback)")
    P = means3D.shape[0]
    H, W = int(rs.image_height), int(rs.image_width)
    M = sh.shape[1] if sh.numel() != 0 else 0
    means3D, colors_precomp, opacities = _f32c(means3D), _f32c(colors_precomp), _f32c(opacities)
    scales, rotations, cov3Ds_precomp, sh = _f32c(scales), _f32c(rotations), _f32c(cov3Ds_precomp), _f32c(sh)
    bg, view, proj, campos = _f32c(rs.bg), _f32c(rs.viewmatrix), _f32c(rs.projmatrix), _f32c(rs.campos)
    color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
    depth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
    radii = torch.empty((P,), dtype=torch.int32, device=dev)
    arena = L.arena(dev)
    with torch.cuda.device(dev):
        rc = lib.dge_rasterize_forward(
            arena.cbs[0], arena.cbs[1], arena.cbs[2], None, P, int(rs.sh_degree), M, L.ptr(bg), W, H,
            L.ptr(means3D), L.ptr(sh), L.ptr(colors_precomp), L.ptr(opacities), L.ptr(scales),
            float(rs.scale_modifier), L.ptr(rotations), L.ptr(cov3Ds_precomp), L.ptr(view), L.ptr(proj),
            L.ptr(campos), float(rs.tanfovx), float(rs.tanfovy), int(bool(rs.prefiltered)),
            L.ptr(color), L.ptr(depth), L.ptr(radii), int(bool(rs.debug)), L.stream_ptr(dev))
    bufs = L.take(arena)
    L.check(rc, "rasterize_gaussians")
    empty = torch.empty(0, dtype=torch.uint8, device=dev)
    geom, binning, img = (b if b is not None else empty for b in bufs)
    return rc, color, depth, radii, geom, binning, img


def _backward_call(rs, means3D, radii, colors_precomp, scales, rotations, cov3Ds_precomp, grad_out_color, sh,
                   geom, num_rendered, binning, img, need):
    """_C.rasterize_gaussians_backward (DGR/rasterize_points.cu:97-157). `need` says which of
    (colors_precomp, cov3D_precomp) gradients are wanted; outputs are written in full by the
    library, so nothing is zero-filled here (the reference fills nine tensors, :120-128)."""
    lib = L.load()
    dev = means3D.device
    P = means3D.shape[0]
    H, W = grad_out_color.shape[1], grad_out_color.shape[2]
    M = sh.shape[1] if sh.numel() != 0 else 0
    f32 = dict(dtype=torch.float32, device=dev)
    new = torch.zeros if P == 0 else torch.empty
    dL_dmeans3D = new((P, 3), **f32)
    dL_dmeans2D = new((P, 3), **f32)
    dL_dopacity = new((P, 1), **f32)
    dL_dsh = new((P, M, 3), **f32)
    dL_dscales = new((P, 3), **f32)
    dL_drotations = new((P, 4), **f32)
    dL_dcolors = new((P, 3), **f32) if need["colors"] else None
    dL_dcov3D = new((P, 6), **f32) if need["cov3D"] else None
    if P != 0:
        grad_out_color = _f32c(grad_out_color)
        bg, view, proj, campos = _f32c(rs.bg), _f32c(rs.viewmatrix), _f32c(rs.projmatrix), _f32c(rs.campos)
        arena = L.arena(dev, 1)
        with torch.cuda.device(dev):
            rc = lib.dge_rasterize_backward(
                arena.cbs[0], None, P, int(rs.sh_degree), M, int(num_rendered), L.ptr(bg), W, H,
                L.ptr(means3D), L.ptr(sh), L.ptr(colors_precomp), L.ptr(scales), float(rs.scale_modifier),
                L.ptr(rotations), L.ptr(cov3Ds_precomp), L.ptr(view), L.ptr(proj), L.ptr(campos),
                float(rs.tanfovx), float(rs.tanfovy), L.ptr(radii), L.ptr(geom), L.ptr(binning), L.ptr(img),
                L.ptr(grad_out_color), L.ptr(dL_dmeans2D), None, L.ptr(dL_dopacity), L.ptr(dL_dcolors),
                L.ptr(dL_dmeans3D), L.ptr(dL_dcov3D), L.ptr(dL_dsh), L.ptr(dL_dscales), L.ptr(dL_drotations),
                0, int(bool(rs.debug)), L.stream_ptr(dev))
        L.take(arena)  # the scratch of the blend-stage sums is dead once the call has been queued (stream-ordered free)
        L.check(rc, "rasterize_gaussians_backward")
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations


class _RasterizeGaussians(torch.autograd.Function):
    # DGR/diff_gaussian_rasterization/__init__.py:50-225
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        args = (raster_settings, means3D, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, sh)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)  # Copy them before they can be corrupted
            try:
                num_rendered, color, depth, radii, geomBuffer, binningBuffer, imgBuffer = _forward_call(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            num_rendered, color, depth, radii, geomBuffer, binningBuffer, imgBuffer = _forward_call(*args)

        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer)
        return color, radii, depth

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_depth):
        # grad_radii / grad_depth are ignored, as in the reference (:137,168): depth is forward-only
        num_rendered = ctx.num_rendered
        raster_settings = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer,
         imgBuffer) = ctx.saved_tensors
        need = {"colors": colors_precomp.numel() != 0, "cov3D": cov3Ds_precomp.numel() != 0}
        args = (raster_settings, means3D, radii, colors_precomp, scales, rotations, cov3Ds_precomp, grad_out_color,
                sh, geomBuffer, num_rendered, binningBuffer, imgBuffer, need)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                out = _backward_call(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            out = _backward_call(*args)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales,
         grad_rotations) = out
        # Gradients for absent (empty) inputs are dropped by autograd in the reference too.
        grad_sh = grad_sh if sh.numel() != 0 else None
        grad_scales = grad_scales if scales.numel() != 0 else None
        grad_rotations = grad_rotations if rotations.numel() != 0 else None
        return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities, grad_scales,
                grad_rotations, grad_cov3Ds_precomp, None)


class GaussianRasterizationSettings(NamedTuple):
    # DGR/diff_gaussian_rasterization/__init__.py:228-240
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def _empty():
    return torch.Tensor([]).to(torch.float32).to("cuda")


class GaussianRasterizer(nn.Module):
    # DGR/diff_gaussian_rasterization/__init__.py:243-364
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        # Mark visible points (based on frustum culling for camera) with a boolean
        with torch.no_grad():
            rs = self.raster_settings
            lib = L.load()
            P = positions.shape[0]
            positions = _f32c(positions)
            view, proj = _f32c(rs.viewmatrix), _f32c(rs.projmatrix)
            visible = torch.zeros((P,), dtype=torch.bool, device=positions.device)
            if P != 0:
                with torch.cuda.device(positions.device):
                    L.check(lib.dge_mark_visible(P, L.ptr(positions), L.ptr(view), L.ptr(proj), L.ptr(visible),
                                                 L.stream_ptr(positions.device)), "mark_visible")
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")

        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception(
                "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")

        if shs is None:
            shs = _empty()
        if colors_precomp is None:
            colors_precomp = _empty()
        if scales is None:
            scales = _empty()
        if rotations is None:
            rotations = _empty()
        if cov3D_precomp is None:
            cov3D_precomp = _empty()

        # Invoke C++/CUDA rasterization routine
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, raster_settings)

    def apply_weights(self, means3D, means2D, opacities, shs=None, weights=None, scales=None, rotations=None,
                      cov3Ds_precomp=None, cnt=None, image_weights=None):
        assert weights is not None
        assert cnt is not None
        assert image_weights is not None

        rs = self.raster_settings
        if shs is None:
            shs = _empty()
        if scales is None:
            scales = _empty()
        if rotations is None:
            rotations = _empty()
        if cov3Ds_precomp is None:
            cov3Ds_precomp = _empty()

        # _C.apply_weights (DGR/rasterize_points.cu:177-234): weights / cnt updated in place
        lib = L.load()
        _check_means(means3D)
        dev = means3D.device
        P = means3D.shape[0]
        H, W = int(rs.image_height), int(rs.image_width)
        M = shs.shape[1] if shs.numel() != 0 else 0
        num_channels = image_weights.shape[0]
        if P == 0:
            return
        if not (weights.is_contiguous() and cnt.is_contiguous()):
            raise RuntimeError("apply_weights: weights and cnt must be contiguous (they are updated in place)")
        if weights.dtype != torch.float32 or cnt.dtype != torch.int32:
            raise RuntimeError("apply_weights: weights must be float32 and cnt int32")
        means3D, opacities, scales = _f32c(means3D), _f32c(opacities), _f32c(scales)
        rotations, cov3Ds_precomp, image_weights = _f32c(rotations), _f32c(cov3Ds_precomp), _f32c(image_weights)
        bg, view, proj, campos = _f32c(rs.bg), _f32c(rs.viewmatrix), _f32c(rs.projmatrix), _f32c(rs.campos)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        arena = L.arena(dev)
        with torch.cuda.device(dev):
            rc = lib.dge_apply_weights(
                arena.cbs[0], arena.cbs[1], arena.cbs[2], None, P, int(rs.sh_degree), M, L.ptr(bg), W, H,
                L.ptr(means3D), L.ptr(shs), L.ptr(weights), L.ptr(opacities), L.ptr(scales),
                float(rs.scale_modifier), L.ptr(rotations), L.ptr(cov3Ds_precomp), L.ptr(view), L.ptr(proj),
                L.ptr(campos), float(rs.tanfovx), float(rs.tanfovy), int(bool(rs.prefiltered)),
                L.ptr(image_weights), L.ptr(radii), L.ptr(cnt), int(num_channels), int(bool(rs.debug)),
                L.stream_ptr(dev))
        L.take(arena)
        L.check(rc, "apply_weights")
