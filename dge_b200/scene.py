"""Synthetic inputs `randgauss-v1` (SURVEY.md §8d / BASELINE.md §3.3).

Seeded random Gaussians and ring cameras, generated with a CPU torch generator so that the
reference arm, the oracle and this implementation all see identical bits. Camera matrices
follow the conventions of the reference's cameras
(gaussiansplatting/scene/cameras.py:92-96, gaussiansplatting/utils/graphics_utils.py:297-344):
`world_view_transform` and `full_proj_transform` are the TRANSPOSED W2C / full projection,
row-major, which is what the rasterizer indexes as m[4*col+row].
"""
import math
from typing import NamedTuple

import numpy as np
import torch


class Gaussians(NamedTuple):
    means3D: torch.Tensor    # [P,3]
    scales: torch.Tensor     # [P,3]  (activated: exp)
    rotations: torch.Tensor  # [P,4]  (unit quaternions r,x,y,z)
    opacities: torch.Tensor  # [P,1]  (activated: in (0,1))
    shs: torch.Tensor        # [P,16,3]


class Camera(NamedTuple):
    image_height: int
    image_width: int
    FoVx: float
    FoVy: float
    world_view_transform: torch.Tensor  # [4,4]
    full_proj_transform: torch.Tensor   # [4,4]
    camera_center: torch.Tensor         # [3]


def make_gaussians(P, seed=1234, sh_degree=3, scale_median=0.012, scale_sigma=0.6):
    g = torch.Generator(device="cpu").manual_seed(seed)
    means = torch.randn(P, 3, generator=g)
    norm = means.norm(dim=1, keepdim=True).clamp_min(1e-12)
    means = means * torch.clamp(3.0 / norm, max=1.0)  # clip to ||x|| <= 3
    scales = torch.exp(math.log(scale_median) + scale_sigma * torch.randn(P, 3, generator=g))
    q = torch.randn(P, 4, generator=g)
    q = q / q.norm(dim=1, keepdim=True).clamp_min(1e-12)
    opac = 0.02 + 0.96 * torch.rand(P, 1, generator=g)
    M = (sh_degree + 1) ** 2
    shs = torch.empty(P, M, 3)
    shs[:, 0, :] = 0.6 * torch.randn(P, 3, generator=g)
    if M > 1:
        shs[:, 1:, :] = 0.08 * torch.randn(P, M - 1, 3, generator=g)
    return Gaussians(means.float().contiguous(), scales.float().contiguous(), q.float().contiguous(),
                     opac.float().contiguous(), shs.float().contiguous())


def _projection(znear, zfar, fovX, fovY):
    # graphics_utils.py:324-344 getProjectionMatrix
    tanY, tanX = math.tan(fovY / 2), math.tan(fovX / 2)
    top, right = tanY * znear, tanX * znear
    bottom, left = -top, -right
    P = torch.zeros(4, 4, dtype=torch.float32)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def look_at_camera(eye, width, height, fovy_deg=50.0, target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0),
                   znear=0.01, zfar=100.0):
    eye, target, up = (np.asarray(v, dtype=np.float64) for v in (eye, target, up))
    f = target - eye
    f /= np.linalg.norm(f)
    right = np.cross(-up, f)
    right /= np.linalg.norm(right)
    down = np.cross(f, right)
    R = np.stack([right, down, f], axis=1)  # C2W rotation (COLMAP axes: x right, y down, z forward)
    T = -R.T @ eye
    # graphics_utils.py:297-310 getWorld2View2 with translate=0, scale=1
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.T
    Rt[:3, 3] = T
    Rt[3, 3] = 1.0
    w2c = np.float32(np.linalg.inv(np.linalg.inv(Rt)))
    fovy = math.radians(fovy_deg)
    fovx = 2.0 * math.atan(math.tan(fovy / 2) * width / height)
    wvt = torch.tensor(w2c).transpose(0, 1).contiguous()
    proj = _projection(znear, zfar, fovx, fovy).transpose(0, 1)
    full = wvt.unsqueeze(0).bmm(proj.unsqueeze(0)).squeeze(0).float().contiguous()
    center = wvt.inverse()[3, :3].contiguous()
    return Camera(int(height), int(width), fovx, fovy, wvt, full, center)


def ring_cameras(V, width, height, radius=4.0, elevation_deg=15.0, fovy_deg=50.0):
    cams = []
    el = math.radians(elevation_deg)
    for k in range(V):
        az = 2.0 * math.pi * k / V
        eye = (radius * math.cos(el) * math.sin(az), radius * math.sin(el), radius * math.cos(el) * math.cos(az))
        cams.append(look_at_camera(eye, width, height, fovy_deg))
    return cams


def upstream_grad(width, height, seed):
    """dL/dcolor ~ N(0,1)/(3*N_pix)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(3, height, width, generator=g) / (3.0 * height * width)).float().contiguous()


def disc_mask(width, height, radius=160.0):
    """Binary 1-channel disc mask centred in the image (config 3)."""
    ys, xs = torch.meshgrid(torch.arange(height), torch.arange(width), indexing="ij")
    m = ((xs - (width - 1) / 2) ** 2 + (ys - (height - 1) / 2) ** 2) <= radius ** 2
    return m.float().unsqueeze(0).contiguous()


def raster_settings(cam, bg, sh_degree=3, scale_modifier=1.0, debug=False, module=None):
    """GaussianRasterizationSettings exactly as gaussian_renderer.render() builds them
    (gaussiansplatting/gaussian_renderer/__init__.py:72-88)."""
    if module is None:
        from . import diff_gaussian_rasterization as module
    return module.GaussianRasterizationSettings(
        image_height=int(cam.image_height), image_width=int(cam.image_width),
        tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5), bg=bg,
        scale_modifier=scale_modifier, viewmatrix=cam.world_view_transform,
        projmatrix=cam.full_proj_transform, sh_degree=sh_degree, campos=cam.camera_center,
        prefiltered=False, debug=debug)


def camera_to(cam, device, non_blocking=False):
    return cam._replace(world_view_transform=cam.world_view_transform.to(device, non_blocking=non_blocking),
                        full_proj_transform=cam.full_proj_transform.to(device, non_blocking=non_blocking),
                        camera_center=cam.camera_center.to(device, non_blocking=non_blocking))
