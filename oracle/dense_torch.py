"""TEST INFRASTRUCTURE — dense torch-CPU evaluation of the splatting equations
(BASELINE.md §3.2).

No tiles, no sort-by-key, no hand-written backward: every (pixel, Gaussian) pair is
evaluated, the reference's hard decisions (near cull, tile-rect membership, power > 0,
alpha < 1/255, T*(1-alpha) < 1e-4) enter as constant masks and gradients come from
autograd. It is an independent derivation of what DGR/cuda_rasterizer/backward.cu computes
by hand, used to validate the C restatement and the CUDA backward, and as the "dense
torch-CPU" baseline named by BASELINE.json. Follows DGR/cuda_rasterizer/forward.cu:20-256
(per-Gaussian stage) and :338-377 (blend rule).
"""
import math

import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]


def _sh_color(deg, d, sh):
    """forward.cu:20-71; d = normalised direction [G,3], sh [G,M,3] -> [G,3] before +0.5"""
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    res = SH_C0 * sh[:, 0]
    if deg > 0:
        res = res - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        res = (res + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] + SH_C2[2] * (2.0 * zz - xx - yy) * sh[:, 6]
               + SH_C2[3] * xz * sh[:, 7] + SH_C2[4] * (xx - yy) * sh[:, 8])
    if deg > 2:
        res = (res + SH_C3[0] * y * (3.0 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10]
               + SH_C3[2] * y * (4.0 * zz - xx - yy) * sh[:, 11]
               + SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy) * sh[:, 12]
               + SH_C3[4] * x * (4.0 * zz - xx - yy) * sh[:, 13] + SH_C3[5] * z * (xx - yy) * sh[:, 14]
               + SH_C3[6] * x * (xx - 3.0 * yy) * sh[:, 15])
    return res


def render(means3D, opacities, viewmatrix, projmatrix, campos, bg, W, H, tanfovx, tanfovy, shs=None,
           colors_precomp=None, scales=None, rotations=None, sh_degree=3, scale_modifier=1.0, means2D=None,
           dtype=torch.float32, chunk=4096, dL_dcolor=None):
    """Returns dict(color [3,H,W], depth [1,H,W], final_T [H,W], n_contrib [H,W], radii [P]).
    If dL_dcolor is given, also accumulates .grad on every input that requires grad
    (chunk-wise backward of sum(color * dL_dcolor)), including `means2D` ([P,3] zeros): its
    gradient follows the reference's convention (NDC-scaled, DGR/cuda_rasterizer/backward.cu:460-461)."""
    t = lambda a: a.to(dtype)
    P = means3D.shape[0]
    V, Pm = t(viewmatrix), t(projmatrix)  # row-vector convention: p_row @ V
    ones = torch.ones(P, 1, dtype=dtype)
    x = t(means3D)
    p_view = torch.cat([x, ones], 1) @ V
    p_hom = torch.cat([x, ones], 1) @ Pm
    depth = p_view[:, 2]
    p_w = 1.0 / (p_hom[:, 3] + 0.0000001)
    p_proj = p_hom[:, :3] * p_w[:, None]
    # cov3D (forward.cu:118-152)
    s = scale_modifier * t(scales)
    q = t(rotations)
    r, qx, qy, qz = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    Rm = torch.stack([1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - r * qz), 2 * (qx * qz + r * qy),
                      2 * (qx * qy + r * qz), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - r * qx),
                      2 * (qx * qz - r * qy), 2 * (qy * qz + r * qx), 1 - 2 * (qx * qx + qy * qy)], 1).view(P, 3, 3)
    Mm = Rm * s[:, None, :]            # R_std @ diag(s)
    Sigma = Mm @ Mm.transpose(1, 2)
    # cov2D (forward.cu:74-113)
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    tz = p_view[:, 2]
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    # The reference's backward treats the clamped t.x / t.y as constants (x_grad_mul = 0, no
    # dependence on t.z through the clamp: DGR/cuda_rasterizer/backward.cu:175-176,262-264).
    txtz, tytz = p_view[:, 0] / tz, p_view[:, 1] / tz
    txc = torch.where((txtz >= -limx) & (txtz <= limx), p_view[:, 0], (torch.clamp(txtz, -limx, limx) * tz).detach())
    tyc = torch.where((tytz >= -limy) & (tytz <= limy), p_view[:, 1], (torch.clamp(tytz, -limy, limy) * tz).detach())
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -(fx * txc) / (tz * tz), zero, fy / tz, -(fy * tyc) / (tz * tz)], 1).view(P, 2, 3)
    Wr = V[:3, :3].T                    # true W2C rotation
    JW = J @ Wr
    cov = JW @ Sigma @ JW.transpose(1, 2)
    a, b, c = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = a * c - b * b
    conic = torch.stack([c / det, -b / det, a / det], 1)
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam)).detach()
    pix = torch.stack([((p_proj[:, 0].double() + 1.0) * W - 1.0) * 0.5, ((p_proj[:, 1].double() + 1.0) * H - 1.0) * 0.5], 1).to(dtype)
    if means2D is not None:  # gradient tap with the reference's scaling: d pix / d ndc = 0.5*W
        pix = pix + t(means2D)[:, :2] * torch.tensor([0.5 * W, 0.5 * H], dtype=dtype)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    pd = pix.detach()
    # auxiliary.h:46-56 in float32, as the rasterizer does
    p32, r32 = pd.float(), radius.float()
    rmin = torch.stack([((p32[:, 0] - r32) * 0.0625).trunc().clamp(0, gx), ((p32[:, 1] - r32) * 0.0625).trunc().clamp(0, gy)], 1)
    rmax = torch.stack([((((p32[:, 0] + r32) + 16.0) - 1.0) * 0.0625).trunc().clamp(0, gx),
                        ((((p32[:, 1] + r32) + 16.0) - 1.0) * 0.0625).trunc().clamp(0, gy)], 1)
    tiles = (rmax[:, 0] - rmin[:, 0]) * (rmax[:, 1] - rmin[:, 1])
    visible = (depth.detach() > 0.2) & (det.detach() != 0) & (tiles > 0)
    radii = torch.where(visible, radius, torch.zeros_like(radius)).to(torch.int32)
    # colour (forward.cu:20-71, :241-247)
    if colors_precomp is None:
        d = x - t(campos)[None, :]
        d = d / d.norm(dim=1, keepdim=True)
        rgb = torch.clamp_min(_sh_color(sh_degree, d, t(shs)) + 0.5, 0.0)
    else:
        rgb = t(colors_precomp)
    # global front-to-back order among visible Gaussians (ties by index: stable sort)
    vis_idx = torch.nonzero(visible).flatten()
    dbits = depth.detach().float()[vis_idx].view(torch.int32).to(torch.int64)
    order = vis_idx[torch.sort(dbits, stable=True).indices]
    G = order.numel()
    o_pix, o_con, o_op = pix[order], conic[order], t(opacities).reshape(-1)[order]
    o_rgb, o_dep = rgb[order], depth[order]
    o_rmin, o_rmax = rmin[order], rmax[order]
    bgv = t(bg)
    color = torch.zeros(3, H * W, dtype=dtype)
    depth_img = torch.zeros(H * W, dtype=dtype)
    final_T = torch.ones(H * W, dtype=dtype)
    n_contrib = torch.zeros(H * W, dtype=torch.int64)
    N = H * W
    for s0 in range(0, N, chunk):
        ids = torch.arange(s0, min(N, s0 + chunk))
        pxi, pyi = ids % W, ids // W
        pxf, pyf = pxi.to(dtype)[:, None], pyi.to(dtype)[:, None]
        tx_, ty_ = (pxi // 16).float()[:, None], (pyi // 16).float()[:, None]
        member = (tx_ >= o_rmin[None, :, 0]) & (tx_ < o_rmax[None, :, 0]) & (ty_ >= o_rmin[None, :, 1]) & (ty_ < o_rmax[None, :, 1])
        dx = o_pix[None, :, 0] - pxf
        dy = o_pix[None, :, 1] - pyf
        power = -0.5 * (o_con[None, :, 0] * dx * dx + o_con[None, :, 2] * dy * dy) - o_con[None, :, 1] * dx * dy
        alpha_raw = o_op[None, :] * torch.exp(power)
        alpha = torch.clamp(alpha_raw, max=0.99)
        # NB the reference back-propagates through min(0.99, .) as if unclamped
        alpha = alpha_raw + (alpha - alpha_raw).detach()
        keep = member & (power.detach() <= 0) & (alpha.detach() >= 1.0 / 255.0)
        a_eff = torch.where(keep, alpha, torch.zeros_like(alpha))
        one_m = 1.0 - a_eff
        T_incl = torch.cumprod(one_m, dim=1)
        T_excl = torch.cat([torch.ones(len(ids), 1, dtype=dtype), T_incl[:, :-1]], 1)
        stop = keep & (T_incl.detach() < 0.0001)           # first such entry terminates the pixel
        alive = (torch.cumsum(stop.to(torch.int32), 1) == 0)
        wgt = torch.where(alive, a_eff * T_excl, torch.zeros_like(a_eff))
        contrib = alive & keep
        # final T = product over contributing entries only
        Tf = torch.prod(torch.where(contrib, one_m, torch.ones_like(one_m)), dim=1)
        col = wgt @ o_rgb + Tf[:, None] * bgv[None, :]
        dep = wgt @ o_dep
        pos = torch.arange(1, G + 1)[None, :].expand(len(ids), G)
        # n_contrib counts positions in the TILE list; here the rank among tile members
        rank = torch.cumsum(member.to(torch.int64), 1)
        last = torch.where(contrib, rank, torch.zeros_like(rank)).max(dim=1).values if G else torch.zeros(len(ids), dtype=torch.int64)
        color[:, ids] = col.detach().T
        depth_img[ids] = dep.detach()
        final_T[ids] = Tf.detach()
        n_contrib[ids] = last
        if dL_dcolor is not None:
            gsel = t(dL_dcolor).reshape(3, N)[:, ids].T
            (col * gsel).sum().backward(retain_graph=True)
        del member, dx, dy, power, alpha, alpha_raw, keep, a_eff, one_m, T_incl, T_excl, stop, alive, wgt, contrib, pos, rank
    return {"color": color.view(3, H, W), "depth": depth_img.view(1, H, W), "final_T": final_T.view(H, W),
            "n_contrib": n_contrib.view(H, W), "radii": radii}
