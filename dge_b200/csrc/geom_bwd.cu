// K8 + K9 + K10 fused: per-Gaussian backward.
//
// Replaces computeCov2DCUDA (DGR/cuda_rasterizer/backward.cu:144-274), the backward
// preprocessCUDA with its SH and cov3D helpers (backward.cu:20-139, :278-396) and the nine
// torch::zeros fills of the glue (DGR/rasterize_points.cu:120-128): one pass reads the
// blend-stage sums acc[P][12] and writes EVERY output row (zeros for culled Gaussians), so
// no separate zero-fill and no second read of the per-Gaussian inputs is needed. cov3D is
// recomputed from scale/rotation (bit-identical to the forward) instead of being stored.
#include "common.cuh"
#include "math_ref.cuh"

namespace dge {

// GLM-style column-major 3x3: m[col][row]; product as glm::operator* (type_mat3x3.inl)
struct M3 {
  float m[3][3];
};
__device__ __forceinline__ M3 mul(const M3& A, const M3& B) {
  M3 R;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      R.m[i][j] = A.m[0][j] * B.m[i][0] + A.m[1][j] * B.m[i][1] + A.m[2][j] * B.m[i][2];
  return R;
}

__device__ __forceinline__ void store3(float* p, size_t idx, float a, float b, float c) {
  p[3 * idx] = a;
  p[3 * idx + 1] = b;
  p[3 * idx + 2] = c;
}
// ACC = true: add into the caller's running sums (fit step: gradients of all views of a step
// accumulate in place, no per-view tensors and no separate add pass); false: overwrite.
template <bool ACC>
__device__ __forceinline__ void out3(float* p, size_t idx, float a, float b, float c) {
  if (ACC) {
    p[3 * idx] += a;
    p[3 * idx + 1] += b;
    p[3 * idx + 2] += c;
  } else {
    store3(p, idx, a, b, c);
  }
}
template <bool ACC>
__device__ __forceinline__ void out4(float* p, size_t idx, float4 v) {
  float4* q = reinterpret_cast<float4*>(p) + idx;
  if (ACC) {
    const float4 o = *q;
    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
  }
  *q = v;
}


// ---- K8 + projection part of K9 for ONE view (backward.cu:144-274, :366-387): given the
// blend-stage sums of this Gaussian, the view's camera and the Gaussian's cov3D, returns the
// view's contribution to dL/dmean3D (without the SH term) and dL/dcov3D.
// The blend backward (render_bwd.cu) leaves MOMENT sums over the Gaussian's pixels, with q = dL/dG * G and
// d = centre - pixel: S = (sum q dx, sum q dy, sum q dx^2, sum q dx dy, sum q dy^2). The reference's
// per-pair expressions (backward.cu:537-551) are linear in them:
//   dL/dmean2D.x = -W/2 (cx Sx + cy Sy),  dL/dmean2D.y = -H/2 (cz Sy + cy Sx),
//   dL/dconic    = -1/2 (Sxx, Sxy, Syy)
// with the conic (cx, cy, cz) recomputed here bit-identically to the forward's (math_ref.cuh:conic_ref).
// dblend returns (dL/dmean2D.x, .y, dL/dconic.x, .y, .w).
__device__ __forceinline__ void view_geom_backward(float mx, float my, float mz, const float* c3,
                                                   const float* V, const float* Pm, float tan_fovx,
                                                   float tan_fovy, float focal_x, float focal_y, int W, int H,
                                                   float Sx, float Sy, float Sxx, float Sxy, float Syy,
                                                   float dblend[5], float dmean[3], float dcov[6]) {
  struct { float tan_fovx, tan_fovy, focal_x, focal_y; } vp = {tan_fovx, tan_fovy, focal_x, focal_y};
  // ---------------- K8: conic -> cov2D -> cov3D, mean (backward.cu:144-274) ----------------
  // The forward's own T = W J, T Sigma and cov2D (math_ref.cuh, bit-identical to preprocess) serve both the
  // conic that turns the moment sums into dL/dmean2D and the backward, which the reference recomputes them for.
  const float txe = xform_row(V, 0, mx, my, mz), tye = xform_row(V, 1, mx, my, mz);
  const float tz = xform_row(V, 2, mx, my, mz);
  float Tf[6], Af[6];
  const float3 cov = cov2d_parts(txe, tye, tz, tan_fovx, tan_fovy, focal_x, focal_y, V, c3, Tf, Af);
  const float a = cov.x, b = cov.y, c = cov.z;
  const float det = FMA(a, c, -MUL(b, b));
  const float inv = __frcp_rn(det);
  const float3 con = make_float3(MUL(c, inv), MUL(b, -inv), MUL(a, inv));  // preprocess_view
  const float dm2x = (-0.5f * (float)W) * (con.x * Sx + con.y * Sy);
  const float dm2y = (-0.5f * (float)H) * (con.z * Sy + con.y * Sx);
  const float dcon_x = -0.5f * Sxx, dcon_y = -0.5f * Sxy, dcon_w = -0.5f * Syy;
  dblend[0] = dm2x; dblend[1] = dm2y; dblend[2] = dcon_x; dblend[3] = dcon_y; dblend[4] = dcon_w;
  const float limx = 1.3f * vp.tan_fovx, limy = 1.3f * vp.tan_fovy;
  const float txtz = txe / tz, tytz = tye / tz;
  const float tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
  const float ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
  const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
  const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
  const float hx = vp.focal_x, hy = vp.focal_y;
  M3 Wm;
  Wm.m[0][0] = V[0]; Wm.m[0][1] = V[4]; Wm.m[0][2] = V[8];
  Wm.m[1][0] = V[1]; Wm.m[1][1] = V[5]; Wm.m[1][2] = V[9];
  Wm.m[2][0] = V[2]; Wm.m[2][1] = V[6]; Wm.m[2][2] = V[10];
  // TV[j][k] = sum_l T[j][l] Sigma[l][k]
  const float TV[2][3] = {{Af[0], Af[1], Af[2]}, {Af[3], Af[4], Af[5]}};

  const float denom = a * c - b * b;
  float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
  const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
#pragma unroll
  for (int k = 0; k < 6; k++) dcov[k] = 0.f;
  const float T00 = Tf[0], T01 = Tf[1], T02 = Tf[2];
  const float T10 = Tf[3], T11 = Tf[4], T12 = Tf[5];
  if (denom2inv != 0.f) {
    dL_da = denom2inv * (-c * c * dcon_x + 2 * b * c * dcon_y + (denom - a * c) * dcon_w);
    dL_dc = denom2inv * (-a * a * dcon_w + 2 * a * b * dcon_y + (denom - a * c) * dcon_x);
    dL_db = denom2inv * 2 * (b * c * dcon_x - (denom + 2 * b * b) * dcon_y + a * b * dcon_w);
    dcov[0] = T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc;
    dcov[3] = T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc;
    dcov[5] = T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc;
    dcov[1] = 2 * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2 * T10 * T11 * dL_dc;
    dcov[2] = 2 * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2 * T10 * T12 * dL_dc;
    dcov[4] = 2 * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2 * T11 * T12 * dL_dc;
  }
  // dL/dT (upper 2x3), TV[j][k] = sum_l T[j][l] Vrk[l][k]
  const float dT00 = 2 * TV[0][0] * dL_da + TV[1][0] * dL_db;
  const float dT01 = 2 * TV[0][1] * dL_da + TV[1][1] * dL_db;
  const float dT02 = 2 * TV[0][2] * dL_da + TV[1][2] * dL_db;
  const float dT10 = 2 * TV[1][0] * dL_dc + TV[0][0] * dL_db;
  const float dT11 = 2 * TV[1][1] * dL_dc + TV[0][1] * dL_db;
  const float dT12 = 2 * TV[1][2] * dL_dc + TV[0][2] * dL_db;
  const float dJ00 = Wm.m[0][0] * dT00 + Wm.m[0][1] * dT01 + Wm.m[0][2] * dT02;
  const float dJ02 = Wm.m[2][0] * dT00 + Wm.m[2][1] * dT01 + Wm.m[2][2] * dT02;
  const float dJ11 = Wm.m[1][0] * dT10 + Wm.m[1][1] * dT11 + Wm.m[1][2] * dT12;
  const float dJ12 = Wm.m[2][0] * dT10 + Wm.m[2][1] * dT11 + Wm.m[2][2] * dT12;
  const float itz = 1.f / tz, itz2 = itz * itz, itz3 = itz2 * itz;
  const float dtx = x_grad_mul * -hx * itz2 * dJ02;
  const float dty = y_grad_mul * -hy * itz2 * dJ12;
  const float dtz = -hx * itz2 * dJ00 - hy * itz2 * dJ11 + (2 * hx * tx) * itz3 * dJ02 +
                    (2 * hy * ty) * itz3 * dJ12;
  // transformVec4x3Transpose (auxiliary.h:89-97)
  dmean[0] = V[0] * dtx + V[1] * dty + V[2] * dtz;
  dmean[1] = V[4] * dtx + V[5] * dty + V[6] * dtz;
  dmean[2] = V[8] * dtx + V[9] * dty + V[10] * dtz;

  // ---------------- K9: projection (backward.cu:366-387) ----------------
  {
    const float mhw = Pm[3] * mx + Pm[7] * my + Pm[11] * mz + Pm[15];
    const float m_w = 1.0f / (mhw + 0.0000001f);
    const float mul1 = (Pm[0] * mx + Pm[4] * my + Pm[8] * mz + Pm[12]) * m_w * m_w;
    const float mul2 = (Pm[1] * mx + Pm[5] * my + Pm[9] * mz + Pm[13]) * m_w * m_w;
    dmean[0] += (Pm[0] * m_w - Pm[3] * mul1) * dm2x + (Pm[1] * m_w - Pm[3] * mul2) * dm2y;
    dmean[1] += (Pm[4] * m_w - Pm[7] * mul1) * dm2x + (Pm[5] * m_w - Pm[7] * mul2) * dm2y;
    dmean[2] += (Pm[8] * m_w - Pm[11] * mul1) * dm2x + (Pm[9] * m_w - Pm[11] * mul2) * dm2y;
  }

}

// SH basis b_k(dir) and its gradient w.r.t. the (normalised) direction, degree <= D
// (the coefficients of DGR/cuda_rasterizer/backward.cu:44-128, regrouped per basis function).
__device__ __forceinline__ void sh_basis_grad(int D, float x, float y, float z, float b[16],
                                              float gx[16], float gy[16], float gz[16]) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
#pragma unroll
    for (int k = 0; k < 16; k++) b[k] = gx[k] = gy[k] = gz[k] = 0.f;
    b[0] = SH_C0;
    if (D > 0) {
      b[1] = -SH_C1 * y; b[2] = SH_C1 * z; b[3] = -SH_C1 * x;
      gy[1] = -SH_C1; gz[2] = SH_C1; gx[3] = -SH_C1;
    }
    if (D > 1) {
      b[4] = SH_C2_0 * xy; b[5] = SH_C2_1 * yz; b[6] = SH_C2_2 * (2.f * zz - xx - yy);
      b[7] = SH_C2_3 * xz; b[8] = SH_C2_4 * (xx - yy);
      gx[4] = SH_C2_0 * y; gy[4] = SH_C2_0 * x;
      gy[5] = SH_C2_1 * z; gz[5] = SH_C2_1 * y;
      gx[6] = SH_C2_2 * 2.f * -x; gy[6] = SH_C2_2 * 2.f * -y; gz[6] = SH_C2_2 * 2.f * 2.f * z;
      gx[7] = SH_C2_3 * z; gz[7] = SH_C2_3 * x;
      gx[8] = SH_C2_4 * 2.f * x; gy[8] = SH_C2_4 * 2.f * -y;
    }
    if (D > 2) {
      b[9] = SH_C3_0 * y * (3.f * xx - yy); b[10] = SH_C3_1 * xy * z;
      b[11] = SH_C3_2 * y * (4.f * zz - xx - yy); b[12] = SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy);
      b[13] = SH_C3_4 * x * (4.f * zz - xx - yy); b[14] = SH_C3_5 * z * (xx - yy);
      b[15] = SH_C3_6 * x * (xx - 3.f * yy);
      gx[9] = SH_C3_0 * 3.f * 2.f * xy; gy[9] = SH_C3_0 * 3.f * (xx - yy);
      gx[10] = SH_C3_1 * yz; gy[10] = SH_C3_1 * xz; gz[10] = SH_C3_1 * xy;
      gx[11] = SH_C3_2 * -2.f * xy; gy[11] = SH_C3_2 * (-3.f * yy + 4.f * zz - xx); gz[11] = SH_C3_2 * 4.f * 2.f * yz;
      gx[12] = SH_C3_3 * -3.f * 2.f * xz; gy[12] = SH_C3_3 * -3.f * 2.f * yz; gz[12] = SH_C3_3 * 3.f * (2.f * zz - xx - yy);
      gx[13] = SH_C3_4 * (-3.f * xx + 4.f * zz - yy); gy[13] = SH_C3_4 * -2.f * xy; gz[13] = SH_C3_4 * 4.f * 2.f * xz;
      gx[14] = SH_C3_5 * 2.f * xz; gy[14] = SH_C3_5 * -2.f * yz; gz[14] = SH_C3_5 * (xx - yy);
      gx[15] = SH_C3_6 * 3.f * (xx - yy); gy[15] = SH_C3_6 * -3.f * 2.f * xy;
    }
}

// cov3D -> scale, quaternion backward (backward.cu:278-341); no quaternion-norm Jacobian.
__device__ __forceinline__ void cov3d_backward(float4 q, const float* sc, float scale_modifier,
                                               const float* dcov, float dscale[3], float4& dq) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    M3 R;
    R.m[0][0] = 1.f - 2.f * (y * y + z * z); R.m[0][1] = 2.f * (x * y - r * z); R.m[0][2] = 2.f * (x * z + r * y);
    R.m[1][0] = 2.f * (x * y + r * z); R.m[1][1] = 1.f - 2.f * (x * x + z * z); R.m[1][2] = 2.f * (y * z - r * x);
    R.m[2][0] = 2.f * (x * z - r * y); R.m[2][1] = 2.f * (y * z + r * x); R.m[2][2] = 1.f - 2.f * (x * x + y * y);
    const float s[3] = {scale_modifier * sc[0], scale_modifier * sc[1], scale_modifier * sc[2]};
    M3 Mm;  // M = S * R : M[col i][row j] = s_j * R[i][j]
#pragma unroll
    for (int ci = 0; ci < 3; ci++)
#pragma unroll
      for (int rj = 0; rj < 3; rj++) Mm.m[ci][rj] = s[rj] * R.m[ci][rj];
    M3 dS;
    dS.m[0][0] = dcov[0]; dS.m[0][1] = 0.5f * dcov[1]; dS.m[0][2] = 0.5f * dcov[2];
    dS.m[1][0] = 0.5f * dcov[1]; dS.m[1][1] = dcov[3]; dS.m[1][2] = 0.5f * dcov[4];
    dS.m[2][0] = 0.5f * dcov[2]; dS.m[2][1] = 0.5f * dcov[4]; dS.m[2][2] = dcov[5];
    M3 dM = mul(Mm, dS);
#pragma unroll
    for (int ci = 0; ci < 3; ci++)
#pragma unroll
      for (int rj = 0; rj < 3; rj++) dM.m[ci][rj] *= 2.0f;
    // Rt[k] = (R[0][k], R[1][k], R[2][k]); dMt[k] = (dM[0][k], dM[1][k], dM[2][k])
    float dMt[3][3];
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
      for (int l = 0; l < 3; l++) dMt[k][l] = dM.m[l][k];
#pragma unroll
    for (int k = 0; k < 3; k++)
      dscale[k] = R.m[0][k] * dMt[k][0] + R.m[1][k] * dMt[k][1] + R.m[2][k] * dMt[k][2];
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
      for (int l = 0; l < 3; l++) dMt[k][l] *= s[k];
    dq.x = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
    dq.y = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
    dq.z = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
    dq.w = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
}

template <bool ACC>
__global__ void __launch_bounds__(256) geom_backward_kernel(
    ViewParams vp, const float* __restrict__ means3D, const float* __restrict__ scales,
    const float* __restrict__ rotations, const float* __restrict__ shs,
    const float* __restrict__ cov3D_precomp, const int* __restrict__ radii,
    const uint8_t* __restrict__ clamped, const float* __restrict__ acc,
    float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconic, float* __restrict__ dL_dopacity,
    float* __restrict__ dL_dcolor, float* __restrict__ dL_dmean3D, float* __restrict__ dL_dcov3D,
    float* __restrict__ dL_dsh, float* __restrict__ dL_dscale, float* __restrict__ dL_drot) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= vp.P) return;
  const size_t i = (size_t)idx;
  const int M = vp.M;

  if (!(radii[idx] > 0)) {
    if (ACC) return;  // nothing to add
    // rows the reference leaves at torch::zeros
    store3(dL_dmean2D, i, 0.f, 0.f, 0.f);
    store3(dL_dmean3D, i, 0.f, 0.f, 0.f);
    if (dL_dopacity) dL_dopacity[i] = 0.f;
    if (dL_dconic) reinterpret_cast<float4*>(dL_dconic)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dL_dcolor) store3(dL_dcolor, i, 0.f, 0.f, 0.f);
    if (dL_dcov3D)
      for (int k = 0; k < 6; k++) dL_dcov3D[6 * i + k] = 0.f;
    if (dL_dsh) {
      if (M == 16) {
#pragma unroll
        for (int j = 0; j < 12; j++)
          reinterpret_cast<float4*>(dL_dsh + 48 * i)[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        for (int k = 0; k < 3 * M; k++) dL_dsh[3 * M * i + k] = 0.f;
      }
    }
    if (dL_dscale) store3(dL_dscale, i, 0.f, 0.f, 0.f);
    if (dL_drot) reinterpret_cast<float4*>(dL_drot)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }

  const float4 a0 = __ldg(reinterpret_cast<const float4*>(acc) + 3 * i);
  const float4 a1 = __ldg(reinterpret_cast<const float4*>(acc) + 3 * i + 1);
  const float4 a2 = __ldg(reinterpret_cast<const float4*>(acc) + 3 * i + 2);
  const float dop = a1.y;
  float dcol[3] = {a1.z, a1.w, a2.x};

  float V[16], Pm[16];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    V[k] = __ldg(vp.view + k);
    Pm[k] = __ldg(vp.proj + k);
  }
  const float mx = __ldg(means3D + 3 * i), my = __ldg(means3D + 3 * i + 1),
              mz = __ldg(means3D + 3 * i + 2);

  // ---------------- cov3D (forward value) ----------------
  float c3[6];
  float4 q = make_float4(0, 0, 0, 0);
  float sc[3] = {0, 0, 0};
  if (cov3D_precomp != nullptr) {
#pragma unroll
    for (int k = 0; k < 6; k++) c3[k] = __ldg(cov3D_precomp + 6 * i + k);
  } else {
    q = __ldg(reinterpret_cast<const float4*>(rotations) + i);
    sc[0] = __ldg(scales + 3 * i);
    sc[1] = __ldg(scales + 3 * i + 1);
    sc[2] = __ldg(scales + 3 * i + 2);
    cov3d_from_scale_rot(sc[0], sc[1], sc[2], vp.scale_modifier, q, c3);
  }

  // ---------------- K8 + projection (backward.cu:144-274, :366-387) ----------------
  float dmean[3], dcov[6], db[5];
  view_geom_backward(mx, my, mz, c3, V, Pm, vp.tan_fovx, vp.tan_fovy, vp.focal_x, vp.focal_y, vp.W, vp.H, a0.x, a0.y,
                     a0.z, a0.w, a1.x, db, dmean, dcov);
  out3<ACC>(dL_dmean2D, i, db[0], db[1], 0.f);
  if (dL_dopacity) {
    if (ACC) dL_dopacity[i] += dop; else dL_dopacity[i] = dop;
  }
  if (dL_dconic) out4<ACC>(dL_dconic, i, make_float4(db[2], db[3], 0.f, db[4]));
  if (dL_dcolor) out3<ACC>(dL_dcolor, i, dcol[0], dcol[1], dcol[2]);
  if (dL_dcov3D)
#pragma unroll
    for (int k = 0; k < 6; k++) {
      if (ACC) dL_dcov3D[6 * i + k] += dcov[k]; else dL_dcov3D[6 * i + k] = dcov[k];
    }

  // ---------------- K9: SH (backward.cu:20-139) ----------------
  // colour = sum_k b_k(dir) * sh_k, so dL/dsh_k = b_k * dL/dRGB and
  // dL/ddir = sum_k grad(b_k) * (sh_k . dL/dRGB): the reference's dRGBdx/dy/dz sums, regrouped
  // so that the 48 SH floats are consumed as they arrive from 12 vector loads.
  if (shs != nullptr && dL_dsh != nullptr) {
    const float* sh = shs + 3 * (size_t)M * i;
    float* dsh = dL_dsh + 3 * (size_t)M * i;
    const float ox = mx - __ldg(vp.campos), oy = my - __ldg(vp.campos + 1),
                oz = mz - __ldg(vp.campos + 2);
    const float len = sqrtf(ox * ox + oy * oy + oz * oz);
    const float x = ox / len, y = oy / len, z = oz / len;
    const uint8_t cl = clamped[i];
    float dRGB[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) dRGB[ch] = (cl >> ch) & 1 ? 0.f : dcol[ch];
    const int D = vp.D;
    float b[16], gx[16], gy[16], gz[16];
    sh_basis_grad(D, x, y, z, b, gx, gy, gz);
    float ddx = 0.f, ddy = 0.f, ddz = 0.f;
    if (M == 16) {
      float4 v[12];
#pragma unroll
      for (int j = 0; j < 12; j++) v[j] = __ldg(reinterpret_cast<const float4*>(sh) + j);
      const float* f = reinterpret_cast<const float*>(v);
#pragma unroll
      for (int k = 1; k < 16; k++) {
        const float sk = f[3 * k] * dRGB[0] + f[3 * k + 1] * dRGB[1] + f[3 * k + 2] * dRGB[2];
        ddx += gx[k] * sk;
        ddy += gy[k] * sk;
        ddz += gz[k] * sk;
      }
#pragma unroll
      for (int j = 0; j < 12; j++) {
        float4 o;
        o.x = b[(4 * j) / 3] * dRGB[(4 * j) % 3];
        o.y = b[(4 * j + 1) / 3] * dRGB[(4 * j + 1) % 3];
        o.z = b[(4 * j + 2) / 3] * dRGB[(4 * j + 2) % 3];
        o.w = b[(4 * j + 3) / 3] * dRGB[(4 * j + 3) % 3];
        out4<ACC>(dsh, j, o);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 16; k++) {  // compile-time k keeps b/gx/gy/gz in registers
        if (k < M) {
          if (k >= 1) {
            const float sk = __ldg(sh + 3 * k) * dRGB[0] + __ldg(sh + 3 * k + 1) * dRGB[1] + __ldg(sh + 3 * k + 2) * dRGB[2];
            ddx += gx[k] * sk;
            ddy += gy[k] * sk;
            ddz += gz[k] * sk;
          }
#pragma unroll
          for (int ch = 0; ch < 3; ch++) {
            if (ACC) dsh[3 * k + ch] += b[k] * dRGB[ch]; else dsh[3 * k + ch] = b[k] * dRGB[ch];
          }
        }
      }
      if (!ACC)
        for (int k = 16; k < M; k++)
          for (int ch = 0; ch < 3; ch++) dsh[3 * k + ch] = 0.f;
    }
    // dnormvdv (auxiliary.h:107-117)
    const float sum2 = ox * ox + oy * oy + oz * oz;
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    dmean[0] += ((+sum2 - ox * ox) * ddx - oy * ox * ddy - oz * ox * ddz) * invsum32;
    dmean[1] += (-ox * oy * ddx + (sum2 - oy * oy) * ddy - oz * oy * ddz) * invsum32;
    dmean[2] += (-ox * oz * ddx - oy * oz * ddy + (sum2 - oz * oz) * ddz) * invsum32;
  } else if (dL_dsh != nullptr && !ACC) {
    for (int k = 0; k < 3 * M; k++) dL_dsh[3 * M * i + k] = 0.f;
  }
  out3<ACC>(dL_dmean3D, i, dmean[0], dmean[1], dmean[2]);

  // ---------------- K9: cov3D -> scale, rotation (backward.cu:278-341) ----------------
  if (scales != nullptr && cov3D_precomp == nullptr) {
    float dscale[3];
    float4 dq;
    cov3d_backward(q, sc, vp.scale_modifier, dcov, dscale, dq);
    if (dL_dscale) out3<ACC>(dL_dscale, i, dscale[0], dscale[1], dscale[2]);
    if (dL_drot) out4<ACC>(dL_drot, i, dq);
  } else if (!ACC) {
    if (dL_dscale) store3(dL_dscale, i, 0.f, 0.f, 0.f);
    if (dL_drot) reinterpret_cast<float4*>(dL_drot)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ---- fit step: the per-Gaussian backward of ALL views of a step in one pass -------------------
// The per-view outputs of K8/K9 are 62 floats per Gaussian, their inputs 59; what differs from view
// to view is only the 9 blend-stage sums (48-byte acc row incl. flags) and the camera. Running K8/K9
// once per view therefore moves ~780 B/Gaussian/view when accumulating; looping over the views
// inside the thread keeps the 62 sums in registers and moves 52 B/Gaussian/view + 484 B once.
// dL/dscale and dL/dq are linear in dL/dcov3D, so that part runs once on the summed dL/dcov3D.
constexpr int CAM_FLOATS = 40;   // view[16] | proj[16] | campos[3] | tan_fovx | tan_fovy | pad[3]
constexpr int MAX_BATCH_VIEWS = 64;

// RAW: the outputs are gradients w.r.t. GaussianModel's raw parameters (gaussian_model.py:221-258:
// scaling = exp(raw), opacity = sigmoid(raw), rotation = normalize(raw), features = cat(f_dc, f_rest)),
// i.e. the activations' backward is applied in the epilogue and dL/dsh is split into f_dc / f_rest
// rows (dL_dsh -> f_dc [P,1,3], dL_drest -> f_rest [P,15,3]); every row is written (no accumulate).
// A/B switches (tests/gpu_r2_ab.sh), measured at config 2: 3 / 4 / 5 CTAs per SM (168 registers without spills / 128
// / 96): 0.525 / 0.498 / 0.572 ms; pass-1 groups of 2 / 4 / 8 views: 0.559 / 0.498 / 0.496 ms.
#ifndef DGE_GEOM_MIN_CTAS
#define DGE_GEOM_MIN_CTAS 4
#endif
#ifndef DGE_GEOM_GROUP
#define DGE_GEOM_GROUP 4
#endif
template <bool RAW>
__global__ void __launch_bounds__(128, DGE_GEOM_MIN_CTAS) geom_backward_batched_kernel(
    int P, int D, int M, int V, const float* __restrict__ cams, int W, int H, float scale_modifier,
    const float* __restrict__ acc, size_t acc_stride, const uint8_t* __restrict__ flags, size_t flags_stride,
    const float* __restrict__ means3D,
    const float* __restrict__ shs, const float* __restrict__ scales, const float* __restrict__ rotations,
    const float* __restrict__ opacities, const float* __restrict__ rotation_raw,
    float* __restrict__ dL_dmean3D, float* __restrict__ dL_dmean2D, float* __restrict__ dL_dsh,
    float* __restrict__ dL_drest, float* __restrict__ dL_dopacity, float* __restrict__ dL_dscale,
    float* __restrict__ dL_drot, bool accumulate) {
  extern __shared__ float s_cam[];
  for (int k = threadIdx.x; k < V * CAM_FLOATS; k += blockDim.x) s_cam[k] = cams[k];
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (!RAW && idx >= P) return;
  // RAW: every thread reaches the barrier of the staged f_rest store; threads past the end work on the
  // last Gaussian and write nothing
  const bool live = idx < P;
  const size_t i = (size_t)min(idx, P - 1);
  const float mx = __ldg(means3D + 3 * i), my = __ldg(means3D + 3 * i + 1), mz = __ldg(means3D + 3 * i + 2);
  const float4 q = __ldg(reinterpret_cast<const float4*>(rotations) + i);
  const float sc[3] = {__ldg(scales + 3 * i), __ldg(scales + 3 * i + 1), __ldg(scales + 3 * i + 2)};
  float c3[6];
  cov3d_from_scale_rot(sc[0], sc[1], sc[2], scale_modifier, q, c3);
  const float* sh = shs + 3 * (size_t)M * i;

  float dmean[3] = {0.f, 0.f, 0.f}, dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float dm2x = 0.f, dm2y = 0.f, dop = 0.f;
  float dsh[48];
#pragma unroll
  for (int k = 0; k < 48; k++) dsh[k] = 0.f;
  // Pass 1: which views gave this Gaussian anything? A row is all zeros when the Gaussian was
  // culled, occluded or below alpha 1/255 everywhere in that view — the common case (ncu: 2 of 32
  // lanes had work per view when the views were walked in lock-step). Every term below is
  // linear in the nine sums, so such views are skipped; each lane then walks ITS OWN list of
  // non-empty views, and a warp iterates max-over-lanes(#non-empty) times instead of V times.
  // Visibility and the SH clamp mask come from the view's flag bytes (one coalesced byte load per view),
  // the sums of the visible views from their 48-byte rows.
  bool any = false;
  unsigned long long todo = 0ull;
  // Four views at a time, all loads of a level issued before any is used: walked one view after the
  // other, the dependent loads per view put 2 V memory round trips on every thread's critical path —
  // that latency, not bandwidth, was the kernel's time.
  constexpr int GRP = DGE_GEOM_GROUP;
  for (int v0 = 0; v0 < V; v0 += GRP) {
    uint32_t fl[GRP];
#pragma unroll
    for (int k = 0; k < GRP; k++) fl[k] = v0 + k < V ? (uint32_t)__ldg(flags + (size_t)(v0 + k) * flags_stride + i) : 0u;
    unsigned vis = 0;
#pragma unroll
    for (int k = 0; k < GRP; k++)
      if (fl[k] & 1u) vis |= 1u << k;
    if (!vis) continue;
    any = true;
    float4 a0[GRP], a1[GRP], a2[GRP];
#pragma unroll
    for (int k = 0; k < GRP; k++)
      if ((vis >> k) & 1u) {
        const float4* row = reinterpret_cast<const float4*>(acc + (size_t)(v0 + k) * acc_stride) + 3 * i;
        a0[k] = __ldg(row);
        a1[k] = __ldg(row + 1);
        a2[k] = __ldg(row + 2);
      }
#pragma unroll
    for (int k = 0; k < GRP; k++)
      if (((vis >> k) & 1u) &&
          (a0[k].x != 0.f || a0[k].y != 0.f || a0[k].z != 0.f || a0[k].w != 0.f || a1[k].x != 0.f || a1[k].y != 0.f ||
           a1[k].z != 0.f || a1[k].w != 0.f || a2[k].x != 0.f))
        todo |= 1ull << (v0 + k);
  }
  while (todo) {
    const int v = __ffsll((long long)todo) - 1;
    todo &= todo - 1ull;
    const float4* row = reinterpret_cast<const float4*>(acc + (size_t)v * acc_stride) + 3 * i;
    const float4 a0 = __ldg(row), a1 = __ldg(row + 1), a2 = __ldg(row + 2);
    const uint32_t fl = __ldg(flags + (size_t)v * flags_stride + i);  // (L1-resident from pass 1)
    const float* cam = s_cam + v * CAM_FLOATS;
    const float tan_fovx = cam[35], tan_fovy = cam[36];
    const float focal_y = H / (2.0f * tan_fovy), focal_x = W / (2.0f * tan_fovx);
    float vm[3], vc[6], db[5];
    view_geom_backward(mx, my, mz, c3, cam, cam + 16, tan_fovx, tan_fovy, focal_x, focal_y, W, H, a0.x, a0.y, a0.z,
                       a0.w, a1.x, db, vm, vc);
#pragma unroll
    for (int k = 0; k < 6; k++) dcov[k] += vc[k];
    dm2x += db[0];
    dm2y += db[1];
    dop += a1.y;
    // SH part (backward.cu:20-139)
    const float ox = mx - cam[32], oy = my - cam[33], oz = mz - cam[34];
    const float len = sqrtf(ox * ox + oy * oy + oz * oz);
    const float x = ox / len, y = oy / len, z = oz / len;
    float dRGB[3] = {a1.z, a1.w, a2.x};
#pragma unroll
    for (int ch = 0; ch < 3; ch++)
      if ((fl >> (1 + ch)) & 1u) dRGB[ch] = 0.f;
    float b[16], gx[16], gy[16], gz[16];
    sh_basis_grad(D, x, y, z, b, gx, gy, gz);
    float ddx = 0.f, ddy = 0.f, ddz = 0.f;
    if (M == 16) {
      float4 w4[12];
#pragma unroll
      for (int j = 0; j < 12; j++) w4[j] = __ldg(reinterpret_cast<const float4*>(sh) + j);
      const float* f = reinterpret_cast<const float*>(w4);
#pragma unroll
      for (int k = 1; k < 16; k++) {
        const float sk = f[3 * k] * dRGB[0] + f[3 * k + 1] * dRGB[1] + f[3 * k + 2] * dRGB[2];
        ddx += gx[k] * sk;
        ddy += gy[k] * sk;
        ddz += gz[k] * sk;
      }
    } else {
#pragma unroll
      for (int k = 1; k < 16; k++)
        if (k < M) {
          const float sk = __ldg(sh + 3 * k) * dRGB[0] + __ldg(sh + 3 * k + 1) * dRGB[1] + __ldg(sh + 3 * k + 2) * dRGB[2];
          ddx += gx[k] * sk;
          ddy += gy[k] * sk;
          ddz += gz[k] * sk;
        }
    }
#pragma unroll
    for (int k = 0; k < 16; k++)
#pragma unroll
      for (int ch = 0; ch < 3; ch++) dsh[3 * k + ch] += b[k] * dRGB[ch];
    const float sum2 = ox * ox + oy * oy + oz * oz;
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    dmean[0] += vm[0] + ((+sum2 - ox * ox) * ddx - oy * ox * ddy - oz * ox * ddz) * invsum32;
    dmean[1] += vm[1] + (-ox * oy * ddx + (sum2 - oy * oy) * ddy - oz * oy * ddz) * invsum32;
    dmean[2] += vm[2] + (-ox * oz * ddx - oy * oz * ddy + (sum2 - oz * oz) * ddz) * invsum32;
  }
  if (!any && accumulate) return;
  float dscale[3] = {0.f, 0.f, 0.f};
  float4 dq = make_float4(0.f, 0.f, 0.f, 0.f);
  if (any) cov3d_backward(q, sc, scale_modifier, dcov, dscale, dq);
  if (RAW) {
    const float o = __ldg(opacities + i);
    const float4 qr = __ldg(reinterpret_cast<const float4*>(rotation_raw) + i);  // normalize
    const float nrm = fmaxf(sqrtf(qr.x * qr.x + qr.y * qr.y + qr.z * qr.z + qr.w * qr.w), 1e-12f);
    const float qd = q.x * dq.x + q.y * dq.y + q.z * dq.z + q.w * dq.w;
    const float inv = 1.0f / nrm;
    if (live) {
      dL_dopacity[i] = dop * o * (1.0f - o);                                 // sigmoid
      out4<false>(dL_drot, i, make_float4((dq.x - q.x * qd) * inv, (dq.y - q.y * qd) * inv,
                                          (dq.z - q.z * qd) * inv, (dq.w - q.w * qd) * inv));
    }
    const int first = blockIdx.x * blockDim.x;
    const int rows = min((int)blockDim.x, P - first);
    __syncthreads();  // s_cam is dead: reuse the dynamic shared memory as the staging area
    // The four [P,3] outputs: three scalar stores each with a 12-byte stride touch every sector three
    // times; staged (stride 3 words: conflict-free), each leaves as one contiguous block of the CTA.
    {
      float* s3 = s_cam;
      const float vals[4][3] = {{dmean[0], dmean[1], dmean[2]},
                                {dm2x, dm2y, 0.f},
                                {dscale[0] * sc[0], dscale[1] * sc[1], dscale[2] * sc[2]},  // exp
                                {dsh[0], dsh[1], dsh[2]}};                                 // f_dc
#pragma unroll
      for (int b = 0; b < 4; b++)
#pragma unroll
        for (int k = 0; k < 3; k++) s3[b * 3 * 128 + threadIdx.x * 3 + k] = vals[b][k];
      __syncthreads();
      float* const outs[4] = {dL_dmean3D, dL_dmean2D, dL_dscale, dL_dsh};
#pragma unroll
      for (int b = 0; b < 4; b++) {
        float* dst = outs[b] + 3 * (size_t)first;  // 128 rows x 12 B: 16-byte aligned when the tensor is
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
          const int n4 = rows * 3 / 4;
          for (int j = threadIdx.x; j < n4; j += blockDim.x)
            reinterpret_cast<float4*>(dst)[j] = reinterpret_cast<const float4*>(s3 + b * 3 * 128)[j];
          for (int j = n4 * 4 + threadIdx.x; j < rows * 3; j += blockDim.x) dst[j] = s3[b * 3 * 128 + j];
        } else {
          for (int j = threadIdx.x; j < rows * 3; j += blockDim.x) dst[j] = s3[b * 3 * 128 + j];
        }
      }
    }
    // f_rest: 45 floats per Gaussian. Written straight from the registers, each of the 45 store
    // instructions of a warp touches 32 different sectors (stride 180 B); staged through shared memory
    // (stride 45 words: conflict-free) the CTA's 128 rows leave as one contiguous, 16-byte-aligned block.
    __syncthreads();  // the [P,3] blocks have left the staging area
    float* s_rest = s_cam;
#pragma unroll
    for (int k = 0; k < 45; k++) s_rest[threadIdx.x * 45 + k] = dsh[3 + k];
    __syncthreads();
    float4* dst = reinterpret_cast<float4*>(dL_drest + 45 * (size_t)first);
    const float4* src = reinterpret_cast<const float4*>(s_rest);
    const int n4 = rows * 45 / 4;
    for (int j = threadIdx.x; j < n4; j += blockDim.x) dst[j] = src[j];
    for (int j = n4 * 4 + threadIdx.x; j < rows * 45; j += blockDim.x) dL_drest[45 * (size_t)first + j] = s_rest[j];
    return;
  }
  if (accumulate) {
    out3<true>(dL_dmean3D, i, dmean[0], dmean[1], dmean[2]);
    out3<true>(dL_dmean2D, i, dm2x, dm2y, 0.f);
    dL_dopacity[i] += dop;
    out3<true>(dL_dscale, i, dscale[0], dscale[1], dscale[2]);
    out4<true>(dL_drot, i, dq);
  } else {
    out3<false>(dL_dmean3D, i, dmean[0], dmean[1], dmean[2]);
    out3<false>(dL_dmean2D, i, dm2x, dm2y, 0.f);
    dL_dopacity[i] = dop;
    out3<false>(dL_dscale, i, dscale[0], dscale[1], dscale[2]);
    out4<false>(dL_drot, i, dq);
  }
  float* o = dL_dsh + 3 * (size_t)M * i;
  if (M == 16) {
#pragma unroll
    for (int j = 0; j < 12; j++) {
      const float4 v4 = make_float4(dsh[4 * j], dsh[4 * j + 1], dsh[4 * j + 2], dsh[4 * j + 3]);
      if (accumulate) out4<true>(o, j, v4); else out4<false>(o, j, v4);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 48; k++)
      if (k < 3 * M) {
        if (accumulate) o[k] += dsh[k]; else o[k] = dsh[k];
      }
    if (!accumulate)
      for (int k = 48; k < 3 * M; k++) o[k] = 0.f;
  }
}

// GaussianModel's activations (gaussian_model.py:221-258) for the whole model in one pass:
// shs = cat(f_dc, f_rest), opacity = sigmoid, scaling = exp, rotation = normalize (eps 1e-12).
__global__ void __launch_bounds__(256) activate_kernel(int P, const float* __restrict__ opacity_raw,
                                                       const float* __restrict__ scaling_raw,
                                                       const float* __restrict__ rotation_raw,
                                                       float* __restrict__ opacities, float* __restrict__ scales,
                                                       float* __restrict__ rotations) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const size_t i = (size_t)idx;
  opacities[i] = 1.0f / (1.0f + expf(-__ldg(opacity_raw + i)));
#pragma unroll
  for (int k = 0; k < 3; k++) scales[3 * i + k] = expf(__ldg(scaling_raw + 3 * i + k));
  const float4 q = __ldg(reinterpret_cast<const float4*>(rotation_raw) + i);
  const float inv = 1.0f / fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
  reinterpret_cast<float4*>(rotations)[i] = make_float4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
}

// shs = cat(f_dc, f_rest): one thread per float4 of the output (a Gaussian's 48 floats are 12 of them), so that a
// warp reads one contiguous run of f_rest and writes 512 contiguous bytes (one thread per Gaussian touched 32
// different 180-byte rows per load instruction: 38 % of the copy bandwidth)
__global__ void __launch_bounds__(256) cat_features_kernel(size_t n4, const float* __restrict__ f_dc,
                                                           const float* __restrict__ f_rest,
                                                           float4* __restrict__ shs) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n4) return;
  const size_t i = j / 12;
  const int c = (int)(j - i * 12) * 4;  // first of this thread's four coefficient slots, 0 .. 44
  float4 v;
  if (c == 0) {
    v = make_float4(__ldg(f_dc + 3 * i), __ldg(f_dc + 3 * i + 1), __ldg(f_dc + 3 * i + 2), __ldg(f_rest + 45 * i));
  } else {
    const float* r = f_rest + 45 * i + (c - 3);
    v = make_float4(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3));
  }
  shs[j] = v;
}

cudaError_t launch_activate(int P, const float* f_dc, const float* f_rest, const float* opacity_raw,
                            const float* scaling_raw, const float* rotation_raw, float* shs,
                            float* opacities, float* scales, float* rotations, cudaStream_t stream) {
  // either half may be left out (NULL inputs): the geometry activations of the next step run before its
  // features have been stepped (multi-GPU pipelining, fit.py)
  const int which = (opacity_raw != nullptr ? 1 : 0) | (f_dc != nullptr ? 2 : 0);
  if (which & 2) {
    const size_t n4 = (size_t)P * 12;
    cat_features_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(n4, f_dc, f_rest,
                                                                        reinterpret_cast<float4*>(shs));
    DGE_LAUNCHED(1);
  }
  if (which & 1) {
    activate_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, opacity_raw, scaling_raw, rotation_raw, opacities, scales,
                                                         rotations);
    DGE_LAUNCHED(1);
  }
  return cudaGetLastError();
}

cudaError_t launch_geom_backward_batched(int P, int D, int M, int V, const float* cams, int W, int H,
                                         float scale_modifier, const float* acc, size_t acc_stride,
                                         const uint8_t* flags, size_t flags_stride,
                                         const float* means3D, const float* shs, const float* scales,
                                         const float* rotations, float* dL_dmean3D, float* dL_dmean2D,
                                         float* dL_dsh, float* dL_dopacity, float* dL_dscale,
                                         float* dL_drot, bool accumulate, cudaStream_t stream,
                                         const float* opacities, const float* rotation_raw,
                                         float* dL_drest) {
  if (V < 1 || V > MAX_BATCH_VIEWS || flags == nullptr) return cudaErrorInvalidValue;
  size_t smem = V * CAM_FLOATS * sizeof(float);
  if (rotation_raw != nullptr) {
    smem = smem > 128 * 45 * sizeof(float) ? smem : 128 * 45 * sizeof(float);  // the staged f_rest block
    if (M != 16 || (reinterpret_cast<uintptr_t>(dL_drest) & 15)) return cudaErrorInvalidValue;  // 16-byte aligned f_rest rows
    geom_backward_batched_kernel<true><<<(P + 127) / 128, 128, smem, stream>>>(
        P, D, M, V, cams, W, H, scale_modifier, acc, acc_stride, flags, flags_stride, means3D, shs, scales, rotations,
        opacities, rotation_raw, dL_dmean3D, dL_dmean2D, dL_dsh, dL_drest, dL_dopacity, dL_dscale, dL_drot, false);
  } else {
    geom_backward_batched_kernel<false><<<(P + 127) / 128, 128, smem, stream>>>(
        P, D, M, V, cams, W, H, scale_modifier, acc, acc_stride, flags, flags_stride, means3D, shs, scales, rotations,
        nullptr, nullptr, dL_dmean3D, dL_dmean2D, dL_dsh, nullptr, dL_dopacity, dL_dscale, dL_drot, accumulate);
  }
  DGE_LAUNCHED(1);
  return cudaGetLastError();
}

cudaError_t launch_geom_backward(const ViewParams& vp, const float* means3D, const float* scales,
                                 const float* rotations, const float* shs, const float* cov3D_precomp,
                                 const int* radii, const GeomState& g, const float* acc,
                                 float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                                 float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                                 float* dL_dscale, float* dL_drot, bool accumulate,
                                 cudaStream_t stream) {
  if (accumulate)
    geom_backward_kernel<true><<<(vp.P + 255) / 256, 256, 0, stream>>>(
        vp, means3D, scales, rotations, shs, cov3D_precomp, radii, g.clamped, acc, dL_dmean2D,
        dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot);
  else
    geom_backward_kernel<false><<<(vp.P + 255) / 256, 256, 0, stream>>>(
        vp, means3D, scales, rotations, shs, cov3D_precomp, radii, g.clamped, acc, dL_dmean2D,
        dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot);
  DGE_LAUNCHED(1);
  return cudaGetLastError();
}

}  // namespace dge
