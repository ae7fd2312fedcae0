// Bit-exact restatements of the reference's per-Gaussian arithmetic, shared by the
// forward preprocess (which must reproduce the reference's tile lists exactly) and the
// per-Gaussian backward (which recomputes cov3D instead of storing it).
#pragma once
#include "common.cuh"

namespace dge {

#define MUL(a, b) __fmul_rn((a), (b))
#define ADD(a, b) __fadd_rn((a), (b))
#define FMA(a, b, c) __fmaf_rn((a), (b), (c))

// a*b + c*d + e*f (+k): the reference contracts left-associated sums of products as
// t = c*d; t = fma(a,b,t); t = fma(e,f,t); t = t + k   
__device__ __forceinline__ float dot3_ref(float a, float b, float c, float d, float e, float f) {
  return FMA(e, f, FMA(a, b, MUL(c, d)));
}

// transformPoint4x3 / 4x4 rows (DGR/cuda_rasterizer/auxiliary.h:58-77)
__device__ __forceinline__ float xform_row(const float* m, int r, float x, float y, float z) {
  return ADD(dot3_ref(m[r], x, m[4 + r], y, m[8 + r], z), m[12 + r]);
}

// DGR/cuda_rasterizer/forward.cu:118-152 computeCov3D, operation order from SASS.
__device__ __forceinline__ void cov3d_from_scale_rot(float sx, float sy, float sz, float mod,
                                                     float4 q, float* cov) {
  const float r = q.x, x = q.y, y = q.z, z = q.w;
  const float xz = MUL(x, z), rx = MUL(r, x), rz = MUL(r, z);
  const float yy = MUL(y, y), zz = MUL(z, z);
  const float xz_p_ry = FMA(r, y, xz);
  const float xz_m_ry = FMA(-r, y, xz);
  const float yz_m_rx = FMA(y, z, -rx);
  const float yz_p_rx = FMA(y, z, rx);
  const float xy_m_rz = FMA(x, y, -rz);
  const float xy_p_rz = FMA(x, y, rz);
  const float xx_p_yy = FMA(x, x, yy);
  const float yy_p_zz = ADD(yy, zz);
  const float xx_p_zz = FMA(x, x, zz);
  // GLM columns of R (forward.cu:134-138)
  const float R00 = ADD(1.0f, -ADD(yy_p_zz, yy_p_zz)), R01 = ADD(xy_m_rz, xy_m_rz),
              R02 = ADD(xz_p_ry, xz_p_ry);
  const float R10 = ADD(xy_p_rz, xy_p_rz), R11 = ADD(1.0f, -ADD(xx_p_zz, xx_p_zz)),
              R12 = ADD(yz_m_rx, yz_m_rx);
  const float R20 = ADD(xz_m_ry, xz_m_ry), R21 = ADD(yz_p_rx, yz_p_rx),
              R22 = ADD(1.0f, -ADD(xx_p_yy, xx_p_yy));
  const float s0 = MUL(sx, mod), s1 = MUL(sy, mod), s2 = MUL(sz, mod);
  // M = S * R as GLM evaluates it: M[i][j] = S[0][j]*R[i][0] + S[1][j]*R[i][1] + S[2][j]*R[i][2] with
  // S diagonal. The literal-zero products are kept (the reference's SASS keeps them as FMUL/FFMA
  // with RZ): they never change a value but they decide the SIGN of a zero result.
  const float Z = 0.0f;
  const float M00 = dot3_ref(s0, R00, Z, R01, Z, R02), M01 = dot3_ref(Z, R00, s1, R01, Z, R02),
              M02 = dot3_ref(Z, R00, Z, R01, s2, R02);
  const float M10 = dot3_ref(s0, R10, Z, R11, Z, R12), M11 = dot3_ref(Z, R10, s1, R11, Z, R12),
              M12 = dot3_ref(Z, R10, Z, R11, s2, R12);
  const float M20 = dot3_ref(s0, R20, Z, R21, Z, R22), M21 = dot3_ref(Z, R20, s1, R21, Z, R22),
              M22 = dot3_ref(Z, R20, Z, R21, s2, R22);
  // Sigma[i][j] = M[j][0]*M[i][0] + M[j][1]*M[i][1] + M[j][2]*M[i][2]
  cov[0] = dot3_ref(M00, M00, M01, M01, M02, M02);
  cov[1] = dot3_ref(M10, M00, M11, M01, M12, M02);
  cov[2] = dot3_ref(M20, M00, M21, M01, M22, M02);
  cov[3] = dot3_ref(M10, M10, M11, M11, M12, M12);
  cov[4] = dot3_ref(M20, M10, M21, M11, M22, M12);
  cov[5] = dot3_ref(M20, M20, M21, M21, M22, M22);
}

// DGR/cuda_rasterizer/forward.cu:74-113 computeCov2D. t = view-space mean. Also hands out the intermediate
// T = W * J (upper 2x3: T[0..2] = row 0, T[3..5] = row 1) and A = T * Sigma (same layout), which the
// per-Gaussian backward needs again (backward.cu:160-273 recomputes them).
__device__ __forceinline__ float3 cov2d_parts(float tx, float ty, float tz, float tan_fovx, float tan_fovy,
                                              float focal_x, float focal_y, const float* V, const float* c,
                                              float* T, float* A) {
  const float limx = MUL(tan_fovx, 1.3f), limy = MUL(tan_fovy, 1.3f);
  const float txtz = __fdiv_rn(tx, tz), tytz = __fdiv_rn(ty, tz);
  const float cx = fminf(fmaxf(txtz, -limx), limx);
  const float cy = fminf(fmaxf(tytz, -limy), limy);
  const float tz2 = MUL(tz, tz);
  const float J00 = __fdiv_rn(focal_x, tz);
  const float J02 = __fdiv_rn(MUL(MUL(tz, -cx), focal_x), tz2);
  const float J11 = __fdiv_rn(focal_y, tz);
  const float J12 = __fdiv_rn(MUL(MUL(tz, -cy), focal_y), tz2);
  // T = W * J (GLM): T[i][j] = W[0][j]*J[i][0] + W[1][j]*J[i][1] + W[2][j]*J[i][2], J's literal zeros kept
  // (they only decide the sign of zero results, as in the reference's SASS)
  const float Z = 0.0f;
  const float T00 = dot3_ref(V[0], J00, V[1], Z, V[2], J02);
  const float T01 = dot3_ref(V[4], J00, V[5], Z, V[6], J02);
  const float T02 = dot3_ref(V[8], J00, V[9], Z, V[10], J02);
  const float T10 = dot3_ref(V[0], Z, V[1], J11, V[2], J12);
  const float T11 = dot3_ref(V[4], Z, V[5], J11, V[6], J12);
  const float T12 = dot3_ref(V[8], Z, V[9], J11, V[10], J12);
  // A = T^T * Vrk^T ; A[k][j] = fma(Tj2, Vrk[2][k], fma(Tj0, Vrk[0][k], Tj1*Vrk[1][k]))
  const float A00 = dot3_ref(T00, c[0], T01, c[1], T02, c[2]);
  const float A10 = dot3_ref(T00, c[1], T01, c[3], T02, c[4]);
  const float A20 = dot3_ref(T00, c[2], T01, c[4], T02, c[5]);
  const float A01 = dot3_ref(T10, c[0], T11, c[1], T12, c[2]);
  const float A11 = dot3_ref(T10, c[1], T11, c[3], T12, c[4]);
  const float A21 = dot3_ref(T10, c[2], T11, c[4], T12, c[5]);
  // cov[i][j] = fma(A[2][j], T[i][2], fma(A[0][j], T[i][0], A[1][j]*T[i][1]))
  float3 cov;
  cov.x = ADD(dot3_ref(T00, A00, T01, A10, T02, A20), 0.3f);
  cov.y = dot3_ref(T00, A01, T01, A11, T02, A21);
  cov.z = ADD(dot3_ref(T10, A01, T11, A11, T12, A21), 0.3f);
  if (T != nullptr) {
    T[0] = T00; T[1] = T01; T[2] = T02; T[3] = T10; T[4] = T11; T[5] = T12;
    A[0] = A00; A[1] = A10; A[2] = A20; A[3] = A01; A[4] = A11; A[5] = A21;
  }
  return cov;
}
__device__ __forceinline__ float3 cov2d_ref(float tx, float ty, float tz, float tan_fovx, float tan_fovy,
                                            float focal_x, float focal_y, const float* V, const float* c) {
  return cov2d_parts(tx, ty, tz, tan_fovx, tan_fovy, focal_x, focal_y, V, c, nullptr, nullptr);
}

// SH constants, DGR/cuda_rasterizer/auxiliary.h:22-39
#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f
#define SH_C2_0 1.0925484305920792f
#define SH_C2_1 -1.0925484305920792f
#define SH_C2_2 0.31539156525252005f
#define SH_C2_3 -1.0925484305920792f
#define SH_C2_4 0.5462742152960396f
#define SH_C3_0 -0.5900435899266435f
#define SH_C3_1 2.890611442640554f
#define SH_C3_2 -0.4570457994644658f
#define SH_C3_3 0.3731763325901154f
#define SH_C3_4 -0.4570457994644658f
#define SH_C3_5 1.445305721320277f
#define SH_C3_6 -0.5900435899266435f


}  // namespace dge
