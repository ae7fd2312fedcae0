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
