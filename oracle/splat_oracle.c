/*
 * TEST INFRASTRUCTURE — CPU restatement of the DGE rasterizer hot path.
 *
 * Plain C (gcc, -ffp-contract=off), one function per stage of the reference's
 * CudaRasterizer (DGR/ = gaussiansplatting/submodules/diff-gaussian-rasterization/),
 * each citing the reference lines it follows. Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load the resulting liboracle.so; it is the
 * checker, never the thing shipped or measured.
 *
 * Pinning: the reference ships no tests or golden vectors, so this
 * restatement is pinned against outputs of the reference itself — the unmodified
 * reference CUDA code rebuilt for sm_100a (oracle/_ref) and run on a B200 by
 * oracle/make_golden.py; the resulting fixtures live in tests/golden/ and are
 * checked by tests/test_oracle_golden.py.
 *
 * Bit-exactness: the per-Gaussian forward stage (everything that decides tile
 * membership and the sort key) reproduces the reference's arithmetic exactly —
 * explicit fmaf() in the places where the reference's own build (nvcc 12.9,
 * -fmad=true, sm_100a) contracts a multiply-add, read off its SASS; IEEE sqrt and
 * division; FP64 ndc2Pix. The blend stages use libm expf, which differs from
 * CUDA's expf by <= 2 ulp, so image outputs are compared with a tolerance.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TILE 16

typedef struct {
  int P, D, M, W, H;
  float tan_fovx, tan_fovy, scale_modifier;
  const float *view, *proj, *campos, *bg;
} OView;

static inline float dot3_ref(float a, float b, float c, float d, float e, float f) {
  /* a*b + c*d + e*f as the reference's build contracts it */
  return fmaf(e, f, fmaf(a, b, c * d));
}
static inline float xform_row(const float* m, int r, float x, float y, float z) {
  /* DGR/cuda_rasterizer/auxiliary.h:58-77 transformPoint4x3/4x4, one row */
  return dot3_ref(m[r], x, m[4 + r], y, m[8 + r], z) + m[12 + r];
}
static inline int f2i_trunc(float v) { /* CUDA F2I.TRUNC saturates; NaN -> 0 */
  if (!(v == v)) return 0;
  if (v >= 2147483648.0f) return INT32_MAX;
  if (v <= -2147483648.0f) return INT32_MIN;
  return (int)v;
}
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* DGR/cuda_rasterizer/auxiliary.h:46-56 getRect (as compiled: +16 then -1 as two adds) */
static void get_rect(float px, float py, int radius, int gx, int gy, int* r) {
  const float rf = (float)radius;
  r[0] = imin(gx, imax(0, f2i_trunc((px - rf) * 0.0625f)));
  r[1] = imin(gy, imax(0, f2i_trunc((py - rf) * 0.0625f)));
  r[2] = imin(gx, imax(0, f2i_trunc((((px + rf) + 16.0f) - 1.0f) * 0.0625f)));
  r[3] = imin(gy, imax(0, f2i_trunc((((py + rf) + 16.0f) - 1.0f) * 0.0625f)));
}

/* DGR/cuda_rasterizer/forward.cu:118-152 computeCov3D (quaternion NOT normalised) */
static void cov3d(const float* s3, float mod, const float* q4, float* cov) {
  const float r = q4[0], x = q4[1], y = q4[2], z = q4[3];
  const float xz = x * z, rx = r * x, rz = r * z, yy = y * y, zz = z * z;
  const float xz_p_ry = fmaf(r, y, xz), xz_m_ry = fmaf(-r, y, xz);
  const float yz_m_rx = fmaf(y, z, -rx), yz_p_rx = fmaf(y, z, rx);
  const float xy_m_rz = fmaf(x, y, -rz), xy_p_rz = fmaf(x, y, rz);
  const float xx_p_yy = fmaf(x, x, yy), yy_p_zz = yy + zz, xx_p_zz = fmaf(x, x, zz);
  const float R00 = 1.0f - (yy_p_zz + yy_p_zz), R01 = xy_m_rz + xy_m_rz, R02 = xz_p_ry + xz_p_ry;
  const float R10 = xy_p_rz + xy_p_rz, R11 = 1.0f - (xx_p_zz + xx_p_zz), R12 = yz_m_rx + yz_m_rx;
  const float R20 = xz_m_ry + xz_m_ry, R21 = yz_p_rx + yz_p_rx, R22 = 1.0f - (xx_p_yy + xx_p_yy);
  const float s0 = s3[0] * mod, s1 = s3[1] * mod, s2 = s3[2] * mod;
  /* M = S * R as GLM evaluates it, literal-zero products kept (they decide the sign of zeros) */
  const volatile float Zv = 0.0f;
  const float Z = Zv;
  const float M00 = dot3_ref(s0, R00, Z, R01, Z, R02), M01 = dot3_ref(Z, R00, s1, R01, Z, R02),
              M02 = dot3_ref(Z, R00, Z, R01, s2, R02);
  const float M10 = dot3_ref(s0, R10, Z, R11, Z, R12), M11 = dot3_ref(Z, R10, s1, R11, Z, R12),
              M12 = dot3_ref(Z, R10, Z, R11, s2, R12);
  const float M20 = dot3_ref(s0, R20, Z, R21, Z, R22), M21 = dot3_ref(Z, R20, s1, R21, Z, R22),
              M22 = dot3_ref(Z, R20, Z, R21, s2, R22);
  cov[0] = dot3_ref(M00, M00, M01, M01, M02, M02);
  cov[1] = dot3_ref(M10, M00, M11, M01, M12, M02);
  cov[2] = dot3_ref(M20, M00, M21, M01, M22, M02);
  cov[3] = dot3_ref(M10, M10, M11, M11, M12, M12);
  cov[4] = dot3_ref(M20, M10, M21, M11, M22, M12);
  cov[5] = dot3_ref(M20, M20, M21, M21, M22, M22);
}

/* DGR/cuda_rasterizer/forward.cu:74-113 computeCov2D -> (a, b, c) with the 0.3 low-pass */
static void cov2d(float tx, float ty, float tz, const OView* v, float fx, float fy, const float* c,
                  float* out) {
  const float* V = v->view;
  const float limx = v->tan_fovx * 1.3f, limy = v->tan_fovy * 1.3f;
  const float txtz = tx / tz, tytz = ty / tz;
  const float cx = fminf(fmaxf(txtz, -limx), limx), cy = fminf(fmaxf(tytz, -limy), limy);
  const float tz2 = tz * tz;
  const float J00 = fx / tz, J02 = ((tz * -cx) * fx) / tz2;
  const float J11 = fy / tz, J12 = ((tz * -cy) * fy) / tz2;
  /* T = W * J (GLM), J's literal zeros kept */
  const volatile float Zv = 0.0f;
  const float Z = Zv;
  const float T00 = dot3_ref(V[0], J00, V[1], Z, V[2], J02), T01 = dot3_ref(V[4], J00, V[5], Z, V[6], J02),
              T02 = dot3_ref(V[8], J00, V[9], Z, V[10], J02);
  const float T10 = dot3_ref(V[0], Z, V[1], J11, V[2], J12), T11 = dot3_ref(V[4], Z, V[5], J11, V[6], J12),
              T12 = dot3_ref(V[8], Z, V[9], J11, V[10], J12);
  const float A00 = dot3_ref(T00, c[0], T01, c[1], T02, c[2]);
  const float A10 = dot3_ref(T00, c[1], T01, c[3], T02, c[4]);
  const float A20 = dot3_ref(T00, c[2], T01, c[4], T02, c[5]);
  const float A01 = dot3_ref(T10, c[0], T11, c[1], T12, c[2]);
  const float A11 = dot3_ref(T10, c[1], T11, c[3], T12, c[4]);
  const float A21 = dot3_ref(T10, c[2], T11, c[4], T12, c[5]);
  out[0] = dot3_ref(T00, A00, T01, A10, T02, A20) + 0.3f;
  out[1] = dot3_ref(T00, A01, T01, A11, T02, A21);
  out[2] = dot3_ref(T10, A01, T11, A11, T12, A21) + 0.3f;
}

/* DGR/cuda_rasterizer/auxiliary.h:22-39 */
static const float SH_C0 = 0.28209479177387814f, SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

/* DGR/cuda_rasterizer/forward.cu:20-71 computeColorFromSH; res = value before +0.5 */
static void sh_to_rgb(int deg, const float* d3, const float* sh, float* res) {
  const float len = sqrtf(fmaf(d3[2], d3[2], fmaf(d3[0], d3[0], d3[1] * d3[1])));
  const float x = d3[0] / len, y = d3[1] / len, z = d3[2] / len;
#define S(k, c) sh[3 * (k) + (c)]
  for (int c = 0; c < 3; c++) res[c] = S(0, c) * SH_C0;
  if (deg < 1) return;
  const float t1 = y * SH_C1, t2 = z * SH_C1, t3 = x * SH_C1;
  for (int c = 0; c < 3; c++) res[c] = fmaf(-t3, S(3, c), fmaf(t2, S(2, c), fmaf(-t1, S(1, c), res[c])));
  if (deg < 2) return;
  const float xx = x * x, yy = y * y, zz = z * z, xy = y * x, yz = z * y, xz = z * x;
  const float zz2 = zz + zz, xx_m_yy = xx - yy;
  const float k4 = xy * SH_C2[0], k5 = yz * SH_C2[1], k6 = (-yy + (-xx + zz2)) * SH_C2[2];
  const float k7 = xz * SH_C2[3], k8 = xx_m_yy * SH_C2[4];
  for (int c = 0; c < 3; c++) {
    float r = fmaf(k4, S(4, c), res[c]);
    r = fmaf(k5, S(5, c), r);
    r = fmaf(k6, S(6, c), r);
    r = fmaf(k7, S(7, c), r);
    res[c] = fmaf(k8, S(8, c), r);
  }
  if (deg < 3) return;
  const float f4 = -yy + fmaf(zz, 4.0f, -xx);
  const float k9 = (y * SH_C3[0]) * fmaf(xx, 3.0f, -yy);
  const float k10 = (xy * SH_C3[1]) * z;
  const float k11 = (y * SH_C3[2]) * f4;
  const float k12 = (z * SH_C3[3]) * fmaf(yy, -3.0f, fmaf(xx, -3.0f, zz2));
  const float k13 = f4 * (x * SH_C3[4]);
  const float k14 = xx_m_yy * (z * SH_C3[5]);
  const float k15 = (x * SH_C3[6]) * fmaf(yy, -3.0f, xx);
  for (int c = 0; c < 3; c++) {
    float r = fmaf(k9, S(9, c), res[c]);
    r = fmaf(k10, S(10, c), r);
    r = fmaf(k11, S(11, c), r);
    r = fmaf(k12, S(12, c), r);
    r = fmaf(k13, S(13, c), r);
    r = fmaf(k14, S(14, c), r);
    res[c] = fmaf(k15, S(15, c), r);
  }
#undef S
}

/* K1 preprocessCUDA (DGR/cuda_rasterizer/forward.cu:155-256) incl. in_frustum
 * (auxiliary.h:139-164) and ndc2Pix (auxiliary.h:41-44). Outputs of culled Gaussians: radii,
 * tiles_touched, rect = 0, everything else untouched (the reference leaves garbage). */
void oracle_preprocess(const OView* v, const float* means3D, const float* scales, const float* rotations,
                       const float* opacities, const float* shs, const float* cov3D_precomp,
                       const float* colors_precomp, int* radii, float* means2D, float* depths, float* cov3D,
                       float* conic_opacity, float* rgb, uint8_t* clamped, uint32_t* tiles_touched, int* rect) {
  const int gx = (v->W + TILE - 1) / TILE, gy = (v->H + TILE - 1) / TILE;
  const float focal_y = v->H / (2.0f * v->tan_fovy), focal_x = v->W / (2.0f * v->tan_fovx);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < v->P; i++) {
    radii[i] = 0;
    tiles_touched[i] = 0;
    rect[4 * i] = rect[4 * i + 1] = rect[4 * i + 2] = rect[4 * i + 3] = 0;
    const float px = means3D[3 * i], py = means3D[3 * i + 1], pz = means3D[3 * i + 2];
    const float depth = xform_row(v->view, 2, px, py, pz);
    if (!(depth > 0.2f)) continue;
    const float p_w = 1.0f / (xform_row(v->proj, 3, px, py, pz) + 0.0000001f);
    const float projx = xform_row(v->proj, 0, px, py, pz) * p_w;
    const float projy = xform_row(v->proj, 1, px, py, pz) * p_w;
    float c3[6];
    if (cov3D_precomp) {
      memcpy(c3, cov3D_precomp + 6 * (size_t)i, sizeof c3);
    } else {
      cov3d(scales + 3 * (size_t)i, v->scale_modifier, rotations + 4 * (size_t)i, c3);
      memcpy(cov3D + 6 * (size_t)i, c3, sizeof c3);
    }
    const float tx = xform_row(v->view, 0, px, py, pz), ty = xform_row(v->view, 1, px, py, pz);
    float cov[3];
    cov2d(tx, ty, depth, v, focal_x, focal_y, c3, cov);
    const float det = fmaf(cov[0], cov[2], -(cov[1] * cov[1]));
    if (det == 0.0f) continue;
    const float inv = 1.0f / det;
    const float mid = (cov[0] + cov[2]) * 0.5f;
    const float s = sqrtf(fmaxf(fmaf(mid, mid, -det), 0.1f));
    const float lam = fmaxf(mid + s, mid - s);
    const int rad = f2i_trunc(ceilf(sqrtf(lam) * 3.0f));
    const float pix_x = (float)(fma((double)projx + 1.0, (double)v->W, -1.0) * 0.5);
    const float pix_y = (float)(fma((double)projy + 1.0, (double)v->H, -1.0) * 0.5);
    int r[4];
    get_rect(pix_x, pix_y, rad, gx, gy, r);
    const uint32_t cnt = (uint32_t)(r[2] - r[0]) * (uint32_t)(r[3] - r[1]);
    if (cnt == 0) continue;
    if (!colors_precomp) {
      const float d3[3] = {px - v->campos[0], py - v->campos[1], pz - v->campos[2]};
      float res[3];
      sh_to_rgb(v->D, d3, shs + 3 * (size_t)v->M * i, res);
      for (int c = 0; c < 3; c++) {
        const float val = res[c] + 0.5f;
        clamped[3 * (size_t)i + c] = val < 0.0f;
        rgb[3 * (size_t)i + c] = fmaxf(val, 0.0f);
      }
    }
    depths[i] = depth;
    radii[i] = rad;
    means2D[2 * (size_t)i] = pix_x;
    means2D[2 * (size_t)i + 1] = pix_y;
    conic_opacity[4 * (size_t)i] = cov[2] * inv;
    conic_opacity[4 * (size_t)i + 1] = cov[1] * -inv;
    conic_opacity[4 * (size_t)i + 2] = cov[0] * inv;
    conic_opacity[4 * (size_t)i + 3] = opacities[i];
    tiles_touched[i] = cnt;
    memcpy(rect + 4 * (size_t)i, r, sizeof r);
  }
}

/* K12 checkFrustum (DGR/cuda_rasterizer/rasterizer_impl.cu:53-63) */
void oracle_mark_visible(int P, const float* means3D, const float* view, uint8_t* present) {
  for (int i = 0; i < P; i++)
    present[i] = xform_row(view, 2, means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]) > 0.2f;
}

/* K2 InclusiveSum (DGR/cuda_rasterizer/rasterizer_impl.cu:229-232); returns num_rendered */
uint32_t oracle_scan(int P, const uint32_t* tiles_touched, uint32_t* offsets) {
  uint32_t run = 0;
  for (int i = 0; i < P; i++) {
    run += tiles_touched[i];
    offsets[i] = run;
  }
  return run;
}

/* K3 duplicateWithKeys (DGR/cuda_rasterizer/rasterizer_impl.cu:67-100) */
void oracle_duplicate(int P, int W, int H, const float* means2D, const float* depths, const uint32_t* offsets,
                      const int* radii, uint64_t* keys, uint32_t* vals) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  for (int i = 0; i < P; i++) {
    if (radii[i] <= 0) continue;
    uint32_t off = i == 0 ? 0 : offsets[i - 1];
    int r[4];
    get_rect(means2D[2 * (size_t)i], means2D[2 * (size_t)i + 1], radii[i], gx, gy, r);
    uint32_t dbits;
    memcpy(&dbits, depths + i, 4);
    for (int y = r[1]; y < r[3]; y++)
      for (int x = r[0]; x < r[2]; x++) {
        keys[off] = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
        vals[off] = (uint32_t)i;
        off++;
      }
  }
}

/* K4 cub::DeviceRadixSort::SortPairs (DGR/cuda_rasterizer/rasterizer_impl.cu:256-261): a stable
 * sort on key bits [0, end_bit) — restated as a stable LSD byte radix sort. */
void oracle_sort_pairs(uint32_t n, int end_bit, uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp,
                       uint32_t* vals_tmp) {
  uint64_t *ka = keys, *kb = keys_tmp;
  uint32_t *va = vals, *vb = vals_tmp;
  for (int shift = 0; shift < end_bit; shift += 8) {
    const int bits = end_bit - shift < 8 ? end_bit - shift : 8;
    const uint64_t mask = (1ull << bits) - 1;
    size_t count[257] = {0};
    for (uint32_t i = 0; i < n; i++) count[((ka[i] >> shift) & mask) + 1]++;
    for (int d = 0; d < 256; d++) count[d + 1] += count[d];
    for (uint32_t i = 0; i < n; i++) {
      const size_t dst = count[(ka[i] >> shift) & mask]++;
      kb[dst] = ka[i];
      vb[dst] = va[i];
    }
    uint64_t* tk = ka; ka = kb; kb = tk;
    uint32_t* tv = va; va = vb; vb = tv;
  }
  if (ka != keys) {
    memcpy(keys, ka, sizeof(uint64_t) * n);
    memcpy(vals, va, sizeof(uint32_t) * n);
  }
}

/* getHigherMsb (DGR/cuda_rasterizer/rasterizer_impl.cu:36-49) */
uint32_t oracle_higher_msb(uint32_t n) {
  uint32_t msb = sizeof(n) * 4, step = msb;
  while (step > 1) {
    step /= 2;
    if (n >> msb) msb += step; else msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}

/* K5 identifyTileRanges + memset (DGR/cuda_rasterizer/rasterizer_impl.cu:105-125, :263-265) */
void oracle_tile_ranges(uint32_t R, const uint64_t* keys, int num_tiles, uint32_t* ranges) {
  memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)num_tiles);
  for (uint32_t i = 0; i < R; i++) {
    const uint32_t cur = (uint32_t)(keys[i] >> 32);
    if (i == 0) ranges[2 * cur] = 0;
    else {
      const uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
      if (cur != prev) { ranges[2 * prev + 1] = i; ranges[2 * cur] = i; }
    }
    if (i == R - 1) ranges[2 * cur + 1] = R;
  }
}

static inline float power_of(const float* xy, const float* co, float pxf, float pyf, float* dx, float* dy) {
  /* DGR/cuda_rasterizer/forward.cu:338-341 as compiled:
   * fma(fma(dx, cx*dx, (cz*dy)*dy), -0.5, -((cy*dx)*dy)) */
  *dx = xy[0] - pxf;
  *dy = xy[1] - pyf;
  const float a = *dy * (*dy * co[2]);
  return fmaf(fmaf(*dx, *dx * co[0], a), -0.5f, -(*dy * (*dx * co[1])));
}

/* K6 renderCUDA forward (DGR/cuda_rasterizer/forward.cu:261-379) */
void oracle_render_forward(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                           const float* means2D, const float* colors, const float* depths,
                           const float* conic_opacity, const float* bg, float* out_color, float* out_depth,
                           float* final_T, uint32_t* n_contrib) {
  const int gx = (W + TILE - 1) / TILE;
  const size_t HW = (size_t)H * W;
#pragma omp parallel for schedule(dynamic, 64)
  for (int pix = 0; pix < H * W; pix++) {
    const int px = pix % W, py = pix / W;
    const int tile = (py / TILE) * gx + px / TILE;
    const uint32_t beg = ranges[2 * tile], end = ranges[2 * tile + 1];
    float T = 1.0f, C[3] = {0, 0, 0}, D = 0.0f;
    uint32_t last = 0;
    for (uint32_t k = beg; k < end; k++) {
      const uint32_t g = point_list[k];
      const float* co = conic_opacity + 4 * (size_t)g;
      float dx, dy;
      const float power = power_of(means2D + 2 * (size_t)g, co, (float)px, (float)py, &dx, &dy);
      if (power > 0.0f) continue;
      const float alpha = fminf(0.99f, co[3] * expf(power));
      if (alpha < 1.0f / 255.0f) continue;
      const float test_T = T * (1.0f - alpha);
      if (test_T < 0.0001f) break;
      for (int c = 0; c < 3; c++) C[c] = fmaf(T, alpha * colors[3 * (size_t)g + c], C[c]);
      D = fmaf(T, alpha * depths[g], D);
      T = test_T;
      last = k - beg + 1;
    }
    final_T[pix] = T;
    n_contrib[pix] = last;
    for (int c = 0; c < 3; c++) out_color[c * HW + pix] = fmaf(bg[c], T, C[c]);
    out_depth[pix] = D;
  }
}

/* K7 renderCUDA backward (DGR/cuda_rasterizer/backward.cu:399-557). Sums over pixels are taken
 * in double, so they are the order-independent value the reference's float atomics approximate
 * (pixels run on all host threads; the order of the double additions then varies at the 1e-16
 * level, twelve orders of magnitude below the 1e-4 the gradients are compared with). Outputs
 * must be zeroed by the caller (they are accumulated). */
void oracle_render_backward(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                            const float* bg, const float* means2D, const float* conic_opacity,
                            const float* colors, const float* final_T, const uint32_t* n_contrib,
                            const float* dL_dpix, double* dL_dmean2D /*[P,2]*/, double* dL_dconic /*[P,3]*/,
                            double* dL_dopacity /*[P]*/, double* dL_dcolor /*[P,3]*/) {
  const int gx = (W + TILE - 1) / TILE;
  const size_t HW = (size_t)H * W;
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
#pragma omp parallel for schedule(dynamic, 256)
  for (int pix = 0; pix < H * W; pix++) {
    const int px = pix % W, py = pix / W;
    const int tile = (py / TILE) * gx + px / TILE;
    const uint32_t beg = ranges[2 * tile];
    const float T_final = final_T[pix];
    float T = T_final;
    float accum[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0.0f;
    const float g3[3] = {dL_dpix[pix], dL_dpix[HW + pix], dL_dpix[2 * HW + pix]};
    for (uint32_t k = n_contrib[pix]; k-- > 0;) {
      const uint32_t g = point_list[beg + k];
      const float* co = conic_opacity + 4 * (size_t)g;
      float dx, dy;
      const float power = power_of(means2D + 2 * (size_t)g, co, (float)px, (float)py, &dx, &dy);
      if (power > 0.0f) continue;
      const float G = expf(power);
      const float alpha = fminf(0.99f, co[3] * G);
      if (alpha < 1.0f / 255.0f) continue;
      T = T / (1.0f - alpha);
      const float dchannel_dcolor = alpha * T;
      float dL_dalpha = 0.0f;
      for (int c = 0; c < 3; c++) {
        const float col = colors[3 * (size_t)g + c];
        accum[c] = last_alpha * last_color[c] + (1.0f - last_alpha) * accum[c];
        last_color[c] = col;
        dL_dalpha += (col - accum[c]) * g3[c];
#pragma omp atomic
        dL_dcolor[3 * (size_t)g + c] += dchannel_dcolor * g3[c];
      }
      dL_dalpha *= T;
      last_alpha = alpha;
      float bg_dot = 0.0f;
      for (int c = 0; c < 3; c++) bg_dot += bg[c] * g3[c];
      dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot;
      const float dL_dG = co[3] * dL_dalpha;
      const float gdx = G * dx, gdy = G * dy;
      const float dG_ddelx = -gdx * co[0] - gdy * co[1];
      const float dG_ddely = -gdy * co[2] - gdx * co[1];
#pragma omp atomic
      dL_dmean2D[2 * (size_t)g] += dL_dG * dG_ddelx * ddelx_dx;
#pragma omp atomic
      dL_dmean2D[2 * (size_t)g + 1] += dL_dG * dG_ddely * ddely_dy;
#pragma omp atomic
      dL_dconic[3 * (size_t)g] += -0.5f * gdx * dx * dL_dG;
#pragma omp atomic
      dL_dconic[3 * (size_t)g + 1] += -0.5f * gdx * dy * dL_dG;
#pragma omp atomic
      dL_dconic[3 * (size_t)g + 2] += -0.5f * gdy * dy * dL_dG;
#pragma omp atomic
      dL_dopacity[g] += G * dL_dalpha;
    }
  }
}

/* GLM-style column-major 3x3 product, (A*B)[i][j] = sum_k A[k][j]*B[i][k] */
static void m3mul(const float A[3][3], const float B[3][3], float R[3][3]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i][j] = A[0][j] * B[i][0] + A[1][j] * B[i][1] + A[2][j] * B[i][2];
}

/* K8 computeCov2DCUDA + K9 backward preprocessCUDA with its SH and cov3D helpers
 * (DGR/cuda_rasterizer/backward.cu:20-139, :144-274, :278-396). dL_dconic is [P,3] = the
 * reference's float4 slots (x, y, w). All outputs fully written (zeros where the reference
 * leaves torch::zeros). cov3D = the forward's value (precomputed or from oracle_preprocess). */
void oracle_geom_backward(const OView* v, const float* means3D, const int* radii, const float* shs,
                          const uint8_t* clamped, const float* scales, const float* rotations,
                          const float* cov3D, int cov3D_is_precomp, const float* dL_dmean2D, const float* dL_dconic,
                          const float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                          float* dL_dscale, float* dL_drot) {
  const float hy = v->H / (2.0f * v->tan_fovy), hx = v->W / (2.0f * v->tan_fovx);
  const float* V = v->view;
  const float* Pm = v->proj;
  const int M = v->M;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < v->P; i++) {
    float* dmean = dL_dmean3D + 3 * (size_t)i;
    float* dcov = dL_dcov3D + 6 * (size_t)i;
    dmean[0] = dmean[1] = dmean[2] = 0.f;
    for (int k = 0; k < 6; k++) dcov[k] = 0.f;
    if (dL_dsh) for (int k = 0; k < 3 * M; k++) dL_dsh[3 * (size_t)M * i + k] = 0.f;
    for (int k = 0; k < 3; k++) dL_dscale[3 * (size_t)i + k] = 0.f;
    for (int k = 0; k < 4; k++) dL_drot[4 * (size_t)i + k] = 0.f;
    if (!(radii[i] > 0)) continue;
    const float mx = means3D[3 * (size_t)i], my = means3D[3 * (size_t)i + 1], mz = means3D[3 * (size_t)i + 2];
    const float* c3 = cov3D + 6 * (size_t)i;
    const float gX = dL_dconic[3 * (size_t)i], gY = dL_dconic[3 * (size_t)i + 1], gZ = dL_dconic[3 * (size_t)i + 2];
    /* ---- backward.cu:144-274 */
    float tx = V[0] * mx + V[4] * my + V[8] * mz + V[12];
    float ty = V[1] * mx + V[5] * my + V[9] * mz + V[13];
    const float tz = V[2] * mx + V[6] * my + V[10] * mz + V[14];
    const float limx = 1.3f * v->tan_fovx, limy = 1.3f * v->tan_fovy;
    const float txtz = tx / tz, tytz = ty / tz;
    tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
    ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
    const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    const float J[3][3] = {{hx / tz, 0.f, -(hx * tx) / (tz * tz)}, {0.f, hy / tz, -(hy * ty) / (tz * tz)}, {0, 0, 0}};
    const float Wm[3][3] = {{V[0], V[4], V[8]}, {V[1], V[5], V[9]}, {V[2], V[6], V[10]}};
    const float Vrk[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
    float T[3][3], Tt[3][3], tmp[3][3], cov2[3][3];
    m3mul(Wm, J, T);
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) Tt[a][b] = T[b][a];
    m3mul(Tt, Vrk, tmp); /* Vrk symmetric: transpose(Vrk) == Vrk */
    m3mul(tmp, T, cov2);
    const float a = cov2[0][0] + 0.3f, b = cov2[0][1], c = cov2[1][1] + 0.3f;
    const float denom = a * c - b * b;
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    if (denom2inv != 0) {
      dL_da = denom2inv * (-c * c * gX + 2 * b * c * gY + (denom - a * c) * gZ);
      dL_dc = denom2inv * (-a * a * gZ + 2 * a * b * gY + (denom - a * c) * gX);
      dL_db = denom2inv * 2 * (b * c * gX - (denom + 2 * b * b) * gY + a * b * gZ);
      dcov[0] = T[0][0] * T[0][0] * dL_da + T[0][0] * T[1][0] * dL_db + T[1][0] * T[1][0] * dL_dc;
      dcov[3] = T[0][1] * T[0][1] * dL_da + T[0][1] * T[1][1] * dL_db + T[1][1] * T[1][1] * dL_dc;
      dcov[5] = T[0][2] * T[0][2] * dL_da + T[0][2] * T[1][2] * dL_db + T[1][2] * T[1][2] * dL_dc;
      dcov[1] = 2 * T[0][0] * T[0][1] * dL_da + (T[0][0] * T[1][1] + T[0][1] * T[1][0]) * dL_db + 2 * T[1][0] * T[1][1] * dL_dc;
      dcov[2] = 2 * T[0][0] * T[0][2] * dL_da + (T[0][0] * T[1][2] + T[0][2] * T[1][0]) * dL_db + 2 * T[1][0] * T[1][2] * dL_dc;
      dcov[4] = 2 * T[0][2] * T[0][1] * dL_da + (T[0][1] * T[1][2] + T[0][2] * T[1][1]) * dL_db + 2 * T[1][1] * T[1][2] * dL_dc;
    }
    float dT[2][3];
    for (int k = 0; k < 3; k++) {
      const float r0 = T[0][0] * Vrk[k][0] + T[0][1] * Vrk[k][1] + T[0][2] * Vrk[k][2];
      const float r1 = T[1][0] * Vrk[k][0] + T[1][1] * Vrk[k][1] + T[1][2] * Vrk[k][2];
      dT[0][k] = 2 * r0 * dL_da + r1 * dL_db;
      dT[1][k] = 2 * r1 * dL_dc + r0 * dL_db;
    }
    const float dJ00 = Wm[0][0] * dT[0][0] + Wm[0][1] * dT[0][1] + Wm[0][2] * dT[0][2];
    const float dJ02 = Wm[2][0] * dT[0][0] + Wm[2][1] * dT[0][1] + Wm[2][2] * dT[0][2];
    const float dJ11 = Wm[1][0] * dT[1][0] + Wm[1][1] * dT[1][1] + Wm[1][2] * dT[1][2];
    const float dJ12 = Wm[2][0] * dT[1][0] + Wm[2][1] * dT[1][1] + Wm[2][2] * dT[1][2];
    const float itz = 1.f / tz, itz2 = itz * itz, itz3 = itz2 * itz;
    const float dtx = x_grad_mul * -hx * itz2 * dJ02;
    const float dty = y_grad_mul * -hy * itz2 * dJ12;
    const float dtz = -hx * itz2 * dJ00 - hy * itz2 * dJ11 + (2 * hx * tx) * itz3 * dJ02 + (2 * hy * ty) * itz3 * dJ12;
    dmean[0] = V[0] * dtx + V[1] * dty + V[2] * dtz;
    dmean[1] = V[4] * dtx + V[5] * dty + V[6] * dtz;
    dmean[2] = V[8] * dtx + V[9] * dty + V[10] * dtz;
    /* ---- backward.cu:366-387 */
    {
      const float gx2 = dL_dmean2D[2 * (size_t)i], gy2 = dL_dmean2D[2 * (size_t)i + 1];
      const float m_w = 1.0f / ((Pm[3] * mx + Pm[7] * my + Pm[11] * mz + Pm[15]) + 0.0000001f);
      const float mul1 = (Pm[0] * mx + Pm[4] * my + Pm[8] * mz + Pm[12]) * m_w * m_w;
      const float mul2 = (Pm[1] * mx + Pm[5] * my + Pm[9] * mz + Pm[13]) * m_w * m_w;
      dmean[0] += (Pm[0] * m_w - Pm[3] * mul1) * gx2 + (Pm[1] * m_w - Pm[3] * mul2) * gy2;
      dmean[1] += (Pm[4] * m_w - Pm[7] * mul1) * gx2 + (Pm[5] * m_w - Pm[7] * mul2) * gy2;
      dmean[2] += (Pm[8] * m_w - Pm[11] * mul1) * gx2 + (Pm[9] * m_w - Pm[11] * mul2) * gy2;
    }
    /* ---- backward.cu:20-139 */
    if (shs) {
      const float* sh = shs + 3 * (size_t)M * i;
      float* dsh = dL_dsh + 3 * (size_t)M * i;
      const float o[3] = {mx - v->campos[0], my - v->campos[1], mz - v->campos[2]};
      const float len = sqrtf(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
      const float x = o[0] / len, y = o[1] / len, z = o[2] / len;
      float dRGB[3], ddx[3] = {0, 0, 0}, ddy[3] = {0, 0, 0}, ddz[3] = {0, 0, 0};
      for (int ch = 0; ch < 3; ch++) dRGB[ch] = clamped[3 * (size_t)i + ch] ? 0.f : dL_dcolor[3 * (size_t)i + ch];
#define S(k, c) sh[3 * (k) + (c)]
#define PUT(k, coef) for (int ch = 0; ch < 3; ch++) dsh[3 * (k) + ch] = (coef) * dRGB[ch]
      PUT(0, SH_C0);
      if (v->D > 0) {
        PUT(1, -SH_C1 * y); PUT(2, SH_C1 * z); PUT(3, -SH_C1 * x);
        for (int ch = 0; ch < 3; ch++) { ddx[ch] = -SH_C1 * S(3, ch); ddy[ch] = -SH_C1 * S(1, ch); ddz[ch] = SH_C1 * S(2, ch); }
        if (v->D > 1) {
          const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
          PUT(4, SH_C2[0] * xy); PUT(5, SH_C2[1] * yz); PUT(6, SH_C2[2] * (2.f * zz - xx - yy));
          PUT(7, SH_C2[3] * xz); PUT(8, SH_C2[4] * (xx - yy));
          for (int ch = 0; ch < 3; ch++) {
            ddx[ch] += SH_C2[0] * y * S(4, ch) + SH_C2[2] * 2.f * -x * S(6, ch) + SH_C2[3] * z * S(7, ch) + SH_C2[4] * 2.f * x * S(8, ch);
            ddy[ch] += SH_C2[0] * x * S(4, ch) + SH_C2[1] * z * S(5, ch) + SH_C2[2] * 2.f * -y * S(6, ch) + SH_C2[4] * 2.f * -y * S(8, ch);
            ddz[ch] += SH_C2[1] * y * S(5, ch) + SH_C2[2] * 2.f * 2.f * z * S(6, ch) + SH_C2[3] * x * S(7, ch);
          }
          if (v->D > 2) {
            PUT(9, SH_C3[0] * y * (3.f * xx - yy)); PUT(10, SH_C3[1] * xy * z);
            PUT(11, SH_C3[2] * y * (4.f * zz - xx - yy)); PUT(12, SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
            PUT(13, SH_C3[4] * x * (4.f * zz - xx - yy)); PUT(14, SH_C3[5] * z * (xx - yy));
            PUT(15, SH_C3[6] * x * (xx - 3.f * yy));
            for (int ch = 0; ch < 3; ch++) {
              ddx[ch] += SH_C3[0] * S(9, ch) * 3.f * 2.f * xy + SH_C3[1] * S(10, ch) * yz + SH_C3[2] * S(11, ch) * -2.f * xy +
                         SH_C3[3] * S(12, ch) * -3.f * 2.f * xz + SH_C3[4] * S(13, ch) * (-3.f * xx + 4.f * zz - yy) +
                         SH_C3[5] * S(14, ch) * 2.f * xz + SH_C3[6] * S(15, ch) * 3.f * (xx - yy);
              ddy[ch] += SH_C3[0] * S(9, ch) * 3.f * (xx - yy) + SH_C3[1] * S(10, ch) * xz +
                         SH_C3[2] * S(11, ch) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * S(12, ch) * -3.f * 2.f * yz +
                         SH_C3[4] * S(13, ch) * -2.f * xy + SH_C3[5] * S(14, ch) * -2.f * yz + SH_C3[6] * S(15, ch) * -3.f * 2.f * xy;
              ddz[ch] += SH_C3[1] * S(10, ch) * xy + SH_C3[2] * S(11, ch) * 4.f * 2.f * yz +
                         SH_C3[3] * S(12, ch) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * S(13, ch) * 4.f * 2.f * xz +
                         SH_C3[5] * S(14, ch) * (xx - yy);
            }
          }
        }
      }
#undef S
#undef PUT
      const float dd[3] = {ddx[0] * dRGB[0] + ddx[1] * dRGB[1] + ddx[2] * dRGB[2],
                           ddy[0] * dRGB[0] + ddy[1] * dRGB[1] + ddy[2] * dRGB[2],
                           ddz[0] * dRGB[0] + ddz[1] * dRGB[1] + ddz[2] * dRGB[2]};
      /* auxiliary.h:107-117 dnormvdv */
      const float sum2 = o[0] * o[0] + o[1] * o[1] + o[2] * o[2];
      const float inv32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
      dmean[0] += ((+sum2 - o[0] * o[0]) * dd[0] - o[1] * o[0] * dd[1] - o[2] * o[0] * dd[2]) * inv32;
      dmean[1] += (-o[0] * o[1] * dd[0] + (sum2 - o[1] * o[1]) * dd[1] - o[2] * o[1] * dd[2]) * inv32;
      dmean[2] += (-o[0] * o[2] * dd[0] - o[1] * o[2] * dd[1] + (sum2 - o[2] * o[2]) * dd[2]) * inv32;
    }
    /* ---- backward.cu:278-341 */
    if (scales && !cov3D_is_precomp) {
      const float* q = rotations + 4 * (size_t)i;
      const float r = q[0], x = q[1], y = q[2], z = q[3];
      const float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                             {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                             {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
      const float s[3] = {v->scale_modifier * scales[3 * (size_t)i], v->scale_modifier * scales[3 * (size_t)i + 1],
                          v->scale_modifier * scales[3 * (size_t)i + 2]};
      float Mm[3][3], dM[3][3], dMt[3][3];
      for (int ci = 0; ci < 3; ci++) for (int rj = 0; rj < 3; rj++) Mm[ci][rj] = s[rj] * R[ci][rj];
      const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                              {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                              {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
      m3mul(Mm, dS, dM);
      for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) dMt[k][l] = 2.0f * dM[l][k];
      for (int k = 0; k < 3; k++)
        dL_dscale[3 * (size_t)i + k] = R[0][k] * dMt[k][0] + R[1][k] * dMt[k][1] + R[2][k] * dMt[k][2];
      for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) dMt[k][l] *= s[k];
      float* dq = dL_drot + 4 * (size_t)i;
      dq[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
      dq[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
      dq[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
      dq[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
    }
  }
}

/* K14 renderCUDA_apply_weights<CH> (DGR/cuda_rasterizer/apply_weights.cu:239-356): weights and
 * cnt are accumulated in place; cnt gets +1 per channel per contributing pair (:331-334). */
void oracle_apply_weights_render(int W, int H, int CH, const uint32_t* ranges, const uint32_t* point_list,
                                 const float* means2D, const float* conic_opacity, const float* image_weights,
                                 double* weights /*[P,CH]*/, int64_t* cnt /*[P]*/) {
  const int gx = (W + TILE - 1) / TILE;
  const size_t HW = (size_t)H * W;
  for (int pix = 0; pix < H * W; pix++) {
    const int px = pix % W, py = pix / W;
    const int tile = (py / TILE) * gx + px / TILE;
    const uint32_t beg = ranges[2 * tile], end = ranges[2 * tile + 1];
    float T = 1.0f;
    for (uint32_t k = beg; k < end; k++) {
      const uint32_t g = point_list[k];
      const float* co = conic_opacity + 4 * (size_t)g;
      float dx, dy;
      const float power = power_of(means2D + 2 * (size_t)g, co, (float)px, (float)py, &dx, &dy);
      if (power > 0.0f) continue;
      const float alpha = fminf(0.99f, co[3] * expf(power));
      if (alpha < 1.0f / 255.0f) continue;
      const float test_T = T * (1.0f - alpha);
      if (test_T < 0.0001f) break;
      for (int c = 0; c < CH; c++) {
        weights[(size_t)g * CH + c] += image_weights[c * HW + pix];
        cnt[g] += 1;
      }
      T = test_T;
    }
  }
}
