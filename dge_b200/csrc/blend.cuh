// Shared machinery of the three per-tile blend kernels (forward, backward, apply_weights).
//
// Layout on the SM: one CTA per 16x16 tile, 64 threads = 2 warps, each thread owns four pixels
// (below), and warp w covers the 16x8 pixel half-tile [y0 + 8w, y0 + 8w + 7]. Instances are staged in
// batches of 128 records; before a warp walks a batch it compacts, IN LIST ORDER (ballot +
// prefix popcount), the indices of the records whose conservative alpha >= 1/255 box (written
// by preprocess next to the pixel centre) overlaps its half-tile, and then only visits those.
// About half of a tile's instances miss a given half-tile, and for those the walk costs one
// compare instead of ~35 instructions. Skipped records would have failed the exact
// per-pixel tests anyway, so results are unchanged bit for bit.
#pragma once
#include "common.cuh"

namespace dge {

// Thread -> pixels. A thread owns FOUR pixels of its warp's 16x8 half-tile, one in each 8x4
// quadrant: (lx, ly), (lx+8, ly), (lx, ly+4), (lx+8, ly+4). The column/row sub-products of
// `power` are shared exactly as for a 2x2 quad (two dx, two dy values), but pixel slot p of
// all 32 lanes together covers ONE compact quadrant, so the expensive per-slot code (expf,
// blending, gradients) runs only for the quadrants a Gaussian's footprint actually reaches
// (2.2 of 4 on average) instead of for nearly every slot as with interleaved 2x2 quads.
constexpr int PX_STEP = 8;
constexpr int PY_STEP = 4;
constexpr int BL_THREADS = 64;
constexpr int BL_WARPS = 2;
constexpr int BL_BATCH = 128;

#define BMUL(a, b) __fmul_rn((a), (b))
#define BADD(a, b) __fadd_rn((a), (b))
#define BFMA(a, b, c) __fmaf_rn((a), (b), (c))

// Fit-step batch: blockIdx.z = view; per-view geometry / image arrays at uniform byte strides, the
// view's instance list at point_list + seg_off[view], its acc rows acc_stride floats apart.
// seg_off == nullptr <=> a single view (the per-view API).
struct BlendBatch {
  size_t geom_stride, img_stride;
  const uint32_t* seg_off;
  size_t acc_stride;
};

struct BlendSmem {
  float4 a[BL_BATCH];    // x, y, conic.x, conic.y
  float4 b[BL_BATCH];    // conic.z, power threshold, opacity, gid (bits)
  float4 c[BL_BATCH];    // r, g, b, depth
  float4 box[BL_BATCH];  // x - hx, x + hx, y - hy, y + hy
  float4 d[BL_BATCH];    // -cy/cz, -cy/cx, limit on the quadratic form, 1 if the exact cull applies
  uint16_t list[BL_WARPS][BL_BATCH];  // record index | quadrant mask << 8
};

// Lower bound on `power` below which opacity*exp(power) < 1/255 for certain.
// 0.01 of slack in the exponent is ~1% in alpha; expf and __logf err by < 1e-6.
__device__ __forceinline__ float power_threshold(float opacity) {
  return opacity > 0.0f ? -(__logf(255.0f * opacity) + 0.01f) : __int_as_float(0x7f800000);
}

// Cooperative gather of `count` records. src(k) gives the position in point_list of record k.
template <bool WITH_COLOR, typename SRC>
__device__ __forceinline__ void stage_batch(BlendSmem& s, int tid, int count, SRC src,
                                            const uint32_t* __restrict__ point_list,
                                            const float4* __restrict__ means2D,
                                            const float4* __restrict__ conic_opacity,
                                            const float4* __restrict__ rgb_depth) {
  for (int k = tid; k < count; k += BL_THREADS) {
    const uint32_t gid = point_list[src(k)];
    const float4 m = means2D[gid];
    const float4 co = conic_opacity[gid];
    const float thr = power_threshold(co.w);
    s.a[k] = make_float4(m.x, m.y, co.x, co.y);
    s.b[k] = make_float4(co.z, thr, co.w, __uint_as_float(gid));
    s.box[k] = make_float4(m.x - m.z, m.x + m.z, m.y - m.w, m.y + m.w);
    // Exact-cull constants (quad_mask): q(u,v) = 0.5 (cx u^2 + cz v^2) + cy u v must stay <= tau =
    // -thr for a pixel to reach alpha >= 1/255. The limit carries a slack proportional to the
    // conditioning kappa = cx cz / det of the form: the fp32 `power` of a pixel differs from the
    // real-number value by < 1e-6 * kappa * q, so a rectangle is only dropped when its minimum of q
    // exceeds tau by 20x that; ill-conditioned needles (kappa > 1e3) are left to the box test.
    const float det = co.x * co.z - co.y * co.y;
    const float ac = co.x * co.z;
    const bool exact = co.x > 0.0f && co.z > 0.0f && det > 1e-3f * ac && thr < 0.0f;
    const float kappa = exact ? __fdividef(ac, det) : 1.0f;
    s.d[k] = make_float4(exact ? __fdividef(-co.y, co.z) : 0.0f, exact ? __fdividef(-co.y, co.x) : 0.0f,
                         -thr * (1.0f + 2e-5f * kappa) + 1e-4f, exact ? 1.0f : 0.0f);
    if (WITH_COLOR) s.c[k] = rgb_depth[gid];
  }
}

// Which of the warp's four 8x4 pixel quadrants (origin X0, Y0 of the 16x8 half-tile, pixel-centre
// coordinates) can hold a pixel with alpha >= 1/255: bit p set <=> quadrant p = 2*row + col must be
// visited. Two tests per quadrant, both only ever DROP pixels that fail the reference's own test:
//  1. the conservative axis-aligned box of {alpha >= 1/255} written by preprocess;
//  2. the minimum of the quadratic form over the quadrant's rectangle. q is convex with its
//     minimum at the Gaussian's centre, so over a rectangle that does not contain the centre the
//     minimum lies on an edge facing it: on u = clamp(0) with v = clamp(-cy u / cz), or on
//     v = clamp(0) with u = clamp(-cy v / cx).
__device__ __forceinline__ uint32_t quad_mask(const float4 a, const float cz, const float4 bx,
                                              const float4 d, float X0, float Y0) {
  uint32_t mask = 0;
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const float xa = X0 + (float)(PX_STEP * (p & 1)), xb = xa + (float)(PX_STEP - 1);
    const float ya = Y0 + (float)(PY_STEP * (p >> 1)), yb = ya + (float)(PY_STEP - 1);
    bool hit = bx.y >= xa && bx.x <= xb && bx.w >= ya && bx.z <= yb;
    if (hit && d.w != 0.0f) {
      const float u0 = xa - a.x, u1 = xb - a.x, v0 = ya - a.y, v1 = yb - a.y;
      const float uc = fminf(fmaxf(0.0f, u0), u1), vc = fminf(fmaxf(0.0f, v0), v1);
      const float vs = fminf(fmaxf(d.x * uc, v0), v1), us = fminf(fmaxf(d.y * vc, u0), u1);
      const float qu = 0.5f * (a.z * uc * uc + cz * vs * vs) + a.w * uc * vs;  // edge u = uc
      const float qv = 0.5f * (a.z * us * us + cz * vc * vc) + a.w * us * vc;  // edge v = vc
      const float inf = __int_as_float(0x7f800000);
      float m = (uc == 0.0f && vc == 0.0f) ? 0.0f : inf;
      if (uc != 0.0f) m = fminf(m, qu);
      if (vc != 0.0f) m = fminf(m, qv);
      hit = !(m > d.z);  // NaN keeps the quadrant
    }
    mask |= hit ? (1u << p) : 0u;
  }
  return mask;
}

// Stable compaction of the records of this batch that can touch the warp's half-tile, each with
// the mask of quadrants to visit. allowed(k) is an extra per-record quadrant mask (e.g. "position
// below the quadrant's last contributor", "quadrant not saturated yet").
template <typename ALLOWED>
__device__ __forceinline__ int compact_batch(BlendSmem& s, int warp, int lane, int count, float X0,
                                             float Y0, ALLOWED allowed) {
  int n = 0;
  const uint32_t lt = (1u << lane) - 1u;
  const float X1 = X0 + 15.0f, Y1 = Y0 + 7.0f;
#pragma unroll
  for (int c = 0; c < BL_BATCH / 32; c++) {
    const int k = c * 32 + lane;
    uint32_t qm = 0;
    if (k < count) {
      const float4 bx = s.box[k];
      if (bx.y >= X0 && bx.x <= X1 && bx.w >= Y0 && bx.z <= Y1) {
        qm = allowed(k);
        if (qm) qm &= quad_mask(s.a[k], s.b[k].x, bx, s.d[k], X0, Y0);
      }
    }
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, qm != 0);
    if (qm) s.list[warp][n + __popc(m & lt)] = (uint16_t)(k | (qm << 8));
    n += __popc(m);
  }
  __syncwarp();
  return n;
}

// `power` of one pixel, bit-exact with the reference (DGR/cuda_rasterizer/forward.cu:338-341 as
// compiled): fma(fma(dx, cx*dx, (cz*dy)*dy), -0.5, -((cy*dx)*dy))
__device__ __forceinline__ float pixel_power(const float4 a, const float cz, float dx, float dy) {
  return BFMA(BFMA(dx, BMUL(dx, a.z), BMUL(dy, BMUL(dy, cz))), -0.5f, -BMUL(dy, BMUL(dx, a.w)));
}

}  // namespace dge
