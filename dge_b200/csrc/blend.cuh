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

struct BlendSmem {
  float4 a[BL_BATCH];    // x, y, conic.x, conic.y
  float4 b[BL_BATCH];    // conic.z, power threshold, opacity, gid (bits)
  float4 c[BL_BATCH];    // r, g, b, depth
  float4 box[BL_BATCH];  // x - hx, x + hx, y - hy, y + hy
  uint8_t list[BL_WARPS][BL_BATCH];
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
    s.a[k] = make_float4(m.x, m.y, co.x, co.y);
    s.b[k] = make_float4(co.z, power_threshold(co.w), co.w, __uint_as_float(gid));
    s.box[k] = make_float4(m.x - m.z, m.x + m.z, m.y - m.w, m.y + m.w);
    if (WITH_COLOR) s.c[k] = rgb_depth[gid];
  }
}

// Stable compaction of the records of this batch that can touch the warp's half-tile.
// keep(k) is an extra per-record predicate (e.g. "position below the warp's last contributor").
template <typename KEEP>
__device__ __forceinline__ int compact_batch(BlendSmem& s, int warp, int lane, int count, float X0,
                                             float X1, float Y0, float Y1, KEEP keep) {
  int n = 0;
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int c = 0; c < BL_BATCH / 32; c++) {
    const int k = c * 32 + lane;
    bool hit = false;
    if (k < count) {
      const float4 bx = s.box[k];
      hit = bx.y >= X0 && bx.x <= X1 && bx.w >= Y0 && bx.z <= Y1 && keep(k);
    }
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, hit);
    if (hit) s.list[warp][n + __popc(m & lt)] = (uint8_t)k;
    n += __popc(m);
  }
  __syncwarp();
  return n;
}

// `power` of the thread's four pixels (p = 2*row + col of the quadrant grid), bit-exact with the reference
// (DGR/cuda_rasterizer/forward.cu:338-341 as compiled): fma(fma(dx, cx*dx, (cz*dy)*dy), -0.5, -((cy*dx)*dy))
struct Quad {
  float dx0, dx1, dy0, dy1;
  float power[4];
};
__device__ __forceinline__ Quad quad_power(const float4 a, const float cz, float fx0, float fx1,
                                           float fy0, float fy1) {
  Quad q;
  q.dx0 = BADD(a.x, -fx0);
  q.dx1 = BADD(a.x, -fx1);
  q.dy0 = BADD(a.y, -fy0);
  q.dy1 = BADD(a.y, -fy1);
  const float bx0 = BMUL(q.dx0, a.z), bx1 = BMUL(q.dx1, a.z);  // conic.x * dx
  const float cx0 = BMUL(q.dx0, a.w), cx1 = BMUL(q.dx1, a.w);  // conic.y * dx
  const float ay0 = BMUL(q.dy0, BMUL(q.dy0, cz)), ay1 = BMUL(q.dy1, BMUL(q.dy1, cz));
  q.power[0] = BFMA(BFMA(q.dx0, bx0, ay0), -0.5f, -BMUL(q.dy0, cx0));
  q.power[1] = BFMA(BFMA(q.dx1, bx1, ay0), -0.5f, -BMUL(q.dy0, cx1));
  q.power[2] = BFMA(BFMA(q.dx0, bx0, ay1), -0.5f, -BMUL(q.dy1, cx0));
  q.power[3] = BFMA(BFMA(q.dx1, bx1, ay1), -0.5f, -BMUL(q.dy1, cx1));
  return q;
}

}  // namespace dge
