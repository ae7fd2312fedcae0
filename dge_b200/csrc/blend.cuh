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
#ifndef DGE_BL_BATCH
#define DGE_BL_BATCH 128
#endif
constexpr int BL_BATCH = DGE_BL_BATCH;  // A/B switch (tests/gpu_r2_ab.sh)

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
  uint32_t P;  // number of Gaussians (checked builds: bound of the staged ids)
};

// A staged record occupies 80 bytes of shared memory: the 64-byte record + one float4 holding the
// Gaussian id. The 80-byte stride (20 words) is what keeps the lane-per-record reads of the compaction
// conflict-free: with a 64-byte stride the eight lanes of a quarter-warp hit two bank groups (4-way
// conflicts made both blend kernels 7-11 % slower).
constexpr int SREC_F4 = REC_F4 + 1;
struct __align__(128) BlendSmem {
  float4 rec[BL_BATCH][SREC_F4];      // TMA destination (first 64 bytes of each slot) | id
  uint16_t list[BL_WARPS][BL_BATCH];  // record index | quadrant mask << 8
  uint64_t bar;                       // mbarrier the bulk copies of a batch complete on
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Once per CTA, before the first stage_batch: every thread arrives once per batch.
__device__ __forceinline__ void stage_init(BlendSmem& s, int tid) {
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&s.bar)), "r"(BL_THREADS));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
}

// Stages `count` records of the tile's list into shared memory with ONE TMA bulk copy
// (cp.async.bulk.shared.global, 64 bytes, SASS UBLKCP) per instance: the list is a gather over the
// per-Gaussian record array, and a bulk copy per record gathers without any register staging or
// per-instance arithmetic (thresholds, boxes and cull constants are part of the record; the probe
// profiles/probes/tma_gather_probe.cu stages 21 M records in 0.31 ms this way vs 0.49 ms with LDG.128 +
// STS). src(k) gives the position in point_list of record k. All threads of the CTA call this (each
// arrives on the mbarrier once per batch); on return the records and ids of the batch are visible to
// all of them. `parity` is the mbarrier phase, flipped by the caller after every batch; the caller
// guarantees that nobody still reads the previous batch. (Double buffering — the next batch in
// flight during the walk — was measured too: no gain, the other 11 CTAs of the SM already cover it.)
template <typename SRC>
__device__ __forceinline__ void stage_batch(BlendSmem& s, int tid, int count, SRC src,
                                            const uint32_t* __restrict__ point_list,
                                            const float4* __restrict__ rec, uint32_t parity,
                                            const float* __restrict__ extra = nullptr, uint32_t num_gaussians = 0) {
  (void)num_gaussians;
  constexpr int PER = BL_BATCH / BL_THREADS;
  uint32_t gid[PER];
  int mine = 0;
#pragma unroll
  for (int i = 0; i < PER; i++) {
    const int k = tid + i * BL_THREADS;
    if (k < count) {
      gid[i] = point_list[src(k)];
      DGE_CHECK(num_gaussians == 0 || gid[i] < num_gaussians);
      s.rec[k][REC_F4].x = __uint_as_float(gid[i]);
      if (extra != nullptr) s.rec[k][REC_F4].y = extra[gid[i]];  // a fourth blended channel (forward only)
      mine++;
    }
  }
  const uint32_t bar = smem_addr(&s.bar);
  // arrive (release: the id stores above) + announce this thread's bytes BEFORE issuing its copies
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(mine * 64) : "memory");
#pragma unroll
  for (int i = 0; i < PER; i++) {
    const int k = tid + i * BL_THREADS;
    if (k < count)
      asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];" ::"r"(
                       smem_addr(&s.rec[k][0])),
                   "l"(rec + (size_t)gid[i] * REC_F4), "r"(bar)
                   : "memory");
  }
  // try_wait with a suspend-time hint: the hardware parks the thread until the phase completes (or the
  // hint expires) instead of returning at once — a bare try_wait loop spun through 5-7 % of these
  // issue-bound kernels' instruction slots (ncu: 1554 M vs 1474 M warp instructions in the forward)
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(bar), "r"(parity), "r"(0x989680)
                 : "memory");
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
                                              const float nbc, const float nba, const float lim, float X0,
                                              float Y0) {
  uint32_t mask = 0;
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const float xa = X0 + (float)(PX_STEP * (p & 1)), xb = xa + (float)(PX_STEP - 1);
    const float ya = Y0 + (float)(PY_STEP * (p >> 1)), yb = ya + (float)(PY_STEP - 1);
    bool hit = bx.y >= xa && bx.x <= xb && bx.w >= ya && bx.z <= yb;
    if (hit && lim >= 0.0f) {
      const float u0 = xa - a.x, u1 = xb - a.x, v0 = ya - a.y, v1 = yb - a.y;
      const float uc = fminf(fmaxf(0.0f, u0), u1), vc = fminf(fmaxf(0.0f, v0), v1);
      const float vs = fminf(fmaxf(nbc * uc, v0), v1), us = fminf(fmaxf(nba * vc, u0), u1);
      const float qu = 0.5f * (a.z * uc * uc + cz * vs * vs) + a.w * uc * vs;  // edge u = uc
      const float qv = 0.5f * (a.z * us * us + cz * vc * vc) + a.w * us * vc;  // edge v = vc
      const float inf = __int_as_float(0x7f800000);
      float m = (uc == 0.0f && vc == 0.0f) ? 0.0f : inf;
      if (uc != 0.0f) m = fminf(m, qu);
      if (vc != 0.0f) m = fminf(m, qv);
      hit = !(m > lim);  // NaN keeps the quadrant
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
      const float4 a = s.rec[k][0], c = s.rec[k][1], d = s.rec[k][3];  // c: conic.z, threshold, opacity, hx
      // x - hx, x + hx, y - hy, y + hy
      const float4 bx = make_float4(a.x - c.w, a.x + c.w, a.y - d.x, a.y + d.x);
      if (bx.y >= X0 && bx.x <= X1 && bx.w >= Y0 && bx.z <= Y1) {
        qm = allowed(k);
        if (qm) qm &= quad_mask(a, c.x, bx, d.y, d.z, d.w, X0, Y0);
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
