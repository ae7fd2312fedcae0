// K6: per-tile front-to-back alpha blend (colour + depth).
//
// Replaces renderCUDA<3> (DGR/cuda_rasterizer/forward.cu:261-379). Results are
// bit-identical to the reference: same per-pixel operation sequence (power as
// fma(fma(dx, cx*dx, (cz*dy)*dy), -0.5, -((cy*dx)*dy)), accurate expf, C =
// fma(T, alpha*c, C), thresholds 1/255 and 1e-4) in the same list order.
//
// What is different is how the work is laid out on the SM:
//  * one CTA per 16x16 tile, 64 threads, each owning a 2x2 pixel quad, so every
//    staged Gaussian is read from shared memory once per four pixels and the
//    column/row sub-products of `power` are shared inside the quad;
//  * a per-Gaussian power threshold (staged next to the conic) rejects pairs whose
//    alpha is certainly < 1/255 before the expf; it is conservative by a margin far
//    above expf's error, so no decision of the reference is ever changed;
//  * Gaussians are staged in batches of 128 records; each warp visits only the records whose
//    conservative alpha >= 1/255 box overlaps its 16x8 half-tile (blend.cuh).
// Staged batch of the FORWARD walk: 64 records (measured at config 2, forward / backward blend per 20-view launch:
// 64: 1.530 / 2.295 ms, 128: 1.576 / 2.243 ms, 192: 1.727 / 2.346 ms — the forward stops early, so shorter batches
// stage fewer records it never walks; the backward starts from the last contributor and keeps 128)
#ifndef DGE_BL_BATCH
#define DGE_BL_BATCH 64
#endif
#include "blend.cuh"

namespace dge {

// EXTRA: additionally blends one per-Gaussian scalar (`extra[P]`) exactly like a colour channel and
// writes it, background added per channel, as a second image — DGE's "semantic" render of the edit
// mask (threestudio/systems/DGE.py:198-204: render(..., override_color=mask repeated 3x)) for the price
// of one FFMA per blended pair instead of a second preprocess + sort + blend of the same view.
// measured at config 2, 20 views per launch: 1.67 / 1.60 / 1.55 ms at 12 / 14 / 16 CTAs per SM (with batches of 64
// records: 1.53 at 16, 1.61 at 18 = 56 registers, 1.99 at 21 = 40 registers)
#ifndef DGE_FWD_MIN_CTAS
#define DGE_FWD_MIN_CTAS 16
#endif
// the variant with the fused semantic channel (tests/gpu_time_render_views.py, 20 views of config 2 forward only):
// 3.69 / 3.59 / 3.56 ms at 12 / 14 / 16 CTAs per SM against 3.42 ms without the channel
#ifndef DGE_FWD_EXTRA_MIN_CTAS
#define DGE_FWD_EXTRA_MIN_CTAS 16
#endif
template <bool EXTRA>
__global__ void __launch_bounds__(BL_THREADS, EXTRA ? DGE_FWD_EXTRA_MIN_CTAS : DGE_FWD_MIN_CTAS) render_forward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
    const float4* __restrict__ rec, const float* __restrict__ background,
    float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
    float* __restrict__ out_depth, BlendBatch bb, const float* __restrict__ extra,
    float* __restrict__ out_extra) {
  __shared__ BlendSmem s;
  if (bb.seg_off) {  // view blockIdx.z of a fit-step batch (blend.cuh)
    const size_t gs = blockIdx.z * bb.geom_stride, is = blockIdx.z * bb.img_stride;
    ranges = shift_ptr(ranges, is);
    final_T = shift_ptr(final_T, is);
    n_contrib = shift_ptr(n_contrib, is);
    rec = shift_ptr(rec, gs);
    point_list += bb.seg_off[blockIdx.z];
    out_color += (size_t)blockIdx.z * 3 * H * W;
    out_depth += (size_t)blockIdx.z * H * W;
    if (EXTRA) out_extra += (size_t)blockIdx.z * 3 * H * W;
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // pixel p of this thread = (px0 + PX_STEP*(p&1), py0 + PY_STEP*(p>>1)): one pixel in each 8x4
  // quadrant of the warp's 16x8 half-tile (see blend.cuh)
  const int px0 = blockIdx.x * DGE_TILE + (lane & 7), py0 = blockIdx.y * DGE_TILE + 8 * warp + (lane >> 3);
  const float fx0 = (float)px0, fx1 = (float)(px0 + PX_STEP), fy0 = (float)py0, fy1 = (float)(py0 + PY_STEP);
  // the warp's half-tile, in pixel-centre coordinates
  const float X0 = (float)(blockIdx.x * DGE_TILE);
  const float Y0 = (float)(blockIdx.y * DGE_TILE + 8 * warp);
  // pixel p = 2*row + col inside the quad
  bool inside[4];
#pragma unroll
  for (int p = 0; p < 4; p++) inside[p] = px0 + PX_STEP * (p & 1) < W && py0 + PY_STEP * (p >> 1) < H;

  const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];
  float T[4], C[4][3], Dp[4], E[4];
  uint32_t last[4];
  // pmax[p]: 0 while pixel p is live, -inf once it has terminated (or lies outside the image). The
  // reference's `power > 0` skip is evaluated as `power > pmax[p]`, which also skips every pair of a
  // terminated pixel at no extra instruction (a flag byte cost a test, a move and a byte insert per pair).
  // A NaN power passes both forms; for a terminated pixel it then yields alpha = fminf(0.99, NaN) = 0.99 and
  // T * 0.01 < 1e-4 again (T * (1 - alpha) was already below 1e-4 for some alpha <= 0.99): still no update.
  const float NEG_INF = __int_as_float(0xff800000);
  float pmax[4];
#pragma unroll
  for (int p = 0; p < 4; p++) {
    T[p] = 1.0f;
    C[p][0] = C[p][1] = C[p][2] = 0.0f;
    Dp[p] = 0.0f;
    E[p] = 0.0f;
    last[p] = 0;
    pmax[p] = inside[p] ? 0.0f : NEG_INF;
  }

  stage_init(s, tid);
  uint32_t parity = 0;
  for (uint32_t base = range.x; base < range.y; base += BL_BATCH) {
    const bool all_done = pmax[0] < 0.0f && pmax[1] < 0.0f && pmax[2] < 0.0f && pmax[3] < 0.0f;
    if (__syncthreads_and(all_done)) break;  // also: everyone has finished walking the previous batch
    const int count = min((uint32_t)BL_BATCH, range.y - base);
    stage_batch(s, tid, count, [&](int k) { return base + k; }, point_list, rec, parity, EXTRA ? extra : nullptr, bb.P);
    parity ^= 1u;
    if (__all_sync(0xFFFFFFFFu, all_done)) continue;  // this half-tile is saturated
    // quadrants in which every pixel has terminated need no further visits
    uint32_t live = 0;
#pragma unroll
    for (int p = 0; p < 4; p++) live |= __all_sync(0xFFFFFFFFu, pmax[p] < 0.0f) ? 0u : (1u << p);
    const int n = compact_batch(s, warp, lane, count, X0, Y0, [&](int) { return live; });
    for (int i = 0; i < n; i++) {
      const uint32_t e = s.list[warp][i];  // warp-uniform
      const int j = e & 0xFF;
      const float4 a = s.rec[j][0];   // x, y, conic.x, conic.y
      const float4 b = s.rec[j][1];   // conic.z, power threshold, opacity, -
      const float4 cd = s.rec[j][2];  // r, g, b, depth
#pragma unroll
      for (int p = 0; p < 4; p++) {
        if (!(e & (0x100u << p))) continue;  // warp-uniform branch
        const float power = pixel_power(a, b.x, BADD(a.x, (p & 1) ? -fx1 : -fx0), BADD(a.y, (p >> 1) ? -fy1 : -fy0));
        if (power > pmax[p] || power < b.y) continue;
        const float alpha = fminf(0.99f, BMUL(b.z, expf(power)));
        if (alpha < 1.0f / 255.0f) continue;
        const float test_T = BMUL(T[p], BADD(1.0f, -alpha));
        const bool keep = !(test_T < 0.0001f);
        pmax[p] = keep ? pmax[p] : NEG_INF;  // terminated: this Gaussian does not contribute (forward.cu:352-357)
        if (!keep) continue;
        C[p][0] = BFMA(T[p], BMUL(alpha, cd.x), C[p][0]);
        C[p][1] = BFMA(T[p], BMUL(alpha, cd.y), C[p][1]);
        C[p][2] = BFMA(T[p], BMUL(alpha, cd.z), C[p][2]);
        Dp[p] = BFMA(T[p], BMUL(alpha, cd.w), Dp[p]);
        if (EXTRA) E[p] = BFMA(T[p], BMUL(alpha, s.rec[j][REC_F4].y), E[p]);
        T[p] = test_T;
        last[p] = base - range.x + j + 1;
      }
    }
  }

  const float bg0 = __ldg(background), bg1 = __ldg(background + 1), bg2 = __ldg(background + 2);
  const size_t HW = (size_t)H * W;
#pragma unroll
  for (int p = 0; p < 4; p++) {
    if (!inside[p]) continue;
    const size_t pix = (size_t)(py0 + PY_STEP * (p >> 1)) * W + (px0 + PX_STEP * (p & 1));
    final_T[pix] = T[p];
    n_contrib[pix] = last[p];
    out_color[pix] = BFMA(bg0, T[p], C[p][0]);
    out_color[HW + pix] = BFMA(bg1, T[p], C[p][1]);
    out_color[2 * HW + pix] = BFMA(bg2, T[p], C[p][2]);
    out_depth[pix] = Dp[p];
    if (EXTRA) {
      out_extra[pix] = BFMA(bg0, T[p], E[p]);
      out_extra[HW + pix] = BFMA(bg1, T[p], E[p]);
      out_extra[2 * HW + pix] = BFMA(bg2, T[p], E[p]);
    }
  }
}

cudaError_t launch_render_forward(const ViewParams& vp, const GeomState& g, const BinState& b,
                                  ImgState& img, const float* background, float* out_color,
                                  float* out_depth, cudaStream_t stream) {
  dim3 grid(vp.grid_x, vp.grid_y);
  render_forward_kernel<false><<<grid, BL_THREADS, 0, stream>>>(
      img.ranges, b.point_list, vp.W, vp.H, g.rec, background,
      img.final_T, img.n_contrib, out_color, out_depth, BlendBatch{0, 0, nullptr, 0, (uint32_t)vp.P}, nullptr, nullptr);
  DGE_LAUNCHED(1);
  return cudaGetLastError();
}

cudaError_t launch_render_forward_batched(const ViewParams& vp, const ViewBatch& vb, const GeomState& g0,
                                          const BinState& b, ImgState& img0, const float* background,
                                          float* out_color, float* out_depth, cudaStream_t stream,
                                          const float* extra, float* out_extra) {
  dim3 grid(vp.grid_x, vp.grid_y, vb.V);
  const BlendBatch bb{vb.geom_stride, vb.img_stride, vb.seg_off, 0, (uint32_t)vp.P};
  if (extra != nullptr && out_extra != nullptr)
    render_forward_kernel<true><<<grid, BL_THREADS, 0, stream>>>(img0.ranges, b.point_list, vp.W, vp.H, g0.rec,
                                                                background, img0.final_T, img0.n_contrib, out_color,
                                                                out_depth, bb, extra, out_extra);
  else
    render_forward_kernel<false><<<grid, BL_THREADS, 0, stream>>>(img0.ranges, b.point_list, vp.W, vp.H, g0.rec,
                                                                 background, img0.final_T, img0.n_contrib, out_color,
                                                                 out_depth, bb, nullptr, nullptr);
  DGE_LAUNCHED(1);
  return cudaGetLastError();
}

}  // namespace dge
