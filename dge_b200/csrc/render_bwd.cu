// K7: backward of the per-tile alpha blend.
//
// Replaces renderCUDA<3> backward (DGR/cuda_rasterizer/backward.cu:399-557). Same
// per-pixel recurrences (T /= 1-alpha walking back to front, accum_rec, the
// background term) and the same skip decisions as the forward (identical `power`
// and alpha arithmetic), but:
//  * the traversal starts at the tile's largest n_contrib instead of the end of the
//    tile list, so occluded instances are never touched;
//  * each thread owns four pixels (one per 8x4 quadrant, blend.cuh) and sums their contributions in registers;
//  * what a pair contributes is reduced to the MOMENTS of q = dL/dG * G over the Gaussian's pixels
//    (sum q d, sum q d d^T with d = centre - pixel): the reference's per-pair products with the conic and
//    with W/2, H/2 (backward.cu:537-551) are linear in them and are applied once per (Gaussian, view) by the
//    per-Gaussian backward (geom_bwd.cu:view_geom_backward) — 9 instead of 18 instructions per pair;
//  * the colour behind a pixel (accum_rec) only ever enters dL/dalpha through its dot product with the
//    pixel's dL/dpixel, so ONE scalar per pixel carries that recurrence instead of three (6 instead of 12
//    instructions per pair, 8 registers less);
//  * the nine per-Gaussian partial sums are reduced across the warp with a
//    transposing butterfly (14 shuffles instead of 45) and leave the SM as ONE
//    red.global instruction per warp and Gaussian (9 active lanes) instead of the
//    reference's 9 atomics per contributing (pixel, Gaussian) pair;
//  * warps in which no lane touches a Gaussian skip the reduction (ballot), and each warp only
//    visits the records whose conservative alpha >= 1/255 box overlaps its half-tile (blend.cuh).
// Sums are accumulated into acc[P][12] (see ACC_* in common.cuh); the per-Gaussian
// kernel in geom_bwd.cu turns them into the reference's output tensors.
#include "blend.cuh"

// 16 CTAs (32 warps) per SM at 64 registers: measured 2.39 / 2.32 / 2.25 ms per 20-view launch at 12 / 14 / 16
// (config 2) — once the pair body shrank, hiding the staging and shared-memory latencies mattered more
// than the 12 bytes of spills the tighter budget costs. (The HAS_BG variant has its own setting below.) Beyond 16 the spills win: 2.53 ms at 18 CTAs
// (56 registers), 3.51 ms at 21 (40 registers).
#ifndef DGE_BWD_MIN_CTAS
#define DGE_BWD_MIN_CTAS 16
#endif
// the HAS_BG variant (non-black background; tests/gpu_time_fit_bg.py, config 2 with bg = 0.4): 2.49 / 2.37 / 2.44 ms
// per 20-view launch at 12 / 14 / 16 CTAs per SM (80 / 72 / 64 registers)
#ifndef DGE_BWD_BG_MIN_CTAS
#define DGE_BWD_BG_MIN_CTAS 14
#endif

namespace dge {

// HAS_BG: a non-black background adds the -T_final/(1-alpha) * (bg . dL/dpixel) term to dL/dalpha
// (backward.cu:526-529); for DGE's black background (DGE.py:87) that term, its division and
// eight registers of per-pixel state disappear at compile time.
template <bool HAS_BG>
__global__ void __launch_bounds__(BL_THREADS, HAS_BG ? DGE_BWD_BG_MIN_CTAS : DGE_BWD_MIN_CTAS) render_backward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
    const float* __restrict__ background, const float4* __restrict__ rec,
    const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib,
    const float* __restrict__ dL_dpixels, float* __restrict__ acc, BlendBatch bb) {
  __shared__ BlendSmem s;
  if (bb.seg_off) {  // view blockIdx.z of a fit-step batch (blend.cuh)
    const size_t gs = blockIdx.z * bb.geom_stride, is = blockIdx.z * bb.img_stride;
    ranges = shift_ptr(ranges, is);
    final_Ts = shift_ptr(final_Ts, is);
    n_contrib = shift_ptr(n_contrib, is);
    rec = shift_ptr(rec, gs);
    point_list += bb.seg_off[blockIdx.z];
    dL_dpixels += (size_t)blockIdx.z * 3 * H * W;
    acc += blockIdx.z * bb.acc_stride;
  }
  __shared__ uint32_t s_max[BL_WARPS];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // pixel p of this thread = (px0 + PX_STEP*(p&1), py0 + PY_STEP*(p>>1)): one pixel in each 8x4
  // quadrant of the warp's 16x8 half-tile (see blend.cuh)
  const int px0 = blockIdx.x * DGE_TILE + (lane & 7), py0 = blockIdx.y * DGE_TILE + 8 * warp + (lane >> 3);
  const float fx0 = (float)px0, fx1 = (float)(px0 + PX_STEP), fy0 = (float)py0, fy1 = (float)(py0 + PY_STEP);
  const float X0 = (float)(blockIdx.x * DGE_TILE);
  const float Y0 = (float)(blockIdx.y * DGE_TILE + 8 * warp);
  const size_t HW = (size_t)H * W;
  const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];

  // behind[p] = (colour accumulated behind the current list position) . dL/dpixel, i.e. the reference's
  // accum_rec AFTER its next update, dotted with this pixel's upstream gradient: alpha*c + (1-alpha)*accum
  // is folded in right after a Gaussian is used instead of right before the next one, and since
  // dL/dalpha only needs sum_ch (c_ch - accum_ch) * dL/dpixel_ch, the dot product is carried instead of
  // the three channels.
  float T[4], T_final[4], behind[4], dpix[4][3], bg_dot[4];
  uint32_t last[4];
  const float bg0 = __ldg(background), bg1 = __ldg(background + 1), bg2 = __ldg(background + 2);
  uint32_t tmax = 0;
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int x = px0 + PX_STEP * (p & 1), y = py0 + PY_STEP * (p >> 1);
    const bool inside = x < W && y < H;
    const size_t pix = (size_t)y * W + x;
    T_final[p] = inside ? final_Ts[pix] : 0.0f;
    T[p] = T_final[p];
    last[p] = inside ? n_contrib[pix] : 0;
    tmax = max(tmax, last[p]);
#pragma unroll
    for (int c = 0; c < 3; c++) dpix[p][c] = inside ? dL_dpixels[c * HW + pix] : 0.0f;
    behind[p] = 0.0f;
    bg_dot[p] = HAS_BG ? bg0 * dpix[p][0] + bg1 * dpix[p][1] + bg2 * dpix[p][2] : 0.0f;
  }
  const uint32_t wmax = __reduce_max_sync(0xFFFFFFFFu, tmax);
  uint32_t qmax[4];  // last contributor of each 8x4 quadrant (warp-uniform)
#pragma unroll
  for (int p = 0; p < 4; p++) qmax[p] = __reduce_max_sync(0xFFFFFFFFu, last[p]);
  if (lane == 0) s_max[warp] = wmax;
  __syncthreads();
  uint32_t bmax = 0;
#pragma unroll
  for (int w = 0; w < BL_WARPS; w++) bmax = max(bmax, s_max[w]);


  stage_init(s, tid);
  uint32_t parity = 0;
  // positions hi-1 ... 0 of the tile list, back to front, in batches
  for (uint32_t hi = bmax; hi > 0; hi -= min(hi, (uint32_t)BL_BATCH)) {
    const int count = min(hi, (uint32_t)BL_BATCH);
    __syncthreads();  // everyone has finished walking the previous batch
    stage_batch(s, tid, count, [&](int k) { return range.x + hi - 1 - k; }, point_list, rec, parity, nullptr, bb.P);
    parity ^= 1u;
    if (hi - count >= wmax) continue;  // nothing in this batch reaches this warp (warp-uniform)
    const int n = compact_batch(s, warp, lane, count, X0, Y0, [&](int k) {
      const uint32_t pos = hi - 1 - k;
      return (pos < qmax[0] ? 1u : 0u) | (pos < qmax[1] ? 2u : 0u) | (pos < qmax[2] ? 4u : 0u) |
             (pos < qmax[3] ? 8u : 0u);
    });
    for (int i = 0; i < n; i++) {
      const uint32_t e = s.list[warp][i];  // warp-uniform
      const int j = e & 0xFF;
      const uint32_t pos = hi - 1 - j;  // 0-based list position
      const float4 a = s.rec[j][0];   // x, y, conic.x, conic.y
      const float4 b = s.rec[j][1];   // conic.z, power threshold, opacity, -
      const float cd0 = s.rec[j][2].x, cd1 = s.rec[j][2].y, cd2 = s.rec[j][2].z;  // r, g, b
      const float opacity = b.z;

      float g[9];
#pragma unroll
      for (int i = 0; i < 9; i++) g[i] = 0.0f;
      bool touched = false;
#pragma unroll
      for (int p = 0; p < 4; p++) {
        if (!(e & (0x100u << p))) continue;  // warp-uniform branch
        const float dx = BADD(a.x, (p & 1) ? -fx1 : -fx0), dy = BADD(a.y, (p >> 1) ? -fy1 : -fy0);
        const float power = pixel_power(a, b.x, dx, dy);
        if (pos >= last[p] || power > 0.0f || power < b.y) continue;
        const float G = expf(power);
        const float oG = BMUL(opacity, G);  // the forward's alpha before the 0.99 clamp, bit for bit
        const float alpha = fminf(0.99f, oG);
        if (alpha < 1.0f / 255.0f) continue;
        touched = true;
        const float one_m = 1.0f - alpha;
        // the reference's recurrence T /= (1 - alpha) (backward.cu:505); the quotient only feeds
        // gradients (tolerance 1e-4), so MUFU.RCP + FMUL replace the IEEE division sequence (1 - alpha is in
        // [0.01, 1]: no range handling needed; ~1 ulp per step stays below 1e-5 over a whole list)
        float rcp;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(one_m));
        T[p] *= rcp;
        const float cd = cd0 * dpix[p][0] + cd1 * dpix[p][1] + cd2 * dpix[p][2];  // colour . dL/dpixel
        float dL_dalpha = (cd - behind[p]) * T[p];
        behind[p] = alpha * cd + one_m * behind[p];
        if (HAS_BG) dL_dalpha -= (T_final[p] * rcp) * bg_dot[p];
        const float w = alpha * T[p];
        g[ACC_R] += w * dpix[p][0];
        g[ACC_G] += w * dpix[p][1];
        g[ACC_B] += w * dpix[p][2];
        g[ACC_OPACITY] += G * dL_dalpha;
        // q = dL/dG * G with dL/dG = opacity * dL/dalpha (the reference back-propagates a clamped alpha as if
        // it were not, backward.cu:531); moments of q over the pixels
        const float q = oG * dL_dalpha;
        const float qx = q * dx, qy = q * dy;
        g[ACC_SX] += qx;
        g[ACC_SY] += qy;
        g[ACC_SXX] += qx * dx;
        g[ACC_SXY] += qx * dy;
        g[ACC_SYY] += qy * dy;
      }
      if (!__any_sync(0xFFFFFFFFu, touched)) continue;

      // ---- transposing butterfly: g[0..7] -> lane L holds the warp total of g[L>>2]
      const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
      float w4[4], w2[2], z;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float send = h16 ? g[i] : g[i + 4];
        const float keep = h16 ? g[i + 4] : g[i];
        w4[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 16);
      }
#pragma unroll
      for (int i = 0; i < 2; i++) {
        const float send = h8 ? w4[i] : w4[i + 2];
        const float keep = h8 ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 8);
      }
      {
        const float send = h4 ? w2[0] : w2[1];
        const float keep = h4 ? w2[1] : w2[0];
        z = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 4);
      }
      z += __shfl_xor_sync(0xFFFFFFFFu, z, 2);
      z += __shfl_xor_sync(0xFFFFFFFFu, z, 1);
      float z8 = g[8];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) z8 += __shfl_xor_sync(0xFFFFFFFFu, z8, o);
      // lanes 0,4,...,28 own slots 0..7, lane 1 owns slot 8: one red.global for all nine
      const uint32_t gid = __float_as_uint(s.rec[j][REC_F4].x);
      const bool owner = (lane & 3) == 0 || lane == 1;
      if (owner) {
        const int slot = lane == 1 ? 8 : (lane >> 2);
        atomicAdd(acc + (size_t)gid * ACC_STRIDE + slot, lane == 1 ? z8 : z);
      }
    }
  }
}

static cudaError_t launch_bwd(dim3 grid, const GeomState& g, const BinState& b, const ImgState& img, int W, int H,
                              const float* background, const float* dL_dpix, float* acc,
                              bool black_background, BlendBatch bb, cudaStream_t stream) {
  if (black_background)
    render_backward_kernel<false><<<grid, BL_THREADS, 0, stream>>>(
        img.ranges, b.point_list, W, H, background, g.rec, img.final_T,
        img.n_contrib, dL_dpix, acc, bb);
  else
    render_backward_kernel<true><<<grid, BL_THREADS, 0, stream>>>(
        img.ranges, b.point_list, W, H, background, g.rec, img.final_T,
        img.n_contrib, dL_dpix, acc, bb);
  DGE_LAUNCHED(1);
  return cudaGetLastError();
}

cudaError_t launch_render_backward(const ViewParams& vp, const GeomState& g, const BinState& b,
                                   const ImgState& img, const float* background,
                                   const float* dL_dpix, float* acc, bool black_background,
                                   cudaStream_t stream) {
  return launch_bwd(dim3(vp.grid_x, vp.grid_y), g, b, img, vp.W, vp.H, background, dL_dpix, acc, black_background,
                    BlendBatch{0, 0, nullptr, 0, (uint32_t)vp.P}, stream);
}

cudaError_t launch_render_backward_batched(const ViewParams& vp, const ViewBatch& vb, const GeomState& g0,
                                           const BinState& b, const ImgState& img0, const float* background,
                                           const float* dL_dpix, float* acc, size_t acc_stride_floats,
                                           bool black_background, cudaStream_t stream) {
  return launch_bwd(dim3(vp.grid_x, vp.grid_y, vb.V), g0, b, img0, vp.W, vp.H, background, dL_dpix, acc,
                    black_background, BlendBatch{vb.geom_stride, vb.img_stride, vb.seg_off, acc_stride_floats, (uint32_t)vp.P},
                    stream);
}

}  // namespace dge
