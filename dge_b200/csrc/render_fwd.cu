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
//  * Gaussians are staged in batches of 128 records of 48 B.
#include "common.cuh"

namespace dge {

constexpr int RF_THREADS = 64;
constexpr int RF_BATCH = 128;

#define MUL(a, b) __fmul_rn((a), (b))
#define ADD(a, b) __fadd_rn((a), (b))
#define FMA(a, b, c) __fmaf_rn((a), (b), (c))

// Lower bound on `power` below which opacity*exp(power) < 1/255 for certain.
// 0.01 of slack in the exponent is ~1% in alpha; expf and __logf err by < 1e-6.
__device__ __forceinline__ float power_threshold(float opacity) {
  return opacity > 0.0f ? -(__logf(255.0f * opacity) + 0.01f) : __int_as_float(0x7f800000);
}

__global__ void __launch_bounds__(RF_THREADS) render_forward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
    const float2* __restrict__ means2D, const float4* __restrict__ conic_opacity,
    const float4* __restrict__ rgb_depth, const float* __restrict__ background,
    float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
    float* __restrict__ out_depth) {
  __shared__ float4 s_a[RF_BATCH];  // x, y, conic.x, conic.y
  __shared__ float4 s_b[RF_BATCH];  // conic.z, power threshold, opacity, unused
  __shared__ float4 s_c[RF_BATCH];  // r, g, b, depth

  const int tid = threadIdx.x;
  const int qx = tid & 7, qy = tid >> 3;
  const int px0 = blockIdx.x * DGE_TILE + 2 * qx, py0 = blockIdx.y * DGE_TILE + 2 * qy;
  const float fx0 = (float)px0, fx1 = (float)(px0 + 1), fy0 = (float)py0, fy1 = (float)(py0 + 1);
  // pixel p = 2*row + col inside the quad
  bool inside[4];
  inside[0] = px0 < W && py0 < H;
  inside[1] = px0 + 1 < W && py0 < H;
  inside[2] = px0 < W && py0 + 1 < H;
  inside[3] = px0 + 1 < W && py0 + 1 < H;

  const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];
  float T[4], C[4][3], Dp[4];
  uint32_t last[4];
  bool done[4];
#pragma unroll
  for (int p = 0; p < 4; p++) {
    T[p] = 1.0f;
    C[p][0] = C[p][1] = C[p][2] = 0.0f;
    Dp[p] = 0.0f;
    last[p] = 0;
    done[p] = !inside[p];
  }

  for (uint32_t base = range.x; base < range.y; base += RF_BATCH) {
    const bool all_done = done[0] && done[1] && done[2] && done[3];
    if (__syncthreads_and(all_done)) break;
    const int count = min((uint32_t)RF_BATCH, range.y - base);
    for (int k = tid; k < count; k += RF_THREADS) {
      const uint32_t gid = point_list[base + k];
      const float2 xy = means2D[gid];
      const float4 co = conic_opacity[gid];
      s_a[k] = make_float4(xy.x, xy.y, co.x, co.y);
      s_b[k] = make_float4(co.z, power_threshold(co.w), co.w, 0.0f);
      s_c[k] = rgb_depth[gid];
    }
    __syncthreads();
    if (!all_done) {
      for (int j = 0; j < count; j++) {
        const float4 a = s_a[j];
        const float2 b = *reinterpret_cast<const float2*>(&s_b[j]);
        const float dx0 = ADD(a.x, -fx0), dx1 = ADD(a.x, -fx1);
        const float dy0 = ADD(a.y, -fy0), dy1 = ADD(a.y, -fy1);
        const float bx0 = MUL(dx0, a.z), bx1 = MUL(dx1, a.z);  // conic.x * dx
        const float cx0 = MUL(dx0, a.w), cx1 = MUL(dx1, a.w);  // conic.y * dx
        const float ay0 = MUL(dy0, MUL(dy0, b.x)), ay1 = MUL(dy1, MUL(dy1, b.x));
        float power[4];
        power[0] = FMA(FMA(dx0, bx0, ay0), -0.5f, -MUL(dy0, cx0));
        power[1] = FMA(FMA(dx1, bx1, ay0), -0.5f, -MUL(dy0, cx1));
        power[2] = FMA(FMA(dx0, bx0, ay1), -0.5f, -MUL(dy1, cx0));
        power[3] = FMA(FMA(dx1, bx1, ay1), -0.5f, -MUL(dy1, cx1));
        bool cand[4];
        bool any = false;
#pragma unroll
        for (int p = 0; p < 4; p++) {
          cand[p] = !done[p] && !(power[p] > 0.0f) && !(power[p] < b.y);
          any |= cand[p];
        }
        if (!any) continue;
        const float opacity = s_b[j].z;
        const float4 cd = s_c[j];
#pragma unroll
        for (int p = 0; p < 4; p++) {
          if (!cand[p]) continue;
          const float alpha = fminf(0.99f, MUL(opacity, expf(power[p])));
          if (alpha < 1.0f / 255.0f) continue;
          const float test_T = MUL(T[p], ADD(1.0f, -alpha));
          if (test_T < 0.0001f) {
            done[p] = true;
            continue;
          }
          C[p][0] = FMA(T[p], MUL(alpha, cd.x), C[p][0]);
          C[p][1] = FMA(T[p], MUL(alpha, cd.y), C[p][1]);
          C[p][2] = FMA(T[p], MUL(alpha, cd.z), C[p][2]);
          Dp[p] = FMA(T[p], MUL(alpha, cd.w), Dp[p]);
          T[p] = test_T;
          last[p] = base - range.x + j + 1;
        }
      }
    }
  }

  const float bg0 = __ldg(background), bg1 = __ldg(background + 1), bg2 = __ldg(background + 2);
  const size_t HW = (size_t)H * W;
#pragma unroll
  for (int p = 0; p < 4; p++) {
    if (!inside[p]) continue;
    const size_t pix = (size_t)(py0 + (p >> 1)) * W + (px0 + (p & 1));
    final_T[pix] = T[p];
    n_contrib[pix] = last[p];
    out_color[pix] = FMA(bg0, T[p], C[p][0]);
    out_color[HW + pix] = FMA(bg1, T[p], C[p][1]);
    out_color[2 * HW + pix] = FMA(bg2, T[p], C[p][2]);
    out_depth[pix] = Dp[p];
  }
}

cudaError_t launch_render_forward(const ViewParams& vp, const GeomState& g, const BinState& b,
                                  ImgState& img, const float* background, float* out_color,
                                  float* out_depth, cudaStream_t stream) {
  dim3 grid(vp.grid_x, vp.grid_y);
  render_forward_kernel<<<grid, RF_THREADS, 0, stream>>>(
      img.ranges, b.point_list, vp.W, vp.H, g.means2D, g.conic_opacity, g.rgb_depth, background,
      img.final_T, img.n_contrib, out_color, out_depth);
  DGE_LAUNCHED(1);
  return cudaGetLastError();
}

}  // namespace dge
