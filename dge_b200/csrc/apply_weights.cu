// K14: DGE's local-editing mask back-projection.
//
// Replaces renderCUDA_apply_weights<CH> (DGR/cuda_rasterizer/apply_weights.cu:239-356):
// walk each tile's list front to back with exactly the forward blend's alpha / T tests and,
// for every contributing (pixel, Gaussian) pair, add the pixel's mask value to
// weights[gid*CH+ch] and 1 to cnt[gid] once PER CHANNEL (apply_weights.cu:331-334).
// The reference issues CH float + CH int atomics per pair; here each thread owns a 2x2
// quad, sums its pixels, the warp reduces (REDUX for the count, shuffles for the floats)
// and one lane per value issues the atomic. With 0/1 masks the float sums are integers
// below 2^24, so the result is bit-identical to the reference whatever the order.
// The out-of-image read of image_weights in the reference (apply_weights.cu:279-283, before
// its `inside` test) is guarded here; those values are never used.
// Staged batch of the back-projection walk: 64 records like the forward blend (it stops early too; config 3:
// 7.08 -> 6.93 ms per 40 views against 128)
#ifndef DGE_BL_BATCH
#define DGE_BL_BATCH 64
#endif
#include "blend.cuh"

namespace dge {

template <int CH>
__global__ void __launch_bounds__(BL_THREADS) apply_weights_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
    const float4* __restrict__ rec, const float* __restrict__ image_weights, float* __restrict__ weights,
    int* __restrict__ cnt, BlendBatch bb) {
  __shared__ BlendSmem s;
  if (bb.seg_off) {  // view blockIdx.z of a batch: its own mask image, the SAME weights / cnt
    const size_t gs = blockIdx.z * bb.geom_stride, is = blockIdx.z * bb.img_stride;
    ranges = shift_ptr(ranges, is);
    rec = shift_ptr(rec, gs);
    point_list += bb.seg_off[blockIdx.z];
    image_weights += (size_t)blockIdx.z * CH * H * W;
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // pixel p of this thread = (px0 + PX_STEP*(p&1), py0 + PY_STEP*(p>>1)): one pixel in each 8x4
  // quadrant of the warp's 16x8 half-tile (see blend.cuh)
  const int px0 = blockIdx.x * DGE_TILE + (lane & 7), py0 = blockIdx.y * DGE_TILE + 8 * warp + (lane >> 3);
  const float fx0 = (float)px0, fx1 = (float)(px0 + PX_STEP), fy0 = (float)py0, fy1 = (float)(py0 + PY_STEP);
  const float X0 = (float)(blockIdx.x * DGE_TILE);
  const float Y0 = (float)(blockIdx.y * DGE_TILE + 8 * warp);
  const size_t HW = (size_t)H * W;
  const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];

  float T[4], Cw[4][CH];
  bool done[4];
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int x = px0 + PX_STEP * (p & 1), y = py0 + PY_STEP * (p >> 1);
    const bool inside = x < W && y < H;
    T[p] = 1.0f;
    done[p] = !inside;
#pragma unroll
    for (int c = 0; c < CH; c++) Cw[p][c] = inside ? image_weights[c * HW + (size_t)y * W + x] : 0.0f;
  }

  stage_init(s, tid);
  uint32_t parity = 0;
  for (uint32_t base = range.x; base < range.y; base += BL_BATCH) {
    const bool all_done = done[0] && done[1] && done[2] && done[3];
    if (__syncthreads_and(all_done)) break;
    const int count = min((uint32_t)BL_BATCH, range.y - base);
    stage_batch(s, tid, count, [&](int k) { return base + k; }, point_list, rec, parity, nullptr, bb.P);
    parity ^= 1u;
    if (__all_sync(0xFFFFFFFFu, all_done)) continue;
    uint32_t live = 0;
#pragma unroll
    for (int p = 0; p < 4; p++) live |= __all_sync(0xFFFFFFFFu, done[p]) ? 0u : (1u << p);
    const int n = compact_batch(s, warp, lane, count, X0, Y0, [&](int) { return live; });
    for (int i = 0; i < n; i++) {
      const uint32_t e = s.list[warp][i];  // warp-uniform
      const int j = e & 0xFF;
      const float4 a = s.rec[j][0];  // x, y, conic.x, conic.y
      const float4 b = s.rec[j][1];  // conic.z, power threshold, opacity, -
      float wsum[CH];
#pragma unroll
      for (int c = 0; c < CH; c++) wsum[c] = 0.0f;
      int nhit = 0;
#pragma unroll
      for (int p = 0; p < 4; p++) {
        if (!(e & (0x100u << p))) continue;  // warp-uniform branch
        const float power = pixel_power(a, b.x, BADD(a.x, (p & 1) ? -fx1 : -fx0), BADD(a.y, (p >> 1) ? -fy1 : -fy0));
        if (done[p] || power > 0.0f || power < b.y) continue;
        const float alpha = fminf(0.99f, BMUL(b.z, expf(power)));
        if (alpha < 1.0f / 255.0f) continue;
        const float test_T = BMUL(T[p], BADD(1.0f, -alpha));
        if (test_T < 0.0001f) {
          done[p] = true;
          continue;
        }
#pragma unroll
        for (int c = 0; c < CH; c++) wsum[c] += Cw[p][c];
        nhit += 1;
        T[p] = test_T;
      }
      const int total = __reduce_add_sync(0xFFFFFFFFu, nhit);
      if (total == 0) continue;
#pragma unroll
      for (int c = 0; c < CH; c++)
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) wsum[c] += __shfl_xor_sync(0xFFFFFFFFu, wsum[c], o);
      const uint32_t gid = __float_as_uint(s.rec[j][REC_F4].x);
      if (lane == 0) atomicAdd(cnt + gid, total * CH);
#pragma unroll
      for (int c = 0; c < CH; c++)
        if (lane == c + 1) atomicAdd(weights + (size_t)gid * CH + c, wsum[c]);
    }
  }
}

cudaError_t launch_apply_weights_render(const ViewParams& vp, const GeomState& g, const BinState& b,
                                        const ImgState& img, float* weights, int* cnt,
                                        const float* image_weights, int num_channels,
                                        cudaStream_t stream) {
  return launch_apply_weights_render_batched(vp, nullptr, g, b, img, weights, cnt, image_weights, num_channels, stream);
}

// vb == nullptr: one view; otherwise all views of the batch (grid.z), image_weights [V,CH,H,W]
cudaError_t launch_apply_weights_render_batched(const ViewParams& vp, const ViewBatch* vb, const GeomState& g,
                                                const BinState& b, const ImgState& img, float* weights, int* cnt,
                                                const float* image_weights, int num_channels,
                                                cudaStream_t stream) {
  dim3 grid(vp.grid_x, vp.grid_y, vb ? vb->V : 1);
  const BlendBatch bb = vb ? BlendBatch{vb->geom_stride, vb->img_stride, vb->seg_off, 0, (uint32_t)vp.P} : BlendBatch{0, 0, nullptr, 0, (uint32_t)vp.P};
#define AW_LAUNCH(CH)                                                                          \
  apply_weights_kernel<CH><<<grid, BL_THREADS, 0, stream>>>(img.ranges, b.point_list, vp.W, vp.H, \
                                                            g.rec, image_weights, weights, cnt, bb)
  if (num_channels == 1) AW_LAUNCH(1);
  else if (num_channels == 2) AW_LAUNCH(2);
  else if (num_channels == 3) AW_LAUNCH(3);
  else return cudaErrorInvalidValue;  // the reference prints and exit(-1)s (apply_weights.cu:377-380)
#undef AW_LAUNCH
  DGE_LAUNCHED(1);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ fused Adam (N3) -----
// torch.optim.Adam (gaussiansplatting/scene/gaussian_model.py:374): no weight decay, no
// amsgrad; bias corrections as torch's single-tensor path:
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void __launch_bounds__(256) fused_adam_kernel(float* __restrict__ p,
                                                         const float* __restrict__ g,
                                                         float* __restrict__ m, float* __restrict__ v,
                                                         size_t n, float step_size, float inv_bc2_sqrt,
                                                         float beta1, float beta2, float eps,
                                                         const uint8_t* __restrict__ mask, int stride,
                                                         size_t first) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float grad = g[i];
  if (mask != nullptr && !mask[(first + i) / stride]) grad = 0.0f;
  const float mi = beta1 * m[i] + (1.0f - beta1) * grad;
  const float vi = beta2 * v[i] + (1.0f - beta2) * grad * grad;
  m[i] = mi;
  v[i] = vi;
  p[i] -= step_size * (mi / (sqrtf(vi) * inv_bc2_sqrt + eps));
}

// four consecutive elements per thread through 16-byte accesses (pointers 16-byte aligned)
__global__ void __launch_bounds__(256) fused_adam_vec4_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                              float4* __restrict__ m, float4* __restrict__ v,
                                                              size_t n4, float step_size, float inv_bc2_sqrt,
                                                              float beta1, float beta2, float eps,
                                                              const uint8_t* __restrict__ mask, int stride) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 G = g[i], M = m[i], Vv = v[i], Pp = p[i];
  float* gr = reinterpret_cast<float*>(&G);
  float* mm = reinterpret_cast<float*>(&M);
  float* vv = reinterpret_cast<float*>(&Vv);
  float* pp = reinterpret_cast<float*>(&Pp);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    float grad = gr[k];
    if (mask != nullptr && !mask[(4 * i + k) / stride]) grad = 0.0f;
    mm[k] = beta1 * mm[k] + (1.0f - beta1) * grad;
    vv[k] = beta2 * vv[k] + (1.0f - beta2) * grad * grad;
    pp[k] -= step_size * (mm[k] / (sqrtf(vv[k]) * inv_bc2_sqrt + eps));
  }
  m[i] = M;
  v[i] = Vv;
  p[i] = Pp;
}

cudaError_t launch_fused_adam(float* param, const float* grad, float* m, float* v, size_t n, float lr,
                              float beta1, float beta2, float eps, int step, const uint8_t* mask,
                              int stride, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  const int st = stride > 0 ? stride : 1;
  const bool aligned = ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                         reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  size_t done = 0;
  if (aligned && n >= 4) {
    const size_t n4 = n / 4;
    fused_adam_vec4_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<float4*>(param), reinterpret_cast<const float4*>(grad), reinterpret_cast<float4*>(m),
        reinterpret_cast<float4*>(v), n4, step_size, inv_bc2_sqrt, beta1, beta2, eps, mask, st);
    DGE_LAUNCHED(1);
    done = n4 * 4;
  }
  if (done < n) {
    // tail (or everything, for unaligned blocks); the mask index needs the element's global position
    fused_adam_kernel<<<(unsigned)((n - done + 255) / 256), 256, 0, stream>>>(
        param + done, grad + done, m + done, v + done, n - done, step_size, inv_bc2_sqrt, beta1, beta2, eps,
        mask, st, done);
    DGE_LAUNCHED(1);
  }
  return cudaGetLastError();
}


// Fused L1 loss and its gradient for one rendered view (DGE.py:672: 10 * L1 over the batch):
// grad = scale * sign(image - target), *loss_accum += scale * sum|image - target|.
// Replaces sub / abs / sum / mul and their four backward kernels.
template <bool VEC>
__global__ void __launch_bounds__(256) l1_loss_grad_kernel(const float* __restrict__ image,
                                                           const float* __restrict__ target, size_t n,
                                                           float scale, float* __restrict__ grad,
                                                           float* __restrict__ loss_accum) {
  float local = 0.f;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
  auto one = [&](float a, float b) {
    const float d = a - b;
    local += fabsf(d);
    return d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
  };
  size_t done = 0;
  if (VEC) {  // all three pointers 16-byte aligned: 128-bit accesses, the n % 4 tail below
    const size_t n4 = n / 4;
    for (size_t i = tid; i < n4; i += nthreads) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(image) + i);
      const float4 b = __ldg(reinterpret_cast<const float4*>(target) + i);
      reinterpret_cast<float4*>(grad)[i] = make_float4(one(a.x, b.x), one(a.y, b.y), one(a.z, b.z), one(a.w, b.w));
    }
    done = n4 * 4;
  }
  for (size_t i = done + tid; i < n; i += nthreads) grad[i] = one(image[i], target[i]);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, o);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += ws[i];
    atomicAdd(loss_accum, t * scale);
  }
}

cudaError_t launch_l1_loss_grad(const float* image, const float* target, size_t n, float scale,
                                float* grad, float* loss_accum, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((n + 256 * 8 - 1) / (256 * 8));
  const bool vec = ((reinterpret_cast<uintptr_t>(image) | reinterpret_cast<uintptr_t>(target) |
                     reinterpret_cast<uintptr_t>(grad)) & 15) == 0;
  if (vec)
    l1_loss_grad_kernel<true><<<blocks < 1 ? 1 : blocks, 256, 0, stream>>>(image, target, n, scale, grad, loss_accum);
  else
    l1_loss_grad_kernel<false><<<blocks < 1 ? 1 : blocks, 256, 0, stream>>>(image, target, n, scale, grad, loss_accum);
  DGE_LAUNCHED(1);
  return cudaGetLastError();
}

// Densification statistics of one step (threestudio/systems/DGE.py:266-284 + GaussianModel.add_densification_stats,
// gaussiansplatting/scene/gaussian_model.py:811-815) for the whole model in one pass: for the Gaussians some view of
// the step saw (max over the views of radii > 0): max_radii2D = max(max_radii2D, radii), xyz_gradient_accum +=
// |screen-space gradient (x, y)|, denom += 1. Replaces ten elementwise / reduce launches of torch.
__global__ void __launch_bounds__(256) update_stats_kernel(int P, const int* __restrict__ radii_max,
                                                           const float* __restrict__ m2d_grad,
                                                           int* __restrict__ max_radii2D,
                                                           float* __restrict__ xyz_gradient_accum,
                                                           float* __restrict__ denom) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int r = radii_max[i];
  if (r <= 0) return;
  max_radii2D[i] = max(max_radii2D[i], r);
  const float gx = m2d_grad[3 * (size_t)i], gy = m2d_grad[3 * (size_t)i + 1];
  xyz_gradient_accum[i] += sqrtf(gx * gx + gy * gy);
  denom[i] += 1.0f;
}

cudaError_t launch_update_stats(int P, const int* radii_max, const float* m2d_grad, int* max_radii2D,
                                float* xyz_gradient_accum, float* denom, cudaStream_t stream) {
  if (P == 0) return cudaSuccess;
  update_stats_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, radii_max, m2d_grad, max_radii2D, xyz_gradient_accum,
                                                           denom);
  DGE_LAUNCHED(1);
  return cudaGetLastError();
}

}  // namespace dge
