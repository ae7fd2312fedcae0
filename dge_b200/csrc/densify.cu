// SURVEY.md §8f N4 on the device: GaussianModel.densify_and_prune
// (gaussiansplatting/scene/gaussian_model.py:543-807: densify_and_clone :728-768, densify_and_split :675-726,
// prune_points :589-607, the optimiser surgery of cat_tensors_to_optimizer / _prune_optimizer :543-640) as TWO
// passes over the flat fit buffers instead of a dozen boolean-mask gathers and torch.cat re-allocations of every
// parameter and Adam-state tensor:
//   densify_select_kernel : per Gaussian, the clone / split / prune decisions and how many rows it contributes
//                           to the new model (itself, its clone, its split children);
//   densify_gather_kernel : given the exclusive scans of those counts, writes the new parameter / exp_avg /
//                           exp_avg_sq buffers and the new edit mask in the reference's row order
//                           [kept originals | kept clones | children copy 0 | copy 1 | ...].
// The arithmetic mirrors what torch evaluates on the same device (separately rounded products and sums, `tensor /
// python_scalar` as a multiplication by the rounded reciprocal), so decisions and copied rows are identical to
// fit.FitModel.densify_and_prune's torch path; the children's positions go through a 3x3 product whose
// accumulation order torch leaves to cuBLAS (they agree to an ulp).
#include "../../include/dge_b200.h"
#include "common.cuh"

namespace dge {

constexpr int DENS_GROUPS = 6;

struct DensifyBuffers {
  const float* src[3][DENS_GROUPS];  // [params, exp_avg, exp_avg_sq][group], group-major [P][width]
  float* dst[3][DENS_GROUPS];
  int width[DENS_GROUPS];
  int g_xyz, g_scaling, g_rotation;  // which groups need arithmetic for the split children
};

__device__ __forceinline__ float sigmoid_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

// sel bit 0: cloned, bit 1: split. keep[0][i]: the original stays, keep[1][i]: its clone stays, keep[2][i]: its
// children stay (each of the N copies), keep[3][i]: split-selected (rank among them indexes the normal samples).
__global__ void __launch_bounds__(256) densify_select_kernel(
    int P, const float* __restrict__ grad, const float* __restrict__ scaling_raw,
    const float* __restrict__ opacity_raw, const uint8_t* __restrict__ mask, float max_grad, float size_threshold,
    float min_opacity, float prune_size, float inv_split, int* __restrict__ keep, uint8_t* __restrict__ sel) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float g = grad[i];
  const float s0 = expf(scaling_raw[3 * i]), s1 = expf(scaling_raw[3 * i + 1]), s2 = expf(scaling_raw[3 * i + 2]);
  const float smax = fmaxf(fmaxf(s0, s1), s2);
  // densify_and_clone :730-735 (torch.norm of a 1-vector = |g|), densify_and_split :680-687
  const bool clone = fabsf(g) >= max_grad && smax <= size_threshold;
  const bool split = g >= max_grad && smax > size_threshold;
  // densify_and_prune :786-796: transparent, or (max_screen_size given) larger than a tenth of the scene; only
  // inside the edit mask. The children carry scaling = log(scale / (0.8 N)) (:694)
  const bool inside = mask == nullptr || mask[i] != 0;
  const bool transparent = sigmoid_ref(opacity_raw[i]) < min_opacity;
  const float c0 = expf(logf(__fmul_rn(s0, inv_split))), c1 = expf(logf(__fmul_rn(s1, inv_split))),
              c2 = expf(logf(__fmul_rn(s2, inv_split)));
  const bool big = prune_size >= 0.0f && smax > prune_size;
  const bool big_child = prune_size >= 0.0f && fmaxf(fmaxf(c0, c1), c2) > prune_size;
  const bool prune_own = inside && (transparent || big);
  const bool prune_child = inside && (transparent || big_child);
  keep[i] = (!split && !prune_own) ? 1 : 0;
  keep[P + i] = (clone && !prune_own) ? 1 : 0;
  keep[2 * P + i] = (split && !prune_child) ? 1 : 0;
  keep[3 * P + i] = split ? 1 : 0;
  sel[i] = (clone ? 1 : 0) | (split ? 2 : 0);
}

// One thread per (Gaussian, float column of its 59). scan: exclusive scans of keep[0..3].
__global__ void __launch_bounds__(256) densify_gather_kernel(
    int P, int cols, int N, const int* __restrict__ keep, const int* __restrict__ scan, int K_orig, int K_clone,
    int K_child, int K_split, DensifyBuffers b, const float* __restrict__ samples, float inv_split,
    const uint8_t* __restrict__ mask_in, uint8_t* __restrict__ mask_out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(t / cols), c = (int)(t - (size_t)i * cols);
  if (i >= P) return;
  int g = 0, k = c;
  while (k >= b.width[g]) k -= b.width[g++];
  const int w = b.width[g];
  const bool k_orig = keep[i] != 0, k_clone = keep[P + i] != 0, k_child = keep[2 * P + i] != 0;
  if (!(k_orig || k_clone || k_child)) return;
  const float p = b.src[0][g][(size_t)i * w + k];
  [[maybe_unused]] const size_t rows_out = (size_t)K_orig + K_clone + (size_t)N * K_child;
  if (k_orig) {
    DGE_CHECK(scan[i] < K_orig);
    const size_t d = (size_t)scan[i] * w + k;
    b.dst[0][g][d] = p;
    b.dst[1][g][d] = b.src[1][g][(size_t)i * w + k];  // the original keeps its Adam moments (_prune_optimizer :568-587)
    b.dst[2][g][d] = b.src[2][g][(size_t)i * w + k];
    if (c == 0 && mask_out) mask_out[scan[i]] = mask_in ? mask_in[i] : 1;
  }
  if (k_clone) {  // densify_and_clone: a copy with zero moments (cat_tensors_to_optimizer :609-640)
    const size_t row = (size_t)K_orig + scan[P + i];
    DGE_CHECK(scan[P + i] < K_clone && row < rows_out);
    const size_t d = row * w + k;
    b.dst[0][g][d] = p;
    b.dst[1][g][d] = 0.0f;
    b.dst[2][g][d] = 0.0f;
    if (c == 0 && mask_out) mask_out[row] = mask_in ? mask_in[i] : 1;
  }
  if (k_child) {  // densify_and_split :675-726
    const int rank = scan[3 * P + i];  // among the split-selected: row of its samples in each copy
    float vq[3] = {0.f, 0.f, 0.f};
    if (g == b.g_xyz) {
      // build_rotation (general_utils.py:78-100) on the RAW quaternion, products and sums rounded one by one
      // as torch's elementwise kernels do
      const float* qr = b.src[0][b.g_rotation] + 4 * (size_t)i;
      const float r0 = qr[0], r1 = qr[1], r2 = qr[2], r3 = qr[3];
      const float norm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0, r0), __fmul_rn(r1, r1)), __fmul_rn(r2, r2)),
                                              __fmul_rn(r3, r3)));
      const float r = __fdiv_rn(r0, norm), x = __fdiv_rn(r1, norm), y = __fdiv_rn(r2, norm), z = __fdiv_rn(r3, norm);
      auto m2 = [](float a, float bb) { return __fmul_rn(a, bb); };
      auto two = [](float a) { return __fmul_rn(2.0f, a); };
      if (k == 0) {
        vq[0] = __fadd_rn(1.0f, -two(__fadd_rn(m2(y, y), m2(z, z))));
        vq[1] = two(__fadd_rn(m2(x, y), -m2(r, z)));
        vq[2] = two(__fadd_rn(m2(x, z), m2(r, y)));
      } else if (k == 1) {
        vq[0] = two(__fadd_rn(m2(x, y), m2(r, z)));
        vq[1] = __fadd_rn(1.0f, -two(__fadd_rn(m2(x, x), m2(z, z))));
        vq[2] = two(__fadd_rn(m2(y, z), -m2(r, x)));
      } else {
        vq[0] = two(__fadd_rn(m2(x, z), -m2(r, y)));
        vq[1] = two(__fadd_rn(m2(y, z), m2(r, x)));
        vq[2] = __fadd_rn(1.0f, -two(__fadd_rn(m2(x, x), m2(y, y))));
      }
    }
    for (int n = 0; n < N; n++) {
      const size_t row = (size_t)K_orig + K_clone + (size_t)n * K_child + scan[2 * P + i];
      DGE_CHECK(scan[2 * P + i] < K_child && rank < K_split && row < rows_out);
      float v = p;
      if (g == b.g_xyz) {
        const float* s = samples + 3 * ((size_t)n * K_split + rank);
        v = __fadd_rn(__fmaf_rn(vq[2], s[2], __fmaf_rn(vq[1], s[1], __fmul_rn(vq[0], s[0]))), p);  // R s + xyz (:690-693)
      } else if (g == b.g_scaling) {
        v = logf(__fmul_rn(expf(p), inv_split));  // log(scale / (0.8 N)) (:694)
      }
      const size_t d = row * w + k;
      b.dst[0][g][d] = v;
      b.dst[1][g][d] = 0.0f;
      b.dst[2][g][d] = 0.0f;
      if (c == 0 && mask_out) mask_out[row] = mask_in ? mask_in[i] : 1;
    }
  }
}

}  // namespace dge

using namespace dge;

extern "C" {

int dge_densify_select(int P, const float* grad, const float* scaling_raw, const float* opacity_raw,
                       const uint8_t* mask, float max_grad, float size_threshold, float min_opacity,
                       float prune_size, int N, int* keep, uint8_t* sel, void* stream) {
  if (P == 0) return 0;
  const float inv_split = 1.0f / (float)(0.8 * N);
  densify_select_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      P, grad, scaling_raw, opacity_raw, mask, max_grad, size_threshold, min_opacity, prune_size, inv_split, keep, sel);
  DGE_LAUNCHED(1);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int dge_densify_gather(int P, int N, const int* keep, const int* scan, int K_orig, int K_clone, int K_child,
                       int K_split, const float* const* src, float* const* dst, const int* widths, int g_xyz,
                       int g_scaling, int g_rotation, const float* samples, const uint8_t* mask_in,
                       uint8_t* mask_out, void* stream) {
  if (P == 0) return 0;
  DensifyBuffers b;
  int cols = 0;
  for (int g = 0; g < DENS_GROUPS; g++) {
    b.width[g] = widths[g];
    cols += widths[g];
    for (int a = 0; a < 3; a++) {
      b.src[a][g] = src[a * DENS_GROUPS + g];
      b.dst[a][g] = dst[a * DENS_GROUPS + g];
    }
  }
  b.g_xyz = g_xyz;
  b.g_scaling = g_scaling;
  b.g_rotation = g_rotation;
  const size_t total = (size_t)P * cols;
  densify_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      P, cols, N, keep, scan, K_orig, K_clone, K_child, K_split, b, samples, 1.0f / (float)(0.8 * N), mask_in, mask_out);
  DGE_LAUNCHED(1);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
