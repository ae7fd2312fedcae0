// Probe (not part of the library): what does staging Gaussian records into shared memory cost when it is
// done (a) with LDG.128 gathers + STS, as blend.cuh:stage_batch does, and (b) with one TMA bulk copy
// (cp.async.bulk.shared.global, 64 B) per record from a packed per-Gaussian record array, completion
// on an mbarrier? Same access pattern as the forward blend: one CTA of 64 threads per "tile", batches
// of 128 records, ids taken from a random list over P Gaussians (the 64 MB record array sits in L2).
// The kernels only stage (and touch one word per record so the copy cannot be elided).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_gather_probe tma_gather_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

constexpr int BATCH = 128, THREADS = 64;

__global__ void __launch_bounds__(THREADS) stage_ldg(const uint32_t* __restrict__ ids, int per_tile,
                                                     const float4* __restrict__ a, const float4* __restrict__ b,
                                                     const float4* __restrict__ c, float* __restrict__ out) {
  __shared__ float4 sa[BATCH], sb[BATCH], sc[BATCH];
  const uint32_t* list = ids + (size_t)blockIdx.x * per_tile;
  float acc = 0.f;
  for (int base = 0; base < per_tile; base += BATCH) {
    __syncthreads();
    for (int k = threadIdx.x; k < BATCH; k += THREADS) {
      const uint32_t g = list[base + k];
      sa[k] = a[g];
      sb[k] = b[g];
      sc[k] = c[g];
    }
    __syncthreads();
    for (int k = threadIdx.x; k < BATCH; k += THREADS) acc += sa[k].x + sb[k].y + sc[k].z;
  }
  if (acc == 123.456f) out[blockIdx.x] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(THREADS) stage_tma(const uint32_t* __restrict__ ids, int per_tile,
                                                     const float4* __restrict__ rec /* 4 float4 per Gaussian */,
                                                     float* __restrict__ out) {
  __shared__ __align__(128) float4 srec[2][BATCH * 4];
  __shared__ __align__(8) uint64_t bar[2];
  const uint32_t* list = ids + (size_t)blockIdx.x * per_tile;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; s++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  const int nb = per_tile / BATCH;
  auto issue = [&](int bi) {
    const int s = bi & 1;
    if (threadIdx.x == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])),
                   "r"(BATCH * 64));
    __syncthreads();  // expect_tx before any copy can complete
    for (int k = threadIdx.x; k < BATCH; k += THREADS) {
      const uint32_t g = list[bi * BATCH + k];
      asm volatile(
          "cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];" ::"r"(
              smem_u32(&srec[s][k * 4])),
          "l"(rec + (size_t)g * 4), "r"(smem_u32(&bar[s]))
          : "memory");
    }
  };
  float acc = 0.f;
  issue(0);
  for (int bi = 0; bi < nb; bi++) {
    if (bi + 1 < nb) issue(bi + 1);  // double buffering: next batch in flight while this one is consumed
    const int s = bi & 1;
    const uint32_t parity = (bi >> 1) & 1;
    uint32_t done = 0;
    while (!done)
      asm volatile(
          "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
          : "=r"(done)
          : "r"(smem_u32(&bar[s])), "r"(parity)
          : "memory");
    for (int k = threadIdx.x; k < BATCH; k += THREADS)
      acc += srec[s][k * 4].x + srec[s][k * 4 + 1].y + srec[s][k * 4 + 2].z;
    __syncthreads();  // everyone is done with buffer s before it is refilled
  }
  if (acc == 123.456f) out[blockIdx.x] = acc;
}

int main(int argc, char** argv) {
  const int P = 1000000, tiles = 20480, per_tile = argc > 1 ? atoi(argv[1]) : 1024;  // 21 M records staged
  std::vector<uint32_t> ids((size_t)tiles * per_tile);
  uint64_t x = 88172645463325252ull;
  for (auto& v : ids) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = (uint32_t)(x % P); }
  uint32_t* d_ids; float4 *a, *b, *c, *rec; float* out;
  cudaMalloc(&d_ids, ids.size() * 4);
  cudaMemcpy(d_ids, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&a, (size_t)P * 16); cudaMalloc(&b, (size_t)P * 16); cudaMalloc(&c, (size_t)P * 16);
  cudaMalloc(&rec, (size_t)P * 64); cudaMalloc(&out, tiles * 4);
  cudaMemset(a, 0, (size_t)P * 16); cudaMemset(b, 0, (size_t)P * 16); cudaMemset(c, 0, (size_t)P * 16);
  cudaMemset(rec, 0, (size_t)P * 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    stage_ldg<<<tiles, THREADS>>>(d_ids, per_tile, a, b, c, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    if (rep == 2) printf("LDG gather + STS : %.3f ms for %zu records (%.1f ns/record/SM-equivalent, 48 B each)\n", ms, ids.size(), ms * 1e6 / ids.size() * 148);
    cudaEventRecord(e0);
    stage_tma<<<tiles, THREADS>>>(d_ids, per_tile, rec, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    if (rep == 2) printf("TMA bulk 64 B/rec: %.3f ms for %zu records (double buffered, mbarrier)\n", ms, ids.size());
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
