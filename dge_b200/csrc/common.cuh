// Shared declarations of the dge_b200 CUDA library (sm_100a only).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#define DGE_TILE 16            // DGR/cuda_rasterizer/config.h:16-17 (BLOCK_X/BLOCK_Y)
#define DGE_NUM_SMS 148

// Checked build (python -m dge_b200.build --variant checks -DDGE_CHECKS=1; tests/gpu_checked_build.sh): bounds
// assertions on every scattered global store of the sort / partition / expand kernels and on every gathered
// index of the blend staging. compute-sanitizer is closed on the GPU pool this was developed on, so the GPU
// suite is run once against this build instead; a failure prints its site and traps.
#ifndef DGE_CHECKS
#define DGE_CHECKS 0
#endif
#if DGE_CHECKS
#include <cstdio>
#define DGE_CHECK(cond)                                                                          \
  do {                                                                                           \
    if (!(cond)) {                                                                               \
      printf("DGE_CHECK failed: %s at %s:%d (block %d,%d,%d thread %d)\n", #cond, __FILE__, __LINE__, \
             blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);                                   \
      __trap();                                                                                  \
    }                                                                                            \
  } while (0)
#else
#define DGE_CHECK(cond) ((void)0)
#endif

namespace dge {

// Number of kernels this library has launched (read by bench.py through dge_launch_count()).
extern std::atomic<unsigned long long> g_kernel_launches;  // (forward and autograd threads both launch)
#define DGE_LAUNCHED(n) (::dge::g_kernel_launches.fetch_add((n), std::memory_order_relaxed))

// ---------------------------------------------------------------- scratch ---
// Our own layout of the three opaque blobs (the reference's is
// DGR/cuda_rasterizer/rasterizer_impl.cu:135-175). All sub-arrays 256-B aligned.
struct GeomState {
  float4* rec;            // [P][4] one 64-byte blend record per Gaussian (REC_* below): what the blend
                          //     kernels stage into shared memory with ONE TMA bulk copy per instance
  ushort4* rect;          // [P] tile rect min.x, min.y, max.x, max.y (all 0 <=> culled)
  uint8_t* clamped;       // [P] bit ch set <=> SH colour channel clamped at 0
  uint32_t* sort_key[2];  // [P] depth bits (0xFFFFFFFF for culled), ping-pong
  uint32_t* sort_val[2];  // [P] Gaussian ids, ping-pong; sort_val[0] ends up depth-sorted
  uint32_t* offsets;      // [P] inclusive scan of tiles_touched in depth order
  uint32_t* block_sums;   // [scan blocks]
  uint32_t* counters;     // [64] counters[0] = num_rendered
  uint32_t* sort_ws;      // radix-sort workspace (tickets, histograms, look-back status)
  size_t sort_ws_bytes;
};

struct BinState {
  uint32_t* point_list;   // [R] Gaussian ids sorted by (tile, depth, id)   (= val[0])
  uint32_t* tile_ids;     // [R] tile id of each sorted instance            (= key[0])
  uint32_t* key_alt;      // [R] ping-pong
  uint32_t* val_alt;      // [R] ping-pong
  uint32_t* sort_ws;
  size_t sort_ws_bytes;
};

struct ImgState {
  float* final_T;         // [N]
  uint32_t* n_contrib;    // [N]
  uint2* ranges;          // [T]
};

// The same arrays `bytes` further on: view v of a batch whose per-view blobs are laid out with a
// uniform stride (fit step, all views of a step in one launch).
template <typename T>
__host__ __device__ __forceinline__ T* shift_ptr(T* p, size_t bytes) {
  return (T*)((const char*)p + bytes);
}
__host__ __device__ __forceinline__ GeomState shift_geom(GeomState g, size_t bytes) {
  g.rec = shift_ptr(g.rec, bytes);
  g.rect = shift_ptr(g.rect, bytes);
  g.clamped = shift_ptr(g.clamped, bytes);
  g.sort_key[0] = shift_ptr(g.sort_key[0], bytes);
  g.sort_key[1] = shift_ptr(g.sort_key[1], bytes);
  g.sort_val[0] = shift_ptr(g.sort_val[0], bytes);
  g.sort_val[1] = shift_ptr(g.sort_val[1], bytes);
  g.offsets = shift_ptr(g.offsets, bytes);
  g.block_sums = shift_ptr(g.block_sums, bytes);
  g.counters = shift_ptr(g.counters, bytes);
  g.sort_ws = shift_ptr(g.sort_ws, bytes);
  return g;
}
__host__ __device__ __forceinline__ ImgState shift_img(ImgState s, size_t bytes) {
  s.final_T = shift_ptr(s.final_T, bytes);
  s.n_contrib = shift_ptr(s.n_contrib, bytes);
  s.ranges = shift_ptr(s.ranges, bytes);
  return s;
}

// Views of one fit step processed by single launches (grid dimension = view). Per-view geometry and
// image blobs have uniform strides; the instance lists of the views lie back to back in ONE binning
// arena, view v at element offset seg_off[v] (seg_off[V] = total number of instances).
struct ViewBatch {
  int V;
  size_t geom_stride;       // bytes between the geometry blobs of consecutive views
  size_t img_stride;        // bytes between the image blobs
  const uint32_t* seg_off;  // device, [V+1]
  const float* cams;        // device, V records of 40 floats (include/dge_b200.h "fit step")
};

size_t carve_geom(char* base, int P, GeomState* st);
size_t carve_binning(char* base, int R, int width, int height, BinState* st);
size_t carve_image(char* base, int width, int height, ImgState* st);

// ------------------------------------------------------------------- sort ---
// Stable LSD onesweep radix sort of (u32 key, u32 value) pairs on key bits
// [0, num_bits). Result lands in (keys[0], vals[0]) when the number of passes
// is even, which sort_pairs guarantees by issuing a plain copy pass otherwise.
// vals[0] == nullptr on entry means "values are 0..n-1" (no iota kernel).
size_t sort_workspace_bytes(uint32_t n);
cudaError_t sort_pairs(uint32_t* keys[2], uint32_t* vals[2], uint32_t n, int num_bits,
                       bool iota_values, uint32_t* ws, size_t ws_bytes, cudaStream_t stream);
int sort_num_passes(int num_bits);
// The same sort over `segs` independent segments in ONE launch per pass (grid.y = segment).
//  * seg_off == nullptr: uniform segments of n items; segment s of every array (keys, vals, ws) lies
//    s * seg_stride_bytes after segment 0 (the per-view geometry blobs of a fit step);
//  * seg_off != nullptr: segment s is elements [seg_off[s], seg_off[s+1]) of the arrays (the views'
//    instance lists inside one binning arena); n = the largest segment, n_total = seg_off[segs]; the
//    workspace is shared: sort_workspace_bytes_segmented(n_total, segs).
size_t sort_workspace_bytes_segmented(uint32_t n_total, int segs);
cudaError_t sort_pairs_segmented(uint32_t* keys[2], uint32_t* vals[2], uint32_t n, int num_bits,
                                 bool iota_values, uint32_t* ws, size_t ws_bytes, int segs,
                                 size_t seg_stride_bytes, const uint32_t* seg_off, uint32_t n_total,
                                 cudaStream_t stream);

// ----------------------------------------------------------------- stages ---
struct ViewParams {
  const float* view;    // device, 16 floats, flat index [4*col+row] (auxiliary.h:58-77)
  const float* proj;    // device, 16 floats
  const float* campos;  // device, 3 floats
  float tan_fovx, tan_fovy, focal_x, focal_y;
  float scale_modifier;
  int W, H, grid_x, grid_y;
  int P, D, M;
};

// colors_mode: 0 = SH -> rgb, 1 = copy colors_precomp into rgb_depth, 2 = no colour (apply_weights)
cudaError_t launch_preprocess(const ViewParams& vp, const float* means3D, const float* scales,
                              const float* rotations, const float* opacities, const float* shs,
                              const float* cov3D_precomp, const float* colors_precomp,
                              int colors_mode, bool prefiltered, int* radii, GeomState& g,
                              uint8_t* flags_out, cudaStream_t stream);
// fit step: per-Gaussian backward of V views in one pass (cams: V records of 40 floats)
cudaError_t launch_geom_backward_batched(int P, int D, int M, int V, const float* cams, int W, int H,
                                         float scale_modifier, const float* acc, size_t acc_stride,
                                         const uint8_t* flags, size_t flags_stride,
                                         const float* means3D, const float* shs, const float* scales,
                                         const float* rotations, float* dL_dmean3D, float* dL_dmean2D,
                                         float* dL_dsh, float* dL_dopacity, float* dL_dscale,
                                         float* dL_drot, bool accumulate, cudaStream_t stream,
                                         const float* opacities = nullptr,
                                         const float* rotation_raw = nullptr, float* dL_drest = nullptr);
cudaError_t launch_activate(int P, const float* f_dc, const float* f_rest, const float* opacity_raw,
                            const float* scaling_raw, const float* rotation_raw, float* shs,
                            float* opacities, float* scales, float* rotations, cudaStream_t stream);
cudaError_t launch_mark_visible(int P, const float* means3D, const float* view, const float* proj,
                                uint8_t* present, cudaStream_t stream);
// depth sort -> scan -> expand -> tile sort -> ranges. R already known on host.
cudaError_t launch_depth_sort(int P, GeomState& g, cudaStream_t stream);
cudaError_t launch_binning(const ViewParams& vp, int R, GeomState& g, BinState& b, ImgState& img,
                           cudaStream_t stream);
cudaError_t launch_render_forward(const ViewParams& vp, const GeomState& g, const BinState& b,
                                  ImgState& img, const float* background, float* out_color,
                                  float* out_depth, cudaStream_t stream);
// acc: [P][12] floats, zeroed by the caller: the ACC_* slots below
cudaError_t launch_render_backward(const ViewParams& vp, const GeomState& g, const BinState& b,
                                   const ImgState& img, const float* background,
                                   const float* dL_dpix, float* acc, bool black_background,
                                   cudaStream_t stream);
cudaError_t launch_geom_backward(const ViewParams& vp, const float* means3D, const float* scales,
                                 const float* rotations, const float* shs, const float* cov3D_precomp,
                                 const int* radii, const GeomState& g, const float* acc,
                                 float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                                 float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                                 float* dL_dscale, float* dL_drot, bool accumulate, cudaStream_t stream);
// fit step: all V views of a step per launch (preprocess.cu, binning.cu, render_fwd.cu, render_bwd.cu)
cudaError_t launch_preprocess_batched(const ViewParams& vp, const ViewBatch& vb, const float* means3D,
                                      const float* scales, const float* rotations, const float* opacities,
                                      const float* shs, GeomState& g0, uint8_t* flags, size_t flags_stride,
                                      int* radii_max, bool prune_lists, cudaStream_t stream);
// colours (SH -> rgb + clamp bits) of the visible (view, Gaussian) pairs into the records of a batched preprocess
// that ran without SH (geometry only)
cudaError_t launch_colour_batched(int P, int D, int V, const float* cams, const float* means3D, const float* shs,
                                  GeomState& g0, size_t geom_stride, uint8_t* flags, size_t flags_stride,
                                  cudaStream_t stream);
cudaError_t launch_seg_offsets(const ViewBatch& vb, const GeomState& g0, uint32_t* seg_off, cudaStream_t stream);
cudaError_t launch_depth_sort_batched(int P, const ViewBatch& vb, GeomState& g0, cudaStream_t stream);
cudaError_t launch_binning_batched(const ViewParams& vp, const ViewBatch& vb, uint32_t R_total, uint32_t R_max,
                                   GeomState& g0, BinState& b, ImgState& img0, cudaStream_t stream);
cudaError_t launch_render_forward_batched(const ViewParams& vp, const ViewBatch& vb, const GeomState& g0,
                                          const BinState& b, ImgState& img0, const float* background,
                                          float* out_color, float* out_depth, cudaStream_t stream,
                                          const float* extra = nullptr, float* out_extra = nullptr);
cudaError_t launch_render_backward_batched(const ViewParams& vp, const ViewBatch& vb, const GeomState& g0,
                                           const BinState& b, const ImgState& img0, const float* background,
                                           const float* dL_dpix, float* acc, size_t acc_stride_floats,
                                           bool black_background, cudaStream_t stream);
size_t carve_binning_batched(char* base, uint32_t R_total, int V, int T, BinState* st);
size_t binning_batched_workspace_bytes(uint32_t R_total, int V, int T);
cudaError_t launch_l1_loss_grad(const float* image, const float* target, size_t n, float scale,
                                float* grad, float* loss_accum, cudaStream_t stream);
cudaError_t launch_update_stats(int P, const int* radii_max, const float* m2d_grad, int* max_radii2D,
                                float* xyz_gradient_accum, float* denom, cudaStream_t stream);
cudaError_t launch_apply_weights_render(const ViewParams& vp, const GeomState& g, const BinState& b,
                                        const ImgState& img, float* weights, int* cnt,
                                        const float* image_weights, int num_channels,
                                        cudaStream_t stream);
cudaError_t launch_apply_weights_render_batched(const ViewParams& vp, const ViewBatch* vb, const GeomState& g,
                                                const BinState& b, const ImgState& img, float* weights, int* cnt,
                                                const float* image_weights, int num_channels,
                                                cudaStream_t stream);
cudaError_t launch_debug_keys(const GeomState& g, const BinState& b, int R, uint64_t* keys_out,
                              cudaStream_t stream);
cudaError_t launch_fused_adam(float* param, const float* grad, float* m, float* v, size_t n, float lr,
                              float beta1, float beta2, float eps, int step, const uint8_t* mask,
                              int stride, cudaStream_t stream);

// per-Gaussian accumulator slots written by the backward blend: moments of q = dL/dG * G over the
// Gaussian's pixels with d = centre - pixel (sum q dx, q dy, q dx^2, q dx dy, q dy^2), dL/dopacity and
// dL/dcolour. geom_bwd.cu turns the
// moments into dL/dmean2D and dL/dconic. Rows are 12 floats (three 16-byte loads), slots 9-11 unused.
enum { ACC_SX = 0, ACC_SY, ACC_SXX, ACC_SXY, ACC_SYY, ACC_OPACITY, ACC_R, ACC_G,
       ACC_B, ACC_STRIDE = 12 };

// ---- the 64-byte blend record of a Gaussian in a view (written by preprocess) ----
//  q0: x, y (pixel-space centre), conic.x, conic.y
//  q1: conic.z, power threshold, opacity, hx
//  q2: r, g, b, view-space depth   (the one quarter that depends on the SH coefficients: dge_fit_views_colour
//      rewrites it as a whole when the projection ran before the features were final)
//  q3: hy, -cy/cz, -cy/cx, limit on the quadratic form (< 0: the exact quadrant cull does not apply)
// (hx, hy): half-extents of the box outside which alpha < 1/255 for certain (-inf = never visible).
constexpr int REC_F4 = 4;

// Lower bound on `power` below which opacity*exp(power) < 1/255 for certain.
// 0.01 of slack in the exponent is ~1% in alpha; expf and __logf err by < 1e-6.
__device__ __forceinline__ float power_threshold(float opacity) {
  return opacity > 0.0f ? -(__logf(255.0f * opacity) + 0.01f) : __int_as_float(0x7f800000);
}

// Packs the record. Exact-cull constants (blend.cuh:quad_mask): q(u,v) = 0.5 (cx u^2 + cz v^2) + cy u v
// must stay <= tau = -thr for a pixel to reach alpha >= 1/255. The limit carries a slack proportional
// to the conditioning kappa = cx cz / det of the form: the fp32 `power` of a pixel differs from the
// real-number value by < 1e-6 * kappa * q, so a rectangle is only dropped when its minimum of q exceeds
// tau by 20x that; ill-conditioned needles (kappa > 1e3) are left to the box test.
__device__ __forceinline__ void pack_record(float4* q, float x, float y, float hx, float hy, float cx,
                                            float cy, float cz, float opacity, float r, float g, float b,
                                            float depth) {
  const float thr = power_threshold(opacity);
  const float det = cx * cz - cy * cy;
  const float ac = cx * cz;
  const bool exact = cx > 0.0f && cz > 0.0f && det > 1e-3f * ac && thr < 0.0f;
  const float kappa = exact ? __fdividef(ac, det) : 1.0f;
  q[0] = make_float4(x, y, cx, cy);
  q[1] = make_float4(cz, thr, opacity, hx);
  q[2] = make_float4(r, g, b, depth);
  q[3] = make_float4(hy, exact ? __fdividef(-cy, cz) : 0.0f, exact ? __fdividef(-cy, cx) : 0.0f,
                     exact ? -thr * (1.0f + 2e-5f * kappa) + 1e-4f : -1.0f);
}

// Tile rect of a Gaussian, bit-exact with getRect (DGR/cuda_rasterizer/auxiliary.h:46-56) as
// compiled for sm_100a: two separate float adds (+16, -1), *0.0625, truncation, clamp.
__device__ __forceinline__ void tile_rect(float px, float py, int radius, int grid_x, int grid_y,
                                          int& min_x, int& min_y, int& max_x, int& max_y) {
  const float r = (float)radius;
  int a = (int)(__fmul_rn(__fadd_rn(px, -r), 0.0625f));
  int b = (int)(__fmul_rn(__fadd_rn(py, -r), 0.0625f));
  int c = (int)(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(px, r), 16.0f), -1.0f), 0.0625f));
  int d = (int)(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(py, r), 16.0f), -1.0f), 0.0625f));
  min_x = min(grid_x, max(0, a));
  min_y = min(grid_y, max(0, b));
  max_x = min(grid_x, max(0, c));
  max_y = min(grid_y, max(0, d));
}

// Lanes of the warp whose value has the same low BITS bits as this lane's, among the lanes in `peers`
// on entry (stable ranking of radix / partition passes): peers & AND(votes of my set bits) & ~OR(votes
// of my clear bits), one ballot per bit. A single match.any is one instruction but iterates over the
// DISTINCT values in the warp (measured 1.5-2x slower per pass); left to the compiler the per-bit
// select costs six instructions — spelled out it is R2P + (vote, two predicated logic ops) per bit.
template <int BITS>
__device__ __forceinline__ uint32_t same_value_lanes(uint32_t d, uint32_t peers) {
  uint32_t others = 0;
#pragma unroll
  for (int b = 0; b < BITS; b++)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 v;\n\t"
        "and.b32 v, %2, %3;\n\t"
        "setp.ne.u32 p, v, 0;\n\t"
        "vote.sync.ballot.b32 v, p, 0xffffffff;\n\t"
        "@p and.b32 %0, %0, v;\n\t"
        "@!p or.b32 %1, %1, v;\n\t}"
        : "+r"(peers), "+r"(others)
        : "r"(d), "r"(1u << b));
  return peers & ~others;
}

}  // namespace dge
