// C-ABI of libdge_b200.so (include/dge_b200.h) and the per-view orchestration that
// replaces CudaRasterizer::Rasterizer::{forward,backward,apply_weights,markVisible}
// (DGR/cuda_rasterizer/rasterizer_impl.cu:128-447).
#include <cstdio>
#include <cstring>
#include "../../include/dge_b200.h"
#include "common.cuh"

#include <mutex>
#include <vector>

namespace dge {

std::atomic<unsigned long long> g_kernel_launches{0};

static thread_local char g_err[512] = "";

// ---- optional per-stage CUDA-event timing on the launching stream (bench.py roofline) ----
enum { ST_PREPROCESS = 0, ST_DEPTH_SORT, ST_BINNING, ST_RENDER_FWD, ST_RENDER_BWD, ST_GEOM_BWD,
       ST_APPLY_WEIGHTS, ST_COUNT };
static unsigned g_profile_mask = 0;
struct EvPair { cudaEvent_t a, b; };
static std::mutex g_ev_mutex;  // the event lists are shared by every thread that calls into the library
static std::vector<EvPair> g_ev_free;
static std::vector<EvPair> g_ev_used[ST_COUNT];
struct StageScope {
  int st; cudaStream_t s; bool on; EvPair ev;
  StageScope(int st_, cudaStream_t s_) : st(st_), s(s_), on((g_profile_mask >> st_) & 1u) {
    if (!on) return;
    {
      std::lock_guard<std::mutex> lock(g_ev_mutex);
      if (g_ev_free.empty()) { cudaEventCreate(&ev.a); cudaEventCreate(&ev.b); }
      else { ev = g_ev_free.back(); g_ev_free.pop_back(); }
    }
    cudaEventRecord(ev.a, s);
  }
  ~StageScope() {
    if (!on) return;
    cudaEventRecord(ev.b, s);
    std::lock_guard<std::mutex> lock(g_ev_mutex);
    g_ev_used[st].push_back(ev);
  }
};

static int fail(const char* where, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "dge_b200: %s: %s (%d)", where, cudaGetErrorString(e), (int)e);
  return -(int)(e == cudaSuccess ? 1 : e);
}
static int fail_msg(const char* msg) {
  snprintf(g_err, sizeof(g_err), "dge_b200: %s", msg);
  return -1;
}

#define CK(where, expr)                               \
  do {                                                \
    cudaError_t _e = (expr);                          \
    if (_e != cudaSuccess) return fail(where, _e);    \
  } while (0)
// debug=True in the reference = synchronise and check after every stage (auxiliary.h:166-173)
#define STAGE(st, where, expr)                                               \
  do {                                                                       \
    {                                                                        \
      StageScope _scope(st, stream);                                         \
      CK(where, (expr));                                                     \
    }                                                                        \
    if (debug) CK(where " (debug sync)", cudaStreamSynchronize(stream));     \
  } while (0)

// ------------------------------------------------------------------- carving -----
template <typename T>
static void take(char*& p, T*& out, size_t count) {
  uintptr_t a = (reinterpret_cast<uintptr_t>(p) + 255) & ~uintptr_t(255);
  out = reinterpret_cast<T*>(a);
  p = reinterpret_cast<char*>(out + count);
}

size_t carve_geom(char* base, int P, GeomState* st) {
  GeomState s;
  char* p = base;
  const size_t n = (size_t)P;
  take(p, s.rec, n * REC_F4);
  take(p, s.rect, n);
  take(p, s.clamped, n);
  take(p, s.sort_key[0], n);
  take(p, s.sort_key[1], n);
  take(p, s.sort_val[0], n);
  take(p, s.sort_val[1], n);
  take(p, s.offsets, n);
  take(p, s.block_sums, (n + 255) / 256 + 1);
  take(p, s.counters, 256);  // [0] num_rendered; view 0 of a batch: [64 .. 64+V] = seg_off
  s.sort_ws_bytes = sort_workspace_bytes((uint32_t)P);
  take(p, s.sort_ws, s.sort_ws_bytes / sizeof(uint32_t));
  if (st) *st = s;
  return (size_t)(p - base) + 256;
}

size_t carve_binning(char* base, int R, int width, int height, BinState* st) {
  (void)width;
  (void)height;
  BinState s;
  char* p = base;
  const size_t n = (size_t)(R > 0 ? R : 1);
  take(p, s.point_list, n);
  take(p, s.tile_ids, n);
  take(p, s.key_alt, n);
  take(p, s.val_alt, n);
  s.sort_ws_bytes = sort_workspace_bytes((uint32_t)n);
  take(p, s.sort_ws, s.sort_ws_bytes / sizeof(uint32_t));
  if (st) *st = s;
  return (size_t)(p - base) + 256;
}

// Instance lists of all V views of a fit step, back to back (view v at seg_off[v]).
size_t carve_binning_batched(char* base, uint32_t R_total, int V, int T, BinState* st) {
  BinState s;
  char* p = base;
  const size_t n = (size_t)(R_total > 0 ? R_total : 1);
  take(p, s.point_list, n);
  take(p, s.tile_ids, n);
  take(p, s.key_alt, n);
  take(p, s.val_alt, n);
  s.sort_ws_bytes = binning_batched_workspace_bytes((uint32_t)n, V, T);
  take(p, s.sort_ws, s.sort_ws_bytes / sizeof(uint32_t));
  if (st) *st = s;
  return (size_t)(p - base) + 256;
}

size_t carve_image(char* base, int width, int height, ImgState* st) {
  ImgState s;
  char* p = base;
  const size_t N = (size_t)width * height;
  const size_t T = (size_t)((width + DGE_TILE - 1) / DGE_TILE) * ((height + DGE_TILE - 1) / DGE_TILE);
  take(p, s.final_T, N);
  take(p, s.n_contrib, N);
  take(p, s.ranges, T);
  if (st) *st = s;
  return (size_t)(p - base) + 256;
}

// Pinned word + event used to bring num_rendered to the host without draining the stream:
// the wait covers preprocess only; the depth sort queued behind it keeps the GPU busy.
struct HostSlot {
  uint32_t* pinned = nullptr;
  cudaEvent_t ev = nullptr;
};
static thread_local HostSlot g_slot;

static cudaError_t ensure_slot() {
  if (g_slot.pinned) return cudaSuccess;
  cudaError_t e = cudaHostAlloc((void**)&g_slot.pinned, 1024, cudaHostAllocDefault);
  if (e != cudaSuccess) return e;
  return cudaEventCreateWithFlags(&g_slot.ev, cudaEventDisableTiming);
}

static ViewParams make_view(int P, int D, int M, int width, int height, const float* viewmatrix,
                            const float* projmatrix, const float* cam_pos, float tan_fovx,
                            float tan_fovy, float scale_modifier) {
  ViewParams vp;
  vp.view = viewmatrix;
  vp.proj = projmatrix;
  vp.campos = cam_pos;
  vp.tan_fovx = tan_fovx;
  vp.tan_fovy = tan_fovy;
  // rasterizer_impl.cu:190-191, host float arithmetic
  vp.focal_y = height / (2.0f * tan_fovy);
  vp.focal_x = width / (2.0f * tan_fovx);
  vp.scale_modifier = scale_modifier;
  vp.W = width;
  vp.H = height;
  vp.grid_x = (width + DGE_TILE - 1) / DGE_TILE;
  vp.grid_y = (height + DGE_TILE - 1) / DGE_TILE;
  vp.P = P;
  vp.D = D;
  vp.M = M;
  return vp;
}

// Shared front half of forward and apply_weights: K1(/K13) .. K5.
static int bin_view(dge_alloc_fn geometryBuffer, dge_alloc_fn binningBuffer, dge_alloc_fn imageBuffer,
                    void* ctx, const ViewParams& vp, const float* means3D, const float* shs,
                    const float* colors_precomp, int colors_mode, const float* opacities,
                    const float* scales, const float* rotations, const float* cov3D_precomp,
                    bool prefiltered, int* radii, bool debug, cudaStream_t stream, GeomState& g,
                    BinState& b, ImgState& img, uint8_t* flags_out = nullptr) {
  if (vp.grid_x > 65535 || vp.grid_y > 65535) return fail_msg("image too large (tile grid > 65535)");
  if (cov3D_precomp == nullptr && (scales == nullptr || rotations == nullptr))
    return fail_msg("need scales+rotations or cov3D_precomp");
  CK("pinned slot", ensure_slot());
  char* gp = geometryBuffer(ctx, carve_geom(nullptr, vp.P, nullptr));
  char* ip = imageBuffer(ctx, carve_image(nullptr, vp.W, vp.H, nullptr));
  if (!gp || !ip) return fail_msg("scratch allocator returned NULL");
  carve_geom(gp, vp.P, &g);
  carve_image(ip, vp.W, vp.H, &img);
  STAGE(ST_PREPROCESS, "preprocess", launch_preprocess(vp, means3D, scales, rotations, opacities, shs, cov3D_precomp,
                                        colors_precomp, colors_mode, prefiltered, radii, g, flags_out, stream));
  CK("num_rendered copy", cudaMemcpyAsync(g_slot.pinned, g.counters, sizeof(uint32_t),
                                          cudaMemcpyDeviceToHost, stream));
  CK("event record", cudaEventRecord(g_slot.ev, stream));
  STAGE(ST_DEPTH_SORT, "depth sort", launch_depth_sort(vp.P, g, stream));
  CK("num_rendered wait", cudaEventSynchronize(g_slot.ev));
  const uint32_t R = *g_slot.pinned;
  if (R >= (1u << 30)) return fail_msg("num_rendered exceeds 2^30 instances");
  char* bp = binningBuffer(ctx, carve_binning(nullptr, (int)R, vp.W, vp.H, nullptr));
  if (!bp) return fail_msg("scratch allocator returned NULL");
  carve_binning(bp, (int)R, vp.W, vp.H, &b);
  STAGE(ST_BINNING, "binning", launch_binning(vp, (int)R, g, b, img, stream));
  return (int)R;
}

// Zeroing of a step's blend-stage sums beside the forward blend: forked from `stream` onto a side stream;
// the backward blend of the same `acc` joins it (wait_acc_zero). One slot per in-flight batch (the chunks of
// a step each have their own rows).
// (Process-wide, not per thread: the backward half of a step may run on autograd's worker thread.)
struct AccZero { const void* acc; cudaEvent_t done; };
static std::mutex g_zero_mutex;
static cudaStream_t g_zero_stream = nullptr;
static cudaEvent_t g_zero_fork = nullptr;
static AccZero g_zero[16];

static cudaError_t zero_acc_async(float* acc, size_t acc_stride_floats, int P, int V, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_zero_mutex);
  cudaError_t e;
  if (g_zero_stream == nullptr) {
    // (default priority: at the highest one the memset's blocks displace the forward blend's from the start and
    // that kernel takes 0.10 ms longer; as it is, the memset fills the slots the blend's tail leaves free. A zeroing
    // kernel of our own instead of cudaMemsetAsync is dispatched ahead of the blend and simply delays it: +0.07 ms)
    if ((e = cudaStreamCreateWithFlags(&g_zero_stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&g_zero_fork, cudaEventDisableTiming)) != cudaSuccess) return e;
    for (AccZero& z : g_zero) {
      z.acc = nullptr;
      if ((e = cudaEventCreateWithFlags(&z.done, cudaEventDisableTiming)) != cudaSuccess) return e;
    }
  }
  AccZero* slot = nullptr;
  for (AccZero& z : g_zero)
    if (z.acc == acc || (slot == nullptr && z.acc == nullptr)) { slot = &z; if (z.acc == acc) break; }
  if (slot == nullptr) {  // more batches in flight than slots: zero in line
    return cudaMemset2DAsync(acc, acc_stride_floats * sizeof(float), 0, sizeof(float) * ACC_STRIDE * (size_t)P, V, stream);
  }
  if ((e = cudaEventRecord(g_zero_fork, stream)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(g_zero_stream, g_zero_fork, 0)) != cudaSuccess) return e;
  if (acc_stride_floats == (size_t)ACC_STRIDE * P)
    e = cudaMemsetAsync(acc, 0, sizeof(float) * acc_stride_floats * (size_t)V, g_zero_stream);
  else
    e = cudaMemset2DAsync(acc, acc_stride_floats * sizeof(float), 0, sizeof(float) * ACC_STRIDE * (size_t)P, V, g_zero_stream);
  if (e != cudaSuccess) return e;
  slot->acc = acc;
  return cudaEventRecord(slot->done, g_zero_stream);
}

static cudaError_t wait_acc_zero(const float* acc, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_zero_mutex);
  for (AccZero& z : g_zero)
    if (z.acc == acc && acc != nullptr) {
      z.acc = nullptr;
      return cudaStreamWaitEvent(stream, z.done, 0);
    }
  return cudaSuccess;  // the caller zeroed the rows itself
}

// Measurement hook: every SM that gets a block stores its cycle counter and the global nanosecond timer.
// Two probes on the same stream, one before and one after a timed region, give the AVERAGE SM clock the
// region ran at (cycles / ns per SM) without a single driver query in between (an NVML / nvidia-smi
// poll stalls the launching thread for ~25 ms on this driver).
__global__ void clock_probe_kernel(unsigned long long* out) {
  if (threadIdx.x == 0) {
    unsigned smid;
    unsigned long long t;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    out[2 * smid] = (unsigned long long)clock64();
    out[2 * smid + 1] = t;
  }
}

}  // namespace dge

using namespace dge;

extern "C" {

const char* dge_last_error(void) { return g_err; }
int dge_abi_version(void) { return 14; }
unsigned long long dge_launch_count(void) { return g_kernel_launches.load(std::memory_order_relaxed); }

int dge_clock_probe(unsigned long long* out, void* stream_) {
  clock_probe_kernel<<<DGE_NUM_SMS * 8, 32, 0, (cudaStream_t)stream_>>>(out);
  CK("clock probe", cudaGetLastError());
  return 0;
}

void dge_profile_enable(unsigned stage_mask) { g_profile_mask = stage_mask; }
// Sums (and clears) the event-timed durations recorded since the last call.
// ms_out / count_out have DGE_NUM_STAGES entries. Synchronises on the recorded events.
int dge_profile_read(float* ms_out, int* count_out) {
  std::vector<EvPair> used[ST_COUNT];
  {
    std::lock_guard<std::mutex> lock(g_ev_mutex);
    for (int st = 0; st < ST_COUNT; st++) used[st].swap(g_ev_used[st]);
  }
  for (int st = 0; st < ST_COUNT; st++) {
    float total = 0.f;
    for (const EvPair& ev : used[st]) {
      cudaError_t e = cudaEventSynchronize(ev.b);
      float ms = 0.f;
      if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, ev.a, ev.b);
      if (e != cudaSuccess) return fail("profile read", e);
      total += ms;
    }
    ms_out[st] = total;
    count_out[st] = (int)used[st].size();
  }
  std::lock_guard<std::mutex> lock(g_ev_mutex);
  for (int st = 0; st < ST_COUNT; st++) g_ev_free.insert(g_ev_free.end(), used[st].begin(), used[st].end());
  return 0;
}

size_t dge_geom_bytes(int P) { return carve_geom(nullptr, P, nullptr); }
size_t dge_binning_bytes(int R, int width, int height) {
  return carve_binning(nullptr, R, width, height, nullptr);
}
size_t dge_image_bytes(int width, int height) { return carve_image(nullptr, width, height, nullptr); }
size_t dge_backward_scratch_bytes(int P) { return sizeof(float) * ACC_STRIDE * (size_t)P + 256; }

int dge_rasterize_forward(dge_alloc_fn geometryBuffer, dge_alloc_fn binningBuffer,
                          dge_alloc_fn imageBuffer, void* alloc_ctx, int P, int D, int M,
                          const float* background, int width, int height, const float* means3D,
                          const float* shs, const float* colors_precomp, const float* opacities,
                          const float* scales, float scale_modifier, const float* rotations,
                          const float* cov3D_precomp, const float* viewmatrix,
                          const float* projmatrix, const float* cam_pos, float tan_fovx,
                          float tan_fovy, int prefiltered, float* out_color, float* out_depth,
                          int* radii, int debug, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (width <= 0 || height <= 0) return fail_msg("empty image");
  const size_t N = (size_t)width * height;
  if (P == 0) {  // rasterize_points.cu:72: kernels skipped, outputs stay zero
    CK("memset", cudaMemsetAsync(out_color, 0, 3 * N * sizeof(float), stream));
    CK("memset", cudaMemsetAsync(out_depth, 0, N * sizeof(float), stream));
    return 0;
  }
  if (shs == nullptr && colors_precomp == nullptr) return fail_msg("need shs or colors_precomp");
  const ViewParams vp = make_view(P, D, M, width, height, viewmatrix, projmatrix, cam_pos, tan_fovx,
                                  tan_fovy, scale_modifier);
  GeomState g;
  BinState b;
  ImgState img;
  const int R = bin_view(geometryBuffer, binningBuffer, imageBuffer, alloc_ctx, vp, means3D, shs,
                         colors_precomp, colors_precomp ? 1 : 0, opacities, scales, rotations,
                         cov3D_precomp, prefiltered != 0, radii, debug != 0, stream, g, b, img);
  if (R < 0) return R;
  STAGE(ST_RENDER_FWD, "render forward",
        launch_render_forward(vp, g, b, img, background, out_color, out_depth, stream));
  return R;
}

// ---- fit-step entry points (include/dge_b200.h "fit step") ----
int dge_fit_forward(dge_alloc_fn geometryBuffer, dge_alloc_fn binningBuffer, dge_alloc_fn imageBuffer,
                        void* alloc_ctx, int P, int D, int M, const float* background, int width,
                        int height, const float* means3D, const float* shs, const float* opacities,
                        const float* scales, float scale_modifier, const float* rotations,
                        const float* cam, float tan_fovx, float tan_fovy, float* out_color,
                        float* out_depth, int* radii, float* acc, uint8_t* flags, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool debug = false;
  if (P == 0 || width <= 0 || height <= 0) return fail_msg("empty problem");
  if (acc != nullptr) CK("acc memset", cudaMemsetAsync(acc, 0, sizeof(float) * ACC_STRIDE * (size_t)P, stream));
  const ViewParams vp = make_view(P, D, M, width, height, cam, cam + 16, cam + 32, tan_fovx, tan_fovy,
                                  scale_modifier);
  GeomState g;
  BinState b;
  ImgState img;
  const int R = bin_view(geometryBuffer, binningBuffer, imageBuffer, alloc_ctx, vp, means3D, shs, nullptr, 0,
                         opacities, scales, rotations, nullptr, false, radii, debug, stream, g, b, img, flags);
  if (R < 0) return R;
  STAGE(ST_RENDER_FWD, "render forward",
        launch_render_forward(vp, g, b, img, background, out_color, out_depth, stream));
  return R;
}

// ---- all V views of a step, one launch per stage ----
#define DGE_MAX_BATCH_VIEWS 64
static uint32_t* batch_seg_off(const GeomState& g0) { return g0.counters + 64; }
// per-view blobs of a batch: the single-view layout repeated at a 256-byte-aligned stride
static size_t batch_stride(size_t bytes) { return (bytes + 255) & ~size_t(255); }

size_t dge_fit_binning_bytes(int R_total, int V, int width, int height) {
  const int T = ((width + DGE_TILE - 1) / DGE_TILE) * ((height + DGE_TILE - 1) / DGE_TILE);
  return carve_binning_batched(nullptr, (uint32_t)(R_total > 0 ? R_total : 0), V, T, nullptr);
}

// Shared front half of the per-step entry points: batched preprocess .. binning for V views.
// shs == nullptr: no colour (mask back-projection). Returns R_total (>= 0) or < 0.
static int bin_views(dge_alloc_fn geometryBuffer, dge_alloc_fn binningBuffer, dge_alloc_fn imageBuffer,
                     void* alloc_ctx, const ViewParams& vp, int V, const float* means3D, const float* shs,
                     const float* opacities, const float* scales, const float* rotations, const float* cams,
                     int* radii_max, uint8_t* flags, size_t flags_stride, int* num_rendered_host,
                     bool prune_lists, cudaStream_t stream, GeomState& g0, BinState& b, ImgState& img0,
                     ViewBatch& vb) {
  const bool debug = false;
  if (V < 1 || V > DGE_MAX_BATCH_VIEWS) return fail_msg("a batch holds 1..64 views");
  if (vp.grid_x > 65535 || vp.grid_y > 65535) return fail_msg("image too large (tile grid > 65535)");
  CK("pinned slot", ensure_slot());
  const size_t gstride = batch_stride(carve_geom(nullptr, vp.P, nullptr));
  const size_t istride = batch_stride(carve_image(nullptr, vp.W, vp.H, nullptr));
  char* gp = geometryBuffer(alloc_ctx, gstride * V);
  char* ip = imageBuffer(alloc_ctx, istride * V);
  if (!gp || !ip) return fail_msg("scratch allocator returned NULL");
  carve_geom(gp, vp.P, &g0);
  carve_image(ip, vp.W, vp.H, &img0);
  vb.V = V;
  vb.geom_stride = gstride;
  vb.img_stride = istride;
  vb.seg_off = batch_seg_off(g0);
  vb.cams = cams;
  STAGE(ST_PREPROCESS, "preprocess (batched)",
        launch_preprocess_batched(vp, vb, means3D, scales, rotations, opacities, shs, g0, flags, flags_stride,
                                  radii_max, prune_lists, stream));
  CK("segment offsets", launch_seg_offsets(vb, g0, batch_seg_off(g0), stream));
  CK("num_rendered copy", cudaMemcpyAsync(g_slot.pinned, batch_seg_off(g0), sizeof(uint32_t) * (V + 1),
                                          cudaMemcpyDeviceToHost, stream));
  CK("event record", cudaEventRecord(g_slot.ev, stream));
  STAGE(ST_DEPTH_SORT, "depth sort (batched)", launch_depth_sort_batched(vp.P, vb, g0, stream));
  CK("num_rendered wait", cudaEventSynchronize(g_slot.ev));
  const uint32_t R_total = g_slot.pinned[V];
  uint32_t R_max = 0;
  for (int v = 0; v < V; v++) {
    const uint32_t r = g_slot.pinned[v + 1] - g_slot.pinned[v];
    if (r > R_max) R_max = r;
    if (num_rendered_host) num_rendered_host[v] = (int)r;
  }
  if (R_total >= (1u << 30)) return fail_msg("the step's views exceed 2^30 instances: use smaller batches");
  const int T = vp.grid_x * vp.grid_y;
  char* bp = binningBuffer(alloc_ctx, carve_binning_batched(nullptr, R_total, V, T, nullptr));
  if (!bp) return fail_msg("scratch allocator returned NULL");
  carve_binning_batched(bp, R_total, V, T, &b);
  STAGE(ST_BINNING, "binning (batched)", launch_binning_batched(vp, vb, R_total, R_max, g0, b, img0, stream));
  return (int)R_total;
}

int dge_fit_views_front(dge_alloc_fn geometryBuffer, dge_alloc_fn binningBuffer, dge_alloc_fn imageBuffer,
                        void* alloc_ctx, int P, int D, int M, int V, int width, int height, const float* means3D,
                        const float* shs, const float* opacities, const float* scales, float scale_modifier,
                        const float* rotations, const float* cams, int* radii_max, uint8_t* flags,
                        size_t flags_stride, int* num_rendered_host, int prune_lists, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (P == 0 || width <= 0 || height <= 0) return fail_msg("empty problem");
  if (flags != nullptr && flags_stride < (size_t)P) return fail_msg("flags_stride >= P is required");
  if (shs != nullptr && M != 16) return fail_msg("the batched fit path needs SH degree-3 storage (M == 16)");
  // tan_fov / focal are per view and filled in by the batched preprocess from the camera records
  const ViewParams vp = make_view(P, D, M, width, height, nullptr, nullptr, nullptr, 1.f, 1.f, scale_modifier);
  GeomState g0;
  ImgState img0;
  BinState b;
  ViewBatch vb;
  return bin_views(geometryBuffer, binningBuffer, imageBuffer, alloc_ctx, vp, V, means3D, shs, opacities, scales,
                   rotations, cams, radii_max, flags, flags_stride, num_rendered_host, prune_lists != 0, stream, g0,
                   b, img0, vb);
}

int dge_fit_views_colour(int P, int D, int M, int V, const float* means3D, const float* shs, const float* cams,
                         char* geom_buffer, uint8_t* flags, size_t flags_stride, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (P == 0) return 0;
  if (M != 16 || shs == nullptr || flags == nullptr) return fail_msg("dge_fit_views_colour needs SH degree-3 storage and the flags of the front half");
  if (V < 1 || V > DGE_MAX_BATCH_VIEWS) return fail_msg("a batch holds 1..64 views");
  GeomState g0;
  const size_t gstride = batch_stride(carve_geom(geom_buffer, P, &g0));
  CK("colour (batched)", launch_colour_batched(P, D, V, cams, means3D, shs, g0, gstride, flags, flags_stride, stream));
  return 0;
}

int dge_fit_views_blend(int P, int V, int R_total, const float* background, int width, int height, char* geom_buffer,
                        char* binning_buffer, char* image_buffer, float* out_color, float* out_depth, float* acc,
                        size_t acc_stride_floats, const float* extra, float* out_extra, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool debug = false;
  if (P == 0 || width <= 0 || height <= 0) return fail_msg("empty problem");
  if (V < 1 || V > DGE_MAX_BATCH_VIEWS) return fail_msg("a batch holds 1..64 views");
  if (acc != nullptr && (acc_stride_floats < (size_t)ACC_STRIDE * P || (acc_stride_floats & 3)))
    return fail_msg("acc_stride_floats >= 12 P (a multiple of 4) is required");
  if ((extra == nullptr) != (out_extra == nullptr)) return fail_msg("extra and out_extra go together");
  const ViewParams vp = make_view(P, 0, 0, width, height, nullptr, nullptr, nullptr, 1.f, 1.f, 1.f);
  GeomState g0;
  BinState b;
  ImgState img0;
  ViewBatch vb;
  vb.V = V;
  vb.geom_stride = batch_stride(carve_geom(geom_buffer, P, &g0));
  vb.img_stride = batch_stride(carve_image(image_buffer, width, height, &img0));
  carve_binning_batched(binning_buffer, (uint32_t)R_total, V, vp.grid_x * vp.grid_y, &b);
  vb.seg_off = batch_seg_off(g0);
  vb.cams = nullptr;
  // the views' rows of blend-stage sums start at zero: a memset on a side stream, forked here so that it runs
  // beside the forward blend (issue-bound, 3 % of the DRAM bandwidth) instead of inside preprocess
  if (acc != nullptr) CK("acc zero", zero_acc_async(acc, acc_stride_floats, P, V, stream));
  STAGE(ST_RENDER_FWD, "render forward (batched)",
        launch_render_forward_batched(vp, vb, g0, b, img0, background, out_color, out_depth, stream, extra,
                                      out_extra));
  return 0;
}

int dge_fit_views_forward(dge_alloc_fn geometryBuffer, dge_alloc_fn binningBuffer, dge_alloc_fn imageBuffer,
                          void* alloc_ctx, int P, int D, int M, int V, const float* background, int width,
                          int height, const float* means3D, const float* shs, const float* opacities,
                          const float* scales, float scale_modifier, const float* rotations,
                          const float* cams, float* out_color, float* out_depth, int* radii_max, float* acc,
                          size_t acc_stride_floats, uint8_t* flags, size_t flags_stride, int* num_rendered_host,
                          const float* extra, float* out_extra, int prune_lists, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool debug = false;
  if (P == 0 || width <= 0 || height <= 0) return fail_msg("empty problem");
  if ((acc == nullptr) != (flags == nullptr)) return fail_msg("acc and flags go together");
  if (acc != nullptr && (acc_stride_floats < (size_t)ACC_STRIDE * P || (acc_stride_floats & 3) || flags_stride < (size_t)P))
    return fail_msg("acc_stride_floats >= 12 P (a multiple of 4) and flags_stride >= P are required");
  if ((extra == nullptr) != (out_extra == nullptr)) return fail_msg("extra and out_extra go together");
  if (M != 16 || shs == nullptr) return fail_msg("the batched fit path needs SH degree-3 storage (M == 16)");
  // tan_fov / focal are per view and filled in by the batched preprocess from the camera records
  const ViewParams vp = make_view(P, D, M, width, height, nullptr, nullptr, nullptr, 1.f, 1.f, scale_modifier);
  GeomState g0;
  ImgState img0;
  BinState b;
  ViewBatch vb;
  const int R_total = bin_views(geometryBuffer, binningBuffer, imageBuffer, alloc_ctx, vp, V, means3D, shs, opacities,
                                scales, rotations, cams, radii_max, flags, flags_stride, num_rendered_host,
                                prune_lists != 0, stream, g0, b, img0, vb);
  if (R_total < 0) return R_total;
  if (acc != nullptr) CK("acc zero", zero_acc_async(acc, acc_stride_floats, P, V, stream));
  STAGE(ST_RENDER_FWD, "render forward (batched)",
        launch_render_forward_batched(vp, vb, g0, b, img0, background, out_color, out_depth, stream, extra,
                                      out_extra));
  return R_total;
}

int dge_fit_views_apply_weights(dge_alloc_fn geometryBuffer, dge_alloc_fn binningBuffer, dge_alloc_fn imageBuffer,
                                void* alloc_ctx, int P, int V, int width, int height, const float* means3D,
                                const float* opacities, const float* scales, float scale_modifier,
                                const float* rotations, const float* cams, const float* image_weights,
                                int num_channels, float* weights, int* cnt, int* num_rendered_host,
                                int prune_lists, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool debug = false;
  if (P == 0) return 0;
  if (width <= 0 || height <= 0) return fail_msg("empty image");
  if (num_channels < 1 || num_channels > 3) return fail_msg("Unsupported number of channels");
  const ViewParams vp = make_view(P, 0, 0, width, height, nullptr, nullptr, nullptr, 1.f, 1.f, scale_modifier);
  GeomState g0;
  ImgState img0;
  BinState b;
  ViewBatch vb;
  const int R_total = bin_views(geometryBuffer, binningBuffer, imageBuffer, alloc_ctx, vp, V, means3D, nullptr,
                                opacities, scales, rotations, cams, nullptr, nullptr, 0, num_rendered_host,
                                prune_lists != 0, stream, g0, b, img0, vb);
  if (R_total <= 0) return R_total;
  STAGE(ST_APPLY_WEIGHTS, "apply_weights blend (batched)",
        launch_apply_weights_render_batched(vp, &vb, g0, b, img0, weights, cnt, image_weights, num_channels, stream));
  return R_total;
}

int dge_fit_views_backward_blend(int P, int V, int R_total, const float* background, int background_is_black,
                                 int width, int height, char* geom_buffer, char* binning_buffer,
                                 char* image_buffer, const float* dL_dpix, float* acc, size_t acc_stride_floats,
                                 void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool debug = false;
  if (P == 0 || R_total == 0) return 0;
  if (V < 1 || V > DGE_MAX_BATCH_VIEWS) return fail_msg("a batch holds 1..64 views");
  const ViewParams vp = make_view(P, 0, 0, width, height, nullptr, nullptr, nullptr, 1.f, 1.f, 1.f);
  GeomState g0;
  BinState b;
  ImgState img0;
  ViewBatch vb;
  vb.V = V;
  vb.geom_stride = batch_stride(carve_geom(geom_buffer, P, &g0));
  vb.img_stride = batch_stride(carve_image(image_buffer, width, height, &img0));
  carve_binning_batched(binning_buffer, (uint32_t)R_total, V, vp.grid_x * vp.grid_y, &b);
  vb.seg_off = batch_seg_off(g0);
  vb.cams = nullptr;
  CK("acc zero wait", wait_acc_zero(acc, stream));
  STAGE(ST_RENDER_BWD, "render backward (batched)",
        launch_render_backward_batched(vp, vb, g0, b, img0, background, dL_dpix, acc, acc_stride_floats,
                                       background_is_black != 0, stream));
  return 0;
}

int dge_fit_backward_blend(int P, int R, const float* background, int background_is_black, int width,
                           int height, char* geom_buffer, char* binning_buffer, char* image_buffer,
                           const float* dL_dpix, float* acc, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool debug = false;
  if (P == 0 || R == 0) return 0;
  const ViewParams vp = make_view(P, 0, 0, width, height, nullptr, nullptr, nullptr, 1.f, 1.f, 1.f);
  GeomState g;
  BinState b;
  ImgState img;
  carve_geom(geom_buffer, P, &g);
  carve_binning(binning_buffer, R, width, height, &b);
  carve_image(image_buffer, width, height, &img);
  STAGE(ST_RENDER_BWD, "render backward",
        launch_render_backward(vp, g, b, img, background, dL_dpix, acc, background_is_black != 0, stream));
  return 0;
}

int dge_fit_backward_geom(int P, int D, int M, int V, const float* cams, int width, int height,
                          float scale_modifier, const float* acc, size_t acc_stride_floats,
                          const uint8_t* flags, size_t flags_stride, const float* means3D, const float* shs, const float* scales, const float* rotations,
                          float* dL_dmean3D, float* dL_dmean2D, float* dL_dsh, float* dL_dopacity,
                          float* dL_dscale, float* dL_drot, int accumulate, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool debug = false;
  if (P == 0) return 0;
  STAGE(ST_GEOM_BWD, "batched geometry backward",
        launch_geom_backward_batched(P, D, M, V, cams, width, height, scale_modifier, acc, acc_stride_floats,
                                     flags, flags_stride, means3D, shs, scales, rotations, dL_dmean3D, dL_dmean2D, dL_dsh,
                                     dL_dopacity, dL_dscale, dL_drot, accumulate != 0, stream));
  return 0;
}

int dge_fit_activate(int P, const float* f_dc, const float* f_rest, const float* opacity_raw,
                     const float* scaling_raw, const float* rotation_raw, float* shs, float* opacities,
                     float* scales, float* rotations, void* stream_) {
  if (P == 0) return 0;
  CK("activate", launch_activate(P, f_dc, f_rest, opacity_raw, scaling_raw, rotation_raw, shs, opacities, scales,
                                 rotations, (cudaStream_t)stream_));
  return 0;
}

int dge_fit_backward_geom_raw(int P, int D, int V, const float* cams, int width, int height,
                              float scale_modifier, const float* acc, size_t acc_stride_floats,
                              const uint8_t* flags, size_t flags_stride, const float* means3D, const float* shs, const float* opacities,
                              const float* scales, const float* rotations, const float* rotation_raw,
                              float* d_xyz, float* d_means2D, float* d_f_dc, float* d_f_rest,
                              float* d_opacity_raw, float* d_scaling_raw, float* d_rotation_raw,
                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool debug = false;
  if (P == 0) return 0;
  if (rotation_raw == nullptr) return fail_msg("rotation_raw is required");
  STAGE(ST_GEOM_BWD, "batched geometry backward (raw)",
        launch_geom_backward_batched(P, D, 16, V, cams, width, height, scale_modifier, acc, acc_stride_floats,
                                     flags, flags_stride, means3D, shs, scales, rotations, d_xyz, d_means2D, d_f_dc, d_opacity_raw,
                                     d_scaling_raw, d_rotation_raw, false, stream, opacities, rotation_raw,
                                     d_f_rest));
  return 0;
}

int dge_rasterize_backward(dge_alloc_fn scratchBuffer, void* alloc_ctx, int P, int D, int M, int R,
                           const float* background, int width, int height, const float* means3D,
                           const float* shs, const float* colors_precomp, const float* scales,
                           float scale_modifier, const float* rotations, const float* cov3D_precomp,
                           const float* viewmatrix, const float* projmatrix, const float* campos,
                           float tan_fovx, float tan_fovy, const int* radii, char* geom_buffer,
                           char* binning_buffer, char* image_buffer, const float* dL_dpix,
                           float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                           float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                           float* dL_drot, int accumulate, int debug, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (P == 0) return 0;
  if (!dL_dmean2D || !dL_dmean3D) return fail_msg("dL_dmean2D and dL_dmean3D are required");
  const ViewParams vp = make_view(P, D, M, width, height, viewmatrix, projmatrix, campos, tan_fovx,
                                  tan_fovy, scale_modifier);
  GeomState g;
  BinState b;
  ImgState img;
  carve_geom(geom_buffer, P, &g);
  carve_binning(binning_buffer, R, width, height, &b);
  carve_image(image_buffer, width, height, &img);
  const size_t acc_bytes = sizeof(float) * ACC_STRIDE * (size_t)P;
  char* sp = scratchBuffer(alloc_ctx, acc_bytes + 256);
  if (!sp) return fail_msg("scratch allocator returned NULL");
  float* acc = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sp) + 255) & ~uintptr_t(255));
  CK("memset acc", cudaMemsetAsync(acc, 0, acc_bytes, stream));
  if (R > 0)
    STAGE(ST_RENDER_BWD, "render backward",
          launch_render_backward(vp, g, b, img, background, dL_dpix, acc, /*black_background=*/false, stream));
  (void)colors_precomp;
  STAGE(ST_GEOM_BWD, "geometry backward",
        launch_geom_backward(vp, means3D, scales, rotations, shs, cov3D_precomp, radii, g, acc,
                             dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D,
                             dL_dsh, dL_dscale, dL_drot, accumulate != 0, stream));
  return 0;
}

int dge_apply_weights(dge_alloc_fn geometryBuffer, dge_alloc_fn binningBuffer,
                      dge_alloc_fn imageBuffer, void* alloc_ctx, int P, int D, int M,
                      const float* background, int width, int height, const float* means3D,
                      const float* shs, float* weights, const float* opacities, const float* scales,
                      float scale_modifier, const float* rotations, const float* cov3D_precomp,
                      const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                      float tan_fovx, float tan_fovy, int prefiltered, const float* image_weights,
                      int* radii, int* cnt, int num_channels, int debug, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  (void)background;
  (void)shs;
  if (P == 0) return 0;
  if (width <= 0 || height <= 0) return fail_msg("empty image");
  if (num_channels < 1 || num_channels > 3) return fail_msg("Unsupported number of channels");
  const ViewParams vp = make_view(P, D, M, width, height, viewmatrix, projmatrix, cam_pos, tan_fovx,
                                  tan_fovy, scale_modifier);
  GeomState g;
  BinState b;
  ImgState img;
  // rasterizer_impl.cu:381-389: `weights` travels as colors_precomp, so SH is skipped (K13)
  const int R = bin_view(geometryBuffer, binningBuffer, imageBuffer, alloc_ctx, vp, means3D, nullptr,
                         nullptr, /*colors_mode=*/2, opacities, scales, rotations, cov3D_precomp,
                         prefiltered != 0, radii, debug != 0, stream, g, b, img);
  if (R < 0) return R;
  if (R > 0)
    STAGE(ST_APPLY_WEIGHTS, "apply_weights blend", launch_apply_weights_render(vp, g, b, img, weights, cnt,
                                                             image_weights, num_channels, stream));
  return R;
}

int dge_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream_) {
  if (P == 0) return 0;
  CK("mark_visible",
     launch_mark_visible(P, means3D, viewmatrix, projmatrix, present, (cudaStream_t)stream_));
  return 0;
}

void dge_geom_pointers(char* chunk, int P, void** out) {
  GeomState g;
  carve_geom(chunk, P, &g);
  out[0] = g.rec;
  out[1] = nullptr;
  out[2] = nullptr;
  out[3] = g.rect;
  out[4] = g.clamped;
  out[5] = g.sort_val[0];
  out[6] = g.offsets;
  out[7] = g.counters;
}
void dge_binning_pointers(char* chunk, int R, int width, int height, void** out) {
  BinState b;
  carve_binning(chunk, R, width, height, &b);
  out[0] = b.point_list;
  out[1] = b.tile_ids;
}
void dge_image_pointers(char* chunk, int width, int height, void** out) {
  ImgState s;
  carve_image(chunk, width, height, &s);
  out[0] = s.final_T;
  out[1] = s.n_contrib;
  out[2] = s.ranges;
}

int dge_debug_sorted_keys(char* geom_buffer, char* binning_buffer, int P, int R, int width,
                          int height, uint64_t* keys_out, void* stream_) {
  GeomState g;
  BinState b;
  carve_geom(geom_buffer, P, &g);
  carve_binning(binning_buffer, R, width, height, &b);
  CK("debug keys", launch_debug_keys(g, b, R, keys_out, (cudaStream_t)stream_));
  return 0;
}

int dge_l1_loss_grad(const float* image, const float* target, size_t n, float scale, float* grad,
                     float* loss_accum, void* stream_) {
  CK("l1 loss", launch_l1_loss_grad(image, target, n, scale, grad, loss_accum, (cudaStream_t)stream_));
  return 0;
}

int dge_fit_update_stats(int P, const int* radii_max, const float* means2D_grad, int* max_radii2D,
                         float* xyz_gradient_accum, float* denom, void* stream_) {
  CK("update stats", launch_update_stats(P, radii_max, means2D_grad, max_radii2D, xyz_gradient_accum, denom,
                                         (cudaStream_t)stream_));
  return 0;
}

int dge_fused_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n,
                   float lr, float beta1, float beta2, float eps, int step, const uint8_t* mask,
                   int stride, void* stream_) {
  CK("fused adam", launch_fused_adam(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step,
                                     mask, stride, (cudaStream_t)stream_));
  return 0;
}

}  // extern "C"
