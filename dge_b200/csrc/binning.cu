// K2-K5: instance lists per tile, front-to-back.
//
// Replaces InclusiveSum + duplicateWithKeys + SortPairs + identifyTileRanges
// (DGR/cuda_rasterizer/rasterizer_impl.cu:67-125, :229-271). The reference
// sorts R (tile|depth) 64-bit keys in 6 byte passes (152 B/instance). The same
// stable order (tile, depth bits, Gaussian id) is produced here with far less
// traffic by sorting the P Gaussians by depth bits ONCE (4 passes over P
// items), emitting instances in that order, and stable-partitioning them by
// tile id (ceil(log2(T)/8) passes over R 8-byte pairs): 8 + 16*passes B/instance.
// Stability of both sorts makes the resulting point_list / ranges bit-identical
// to the reference's (ties in depth keep Gaussian-id order, exactly as CUB's
// stable sort of the id-ordered unsorted list does).
#include <cstdlib>
#include "common.cuh"

namespace dge {

#ifndef DGE_SCAN_THREADS
#define DGE_SCAN_THREADS 256
#endif
constexpr int SCAN_THREADS = DGE_SCAN_THREADS;  // A/B switch (>= 256 only: api.cu sizes block_sums for blocks of 256); 512: binning 0.829 -> 0.847 ms

__device__ __forceinline__ uint32_t rect_count(ushort4 r) {
  return (uint32_t)(r.z - r.x) * (uint32_t)(r.w - r.y);
}

// block sums of tiles_touched taken in depth order. A block covers SCAN_THREADS consecutive ranks with
// SCAN_THREADS / 4 threads, four ranks each: the two dependent loads per rank (order -> rect, a random 8-byte
// gather) are latency, and with one rank per thread a warp had one gather in flight (0.121 ms for config 2's 20 M
// ranks at 26 % of the issue rate and of the DRAM bandwidth); with four they overlap.
constexpr int SCAN_RED_THREADS = SCAN_THREADS / 4;
__global__ void __launch_bounds__(SCAN_RED_THREADS) scan_reduce_kernel(
    int P, const uint32_t* __restrict__ order, const ushort4* __restrict__ rect,
    uint32_t* __restrict__ block_sums, size_t view_stride) {
  if (blockIdx.y) {  // view of a batch (fit step): same arrays, view_stride bytes per view further on
    const size_t sh = blockIdx.y * view_stride;
    order = shift_ptr(order, sh);
    rect = shift_ptr(rect, sh);
    block_sums = shift_ptr(block_sums, sh);
  }
  const int r0 = blockIdx.x * SCAN_THREADS + threadIdx.x;
  uint32_t gid[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int r = r0 + k * SCAN_RED_THREADS;
    gid[k] = r < P ? order[r] : 0xFFFFFFFFu;
  }
  uint32_t c = 0;
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (gid[k] != 0xFFFFFFFFu) c += rect_count(rect[gid[k]]);
  c = __reduce_add_sync(0xFFFFFFFFu, c);
  __shared__ uint32_t ws[SCAN_RED_THREADS / 32];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < SCAN_RED_THREADS / 32; i++) t += ws[i];
    block_sums[blockIdx.x] = t;
  }
}

// in-place exclusive scan of the block sums (single CTA, 1024 threads, carried chunks)
__global__ void __launch_bounds__(1024) scan_block_sums_kernel(int n, uint32_t* __restrict__ sums,
                                                              size_t view_stride) {
  sums = shift_ptr(sums, blockIdx.x * view_stride);  // one CTA per view
  __shared__ uint32_t ws[32];
  __shared__ uint32_t carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < n ? sums[i] : 0;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = ws[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, o);
        if (lane >= o) w += y;
      }
      ws[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t carry = carry_s;
    const uint32_t wbase = warp ? ws[warp - 1] : 0;
    if (i < n) sums[i] = carry + wbase + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wbase + x;
    __syncthreads();
  }
}

// Emits the (tile id, Gaussian id) pairs of 256 depth-consecutive Gaussians.
// Offsets come from a CTA scan + the scanned block sums; each warp then writes its
// Gaussians' instances cooperatively (lane k handles the k-th instance of the warp,
// source Gaussian found from a ballot of the Gaussians' first positions) so the stores
// are coalesced however skewed the per-Gaussian tile counts are.
__global__ void __launch_bounds__(SCAN_THREADS) expand_kernel(
    int P, int grid_x, const uint32_t* __restrict__ order, const ushort4* __restrict__ rect,
    const uint32_t* __restrict__ block_prefix, uint32_t* __restrict__ offsets,
    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, size_t view_stride,
    const uint32_t* __restrict__ seg_off) {
  if (seg_off) {  // view of a batch: geometry arrays at a uniform stride, instances at seg_off[view]
    const size_t sh = blockIdx.y * view_stride;
    order = shift_ptr(order, sh);
    rect = shift_ptr(rect, sh);
    block_prefix = shift_ptr(block_prefix, sh);
    offsets = shift_ptr(offsets, sh);
    keys_out += seg_off[blockIdx.y];
    vals_out += seg_off[blockIdx.y];
  }
  __shared__ uint32_t ws[SCAN_THREADS / 32];
  __shared__ uint32_t s_off[SCAN_THREADS / 32][33];
  __shared__ uint32_t s_gid[SCAN_THREADS / 32][32];
  __shared__ ushort4 s_rect[SCAN_THREADS / 32][32];
  __shared__ float s_invw[SCAN_THREADS / 32][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = blockIdx.x * SCAN_THREADS + threadIdx.x;
  uint32_t gid = 0;
  ushort4 rc = make_ushort4(0, 0, 0, 0);
  if (r < P) {
    gid = order[r];
    rc = rect[gid];
  }
  const uint32_t cnt = rect_count(rc);
  uint32_t x = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) ws[warp] = x;
  __syncthreads();
  uint32_t wbase = block_prefix[blockIdx.x];
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; w++)
    if (w < warp) wbase += ws[w];
  const uint32_t incl = wbase + x;
  if (r < P) offsets[r] = incl;
  // The warp's Gaussians that emit anything, compacted in order: their first output position (relative to the
  // warp's), id, rect and 1 / rect width. Lane k of an iteration then needs "the last start <= k": starts are
  // distinct, so that is (#starts before the iteration's 32 positions) + (#head bits at or below k) - 1 — one
  // ballot, one redux.or and two popcounts instead of a 5-step search through shared memory.
  const uint32_t full = 0xFFFFFFFFu;
  const uint32_t ne = __ballot_sync(full, cnt > 0);
  const uint32_t wstart = __shfl_sync(full, incl - cnt, 0);
  const uint32_t wtotal = __shfl_sync(full, incl, 31) - wstart;
  if (cnt > 0) {
    const int ci = __popc(ne & ((1u << lane) - 1u));
    s_off[warp][ci] = incl - cnt - wstart;
    s_gid[warp][ci] = gid;
    s_rect[warp][ci] = rc;
    // row = j / w through a float reciprocal: j, w < 2^16, so floor((j + 0.5) * (1/w)) is exact
    s_invw[warp][ci] = __frcp_rn((float)((int)rc.z - (int)rc.x));
  }
  __syncwarp();
  const uint32_t my_start = lane < __popc(ne) ? s_off[warp][lane] : 0xFFFFFFFFu;
  const uint32_t upto = full >> (31 - lane);  // bits 0 .. lane
  for (uint32_t base = 0; base < wtotal; base += 32) {
    const uint32_t rel = my_start - base;  // (wraps for starts before this iteration: not < 32)
    const uint32_t heads = __reduce_or_sync(full, rel < 32u ? 1u << rel : 0u);
    const int before = __popc(__ballot_sync(full, my_start < base));
    const uint32_t k = base + lane;
    if (k < wtotal) {
      const int l = before + __popc(heads & upto) - 1;
      const uint32_t j = k - s_off[warp][l];
      const ushort4 q = s_rect[warp][l];
      const uint32_t w = q.z - q.x;
      const uint32_t row = __float2uint_rz(((float)j + 0.5f) * s_invw[warp][l]);
      const uint32_t ty = q.y + row, tx = q.x + (j - row * w);
      const uint32_t target = wstart + k;
      DGE_CHECK(seg_off == nullptr || target < seg_off[blockIdx.y + 1] - seg_off[blockIdx.y]);
      DGE_CHECK(l >= 0 && j < (uint32_t)(q.z - q.x) * (uint32_t)(q.w - q.y) && tx < (uint32_t)grid_x);
      keys_out[target] = ty * (uint32_t)grid_x + tx;
      vals_out[target] = s_gid[warp][l];
    }
  }
}

// identifyTileRanges (DGR/cuda_rasterizer/rasterizer_impl.cu:105-125) on 32-bit tile ids
__global__ void __launch_bounds__(256) tile_ranges_kernel(uint32_t R,
                                                         const uint32_t* __restrict__ tile_ids,
                                                         uint2* __restrict__ ranges, size_t img_stride,
                                                         const uint32_t* __restrict__ seg_off) {
  if (seg_off) {  // view of a batch; ranges stay relative to the view's own segment
    const uint32_t o = seg_off[blockIdx.y];
    R = seg_off[blockIdx.y + 1] - o;
    tile_ids += o;
    ranges = shift_ptr(ranges, blockIdx.y * img_stride);
  }
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= R) return;
  const uint32_t cur = tile_ids[i];
  if (i == 0) {
    ranges[cur].x = 0;
  } else {
    const uint32_t prev = tile_ids[i - 1];
    if (cur != prev) {
      ranges[prev].y = i;
      ranges[cur].x = i;
    }
  }
  if (i == R - 1) ranges[cur].y = R;
}

__global__ void debug_keys_kernel(uint32_t R, const uint32_t* __restrict__ tile_ids,
                                  const uint32_t* __restrict__ point_list,
                                  const float4* __restrict__ rec,
                                  uint64_t* __restrict__ keys_out) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= R) return;
  keys_out[i] = ((uint64_t)tile_ids[i] << 32) | __float_as_uint(rec[(size_t)point_list[i] * REC_F4 + 2].w);
}

cudaError_t launch_depth_sort(int P, GeomState& g, cudaStream_t stream) {
  // culled Gaussians carry key 0xFFFFFFFF and tile count 0: they sort last and emit nothing
  return sort_pairs(g.sort_key, g.sort_val, (uint32_t)P, 32, /*iota=*/true, g.sort_ws,
                    g.sort_ws_bytes, stream);
}

static int tile_bits(int T) {
  int bits = 0;
  while (bits < 32 && (1ll << bits) < (long long)T) bits++;
  return bits;
}

cudaError_t launch_binning(const ViewParams& vp, int R, GeomState& g, BinState& b, ImgState& img,
                           cudaStream_t stream) {
  const int T = vp.grid_x * vp.grid_y;
  cudaError_t e = cudaMemsetAsync(img.ranges, 0, sizeof(uint2) * (size_t)T, stream);
  if (e != cudaSuccess || R == 0) return e;
  const int P = vp.P;
  const int blocks = (P + SCAN_THREADS - 1) / SCAN_THREADS;
  const uint32_t* order = g.sort_val[0];
  scan_reduce_kernel<<<blocks, SCAN_RED_THREADS, 0, stream>>>(P, order, g.rect, g.block_sums, 0);
  scan_block_sums_kernel<<<1, 1024, 0, stream>>>(blocks, g.block_sums, 0);
  const int bits = tile_bits(T);
  const int passes = sort_num_passes(bits);
  uint32_t* keys[2] = {b.tile_ids, b.key_alt};
  uint32_t* vals[2] = {b.point_list, b.val_alt};
  expand_kernel<<<blocks, SCAN_THREADS, 0, stream>>>(P, vp.grid_x, order, g.rect, g.block_sums,
                                                     g.offsets, keys[passes & 1], vals[passes & 1], 0,
                                                     nullptr);
  DGE_LAUNCHED(3);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (passes > 0) {
    e = sort_pairs(keys, vals, (uint32_t)R, bits, /*iota=*/false, b.sort_ws, b.sort_ws_bytes, stream);
    if (e != cudaSuccess) return e;
  }
  tile_ranges_kernel<<<(R + 255) / 256, 256, 0, stream>>>((uint32_t)R, b.tile_ids, img.ranges, 0, nullptr);
  DGE_LAUNCHED(1);
  return cudaGetLastError();
}

// ---- fit step: the same stages for all V views of a step, one launch each (grid.y = view) ----
// seg_off[v] = sum of num_rendered of the views before v (num_rendered of view v = counters[0] of its
// geometry blob, written by the batched preprocess); lives in view 0's counters[64 .. 64+V].
__global__ void seg_offsets_kernel(int V, const uint32_t* __restrict__ counters0, size_t geom_stride,
                                   uint32_t* __restrict__ seg_off) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint32_t run = 0;
  for (int v = 0; v < V; v++) {
    seg_off[v] = run;
    run += *shift_ptr(counters0, v * geom_stride);
  }
  seg_off[V] = run;
}

cudaError_t launch_seg_offsets(const ViewBatch& vb, const GeomState& g0, uint32_t* seg_off, cudaStream_t stream) {
  seg_offsets_kernel<<<1, 32, 0, stream>>>(vb.V, g0.counters, vb.geom_stride, seg_off);
  DGE_LAUNCHED(1);
  return cudaGetLastError();
}

cudaError_t launch_depth_sort_batched(int P, const ViewBatch& vb, GeomState& g0, cudaStream_t stream) {
  return sort_pairs_segmented(g0.sort_key, g0.sort_val, (uint32_t)P, 32, /*iota=*/true, g0.sort_ws,
                              g0.sort_ws_bytes, vb.V, vb.geom_stride, nullptr, 0, stream);
}

// ---- single-pass stable partition by tile id (fit step, T <= PART_MAX_TILES) ----------------------
// The generic path sorts the (tile id, Gaussian id) pairs with ceil(log2 T / 8) onesweep passes, each
// with a decoupled look-back whose chains are as long as a view has sort tiles. For the batched fit
// step the partition is done in ONE deterministic pass without any look-back:
//   part_count    every CTA counts, per tile id, the instances of its run of PART_RUN consecutive
//                 instances (instances are in depth order) -> table[run][tile]
//   part_scan_*   per tile: exclusive prefix over the runs, on top of the tile's start in the view's
//                 list (exclusive scan of the per-tile totals, which also IS the `ranges` array, so
//                 identifyTileRanges and the sorted tile-id array disappear) -> table[run][tile] =
//                 position of the run's first instance of that tile
//   part_scatter  every CTA ranks its run stably (warp w owns the w-th eighth of the run; ballot
//                 ranking inside a warp, per-warp running positions in shared memory) and writes the
//                 Gaussian ids straight to their final position.
// Traffic per instance: 8 B written by expand, 4 + 8 B read, 4 B written (was 8 + 2 x 16 + 4 + 4).
constexpr int PART_THREADS = 256;
constexpr int PART_WARPS = PART_THREADS / 32;
#ifndef DGE_PART_WARP_ITEMS
#define DGE_PART_WARP_ITEMS 512
#endif
constexpr int PART_WARP_ITEMS = DGE_PART_WARP_ITEMS;  // A/B switch: 256 / 512 / 1024 -> binning 1.028 / 0.825 / 0.831 ms (config 2)
constexpr int PART_IPL = PART_WARP_ITEMS / 32;          // instances per lane, register resident
constexpr int PART_RUN = PART_WARPS * PART_WARP_ITEMS;  // instances per CTA
constexpr int PART_GROUPS = 16;                         // run groups of the two-level scan over runs
constexpr int PART_MAX_TILES = 2048;

// rows of the count table owned by view v start at row (seg_off[v] / PART_RUN + v): disjoint, since a
// view of n instances has ceil(n / PART_RUN) <= floor(n / PART_RUN) + 1 runs
__device__ __forceinline__ size_t part_row0(const uint32_t* seg_off, int v) {
  return (size_t)(seg_off[v] / PART_RUN) + (size_t)v;
}
size_t part_table_rows(uint32_t R_total, int V) { return (size_t)(R_total / PART_RUN) + (size_t)V + 1; }
size_t part_workspace_bytes(uint32_t R_total, int V, int T) {
  // count table + per-(view, group, tile) partial sums
  return sizeof(uint32_t) * (part_table_rows(R_total, V) * (size_t)T + (size_t)V * PART_GROUPS * T) + 256;
}

// Which run of PART_RUN instances a CTA of part_count / part_scatter owns. S_flat == 0: grid (run, segment).
// S_flat > 0 (second level of the two-level partition, many short segments of unknown length): a flat grid
// over the rows of the count table; the segment q owning row r is the last one with part_row0(q) <= r
// (part_row0 is strictly increasing), rows past the segment's last run do nothing.
struct PartRunRef {
  uint32_t o, n, lo;  // segment start in the arena, segment length, first instance of the run in the segment
  size_t row;         // row of the count table
  int seg;
};
__device__ __forceinline__ bool part_locate(const uint32_t* __restrict__ seg_off, int S_flat, PartRunRef& pr) {
  if (S_flat == 0) {
    pr.seg = blockIdx.y;
    pr.row = part_row0(seg_off, pr.seg) + blockIdx.x;
    pr.lo = blockIdx.x * (uint32_t)PART_RUN;
  } else {
    const size_t row = blockIdx.x;
    int a = 0, b = S_flat - 1;
    while (a < b) {
      const int m = (a + b + 1) >> 1;
      if (part_row0(seg_off, m) <= row) a = m; else b = m - 1;
    }
    pr.seg = a;
    pr.row = row;
    const size_t r = row - part_row0(seg_off, a);
    if (r >= (size_t)(0xFFFFFFFFu / PART_RUN)) return false;
    pr.lo = (uint32_t)r * (uint32_t)PART_RUN;
  }
  pr.o = seg_off[pr.seg];
  pr.n = seg_off[pr.seg + 1] - pr.o;
  return pr.lo < pr.n;
}

// bucket of an instance = (tile id >> shift) & mask: the tile id itself (single level), the tile group
// (first level of the two-level partition) or the tile inside its group (second level)
__global__ void __launch_bounds__(PART_THREADS) part_count_kernel(const uint32_t* __restrict__ tile_ids,
                                                                  const uint32_t* __restrict__ seg_off,
                                                                  int T, uint32_t* __restrict__ table, int S_flat,
                                                                  int shift, uint32_t mask) {
  extern __shared__ uint32_t s_cnt[];  // [T]
  PartRunRef pr;
  if (!part_locate(seg_off, S_flat, pr)) return;
  const uint32_t lo = pr.lo, n = pr.n;
  const uint32_t hi = min(n, lo + (uint32_t)PART_RUN);
  for (int t = threadIdx.x; t < T; t += PART_THREADS) s_cnt[t] = 0;
  __syncthreads();
  const uint32_t* keys = tile_ids + pr.o;
  // plain shared-memory atomics also for the few buckets of the first level: a warp-aggregated variant
  // (ballots per bucket bit, one add per distinct bucket) took 0.29 ms where this takes 0.08 ms (1080p / 2 M)
  for (uint32_t i = lo + threadIdx.x; i < hi; i += PART_THREADS) atomicAdd(&s_cnt[(keys[i] >> shift) & mask], 1u);
  __syncthreads();
  uint32_t* row = table + pr.row * (size_t)T;
  for (int t = threadIdx.x; t < T; t += PART_THREADS) row[t] = s_cnt[t];
}

// partial[v][g][t] = sum over the runs of group g of table[run][t]
__global__ void __launch_bounds__(256) part_scan_partial_kernel(const uint32_t* __restrict__ seg_off, int T,
                                                                const uint32_t* __restrict__ table,
                                                                uint32_t* __restrict__ partial) {
  const int v = blockIdx.z, g = blockIdx.y;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= T) return;
  const uint32_t n = seg_off[v + 1] - seg_off[v];
  const uint32_t runs = (n + PART_RUN - 1) / PART_RUN, per = (runs + PART_GROUPS - 1) / PART_GROUPS;
  const uint32_t r0 = min(runs, g * per), r1 = min(runs, r0 + per);
  const uint32_t* col = table + part_row0(seg_off, v) * (size_t)T + t;
  uint32_t sum = 0;
#pragma unroll 4
  for (uint32_t r = r0; r < r1; r++) sum += col[(size_t)r * T];
  partial[((size_t)v * PART_GROUPS + g) * T + t] = sum;
}

// Where the per-segment exclusive scan over the buckets goes (CTA = segment):
//  single level   segment = view, bucket = tile: the scan IS the view's `ranges` array
//  first level    segment = view, bucket = tile group c: start of segment q = view * G + c of the second
//                 level, seg2[q] = view_off[view] + scan (absolute arena offsets; seg2[V * G] = end)
//  second level   segment = q, bucket = tile c * 2^ts_shift + t of view q / G: `ranges` again, relative to
//                 the VIEW's list (the blend kernels add view_off[view] themselves)
struct PartScanOut {
  uint2* ranges;
  size_t img_stride;
  uint32_t* seg2;             // first level only
  const uint32_t* view_off;   // first and second level
  const uint32_t* seg_abs;    // second level: seg2
  int G, ts_shift, T_total;
};

// per segment: totals per bucket -> exclusive scan over buckets (PartScanOut); partial[v][g][t] becomes the
// position of group g's first instance of bucket t
__global__ void __launch_bounds__(1024) part_scan_tiles_kernel(int T, uint32_t* __restrict__ partial,
                                                               PartScanOut po) {
  __shared__ uint32_t ws[32];
  __shared__ uint32_t carry_s;
  const int v = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* part = partial + (size_t)v * PART_GROUPS * T;
  uint2* rg = nullptr;
  uint32_t shift_by = 0;  // second level: the segment's start inside its view's list
  int tile0 = 0;
  if (po.seg_abs) {
    const int view = v / po.G;
    rg = shift_ptr(po.ranges, view * po.img_stride);
    shift_by = po.seg_abs[v] - po.view_off[view];
    tile0 = (v - view * po.G) << po.ts_shift;
  } else if (!po.seg2) {
    rg = shift_ptr(po.ranges, v * po.img_stride);
  } else if (v == 0 && threadIdx.x == 0) {
    po.seg2[(size_t)gridDim.x * po.G] = po.view_off[gridDim.x];
  }
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < T; base += 1024) {
    const int t = base + threadIdx.x;
    uint32_t total = 0;
    if (t < T)
#pragma unroll
      for (int g = 0; g < PART_GROUPS; g++) total += part[(size_t)g * T + t];
    uint32_t x = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = ws[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, o);
        if (lane >= o) w += y;
      }
      ws[lane] = w;
    }
    __syncthreads();
    const uint32_t excl = carry_s + (warp ? ws[warp - 1] : 0) + x - total;
    if (t < T) {
      // identifyTileRanges leaves (0, 0) for tiles without instances (rasterizer_impl.cu:263-271)
      if (po.seg2)
        po.seg2[(size_t)v * po.G + t] = po.view_off[v] + excl;
      else if (tile0 + t < po.T_total)
        rg[tile0 + t] = total ? make_uint2(shift_by + excl, shift_by + excl + total) : make_uint2(0u, 0u);
      uint32_t run = excl;
#pragma unroll
      for (int g = 0; g < PART_GROUPS; g++) {
        const uint32_t c = part[(size_t)g * T + t];
        part[(size_t)g * T + t] = run;
        run += c;
      }
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + total;
    __syncthreads();
  }
}

// table[run][t]: count -> position of the run's first instance of tile t in the view's list
__global__ void __launch_bounds__(256) part_scan_runs_kernel(const uint32_t* __restrict__ seg_off, int T,
                                                             uint32_t* __restrict__ table,
                                                             const uint32_t* __restrict__ partial) {
  const int v = blockIdx.z, g = blockIdx.y;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= T) return;
  const uint32_t n = seg_off[v + 1] - seg_off[v];
  const uint32_t runs = (n + PART_RUN - 1) / PART_RUN, per = (runs + PART_GROUPS - 1) / PART_GROUPS;
  const uint32_t r0 = min(runs, g * per), r1 = min(runs, r0 + per);
  uint32_t* col = table + part_row0(seg_off, v) * (size_t)T + t;
  uint32_t run = partial[((size_t)v * PART_GROUPS + g) * T + t];
  for (uint32_t r = r0; r < r1; r += 8) {  // eight independent loads in flight, then the eight stores
    uint32_t c[8];
#pragma unroll
    for (int k = 0; k < 8; k++) c[k] = (r + k < r1) ? col[(size_t)(r + k) * T] : 0u;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (r + k < r1) col[(size_t)(r + k) * T] = run;
      run += c[k];
    }
  }
}

template <int BITS>  // bits of a bucket index (T <= 1 << BITS), compile-time so that the ranking unrolls
__global__ void __launch_bounds__(PART_THREADS) part_scatter_kernel(
    const uint32_t* __restrict__ tile_ids, const uint32_t* __restrict__ gids,
    const uint32_t* __restrict__ seg_off, int T, const uint32_t* __restrict__ table,
    uint32_t* __restrict__ point_list, uint32_t* __restrict__ keys_out, int S_flat, int shift, uint32_t mask) {
  extern __shared__ uint32_t s_part[];
  const int TP = (T + 1) >> 1;                      // packed u16 pairs per warp row
  uint32_t* s_cnt = s_part;                         // [PART_WARPS][TP]  two 16-bit counters per word
  uint32_t* s_base = s_part + PART_WARPS * TP;      // [PART_WARPS][T]
  PartRunRef pr;
  if (!part_locate(seg_off, S_flat, pr)) return;
  const uint32_t lo = pr.lo, n = pr.n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t* keys = tile_ids + pr.o;
  const uint32_t* vals = gids + pr.o;
  uint32_t* out = point_list + pr.o;
  uint32_t* kout = keys_out ? keys_out + pr.o : nullptr;  // first level: the tile ids travel along
  // (PART_WARPS * TP is a multiple of 4 words and s_cnt is 16-byte aligned)
  for (int i = threadIdx.x; i < PART_WARPS * TP / 4; i += PART_THREADS)
    reinterpret_cast<uint4*>(s_cnt)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  // A: per-warp counts of the warp's own PART_WARP_ITEMS consecutive instances. Keys and Gaussian ids
  // are loaded ONCE, all loads in flight together, and stay in registers for the ranking below.
  const uint32_t wlo = lo + warp * (uint32_t)PART_WARP_ITEMS;
  const uint32_t whi = min(n, wlo + (uint32_t)PART_WARP_ITEMS);
  uint32_t key[PART_IPL], val[PART_IPL];
#pragma unroll
  for (int k = 0; k < PART_IPL; k++) {
    const uint32_t i = wlo + k * 32 + lane;
    key[k] = i < whi ? keys[i] : 0xFFFFFFFFu;
  }
#pragma unroll
  for (int k = 0; k < PART_IPL; k++) {
    const uint32_t i = wlo + k * 32 + lane;
    val[k] = i < whi ? vals[i] : 0u;
  }
#pragma unroll
  for (int k = 0; k < PART_IPL; k++)
    if (key[k] != 0xFFFFFFFFu) {
      const uint32_t bk = (key[k] >> shift) & mask;
      atomicAdd(&s_cnt[warp * TP + (bk >> 1)], 1u << (16 * (bk & 1)));
    }
  __syncthreads();
  // B: position of each warp's first instance of every bucket, two buckets (one packed counter word) per
  // step: the table is as large as the run (8 warps x T entries for 4096 instances), so this loop
  // was 40 % of the kernel's instructions when written tile by tile
  const uint32_t* row = table + pr.row * (size_t)T;
  for (int t2 = threadIdx.x; t2 < TP; t2 += PART_THREADS) {
    const int t = 2 * t2;
    uint32_t run0 = row[t], run1 = (t + 1 < T) ? row[t + 1] : 0u;
#pragma unroll
    for (int w = 0; w < PART_WARPS; w++) {
      const uint32_t c = s_cnt[w * TP + t2];
      s_base[w * T + t] = run0;
      if (t + 1 < T) s_base[w * T + t + 1] = run1;
      run0 += c & 0xFFFFu;
      run1 += c >> 16;
    }
  }
  __syncthreads();
  // C: stable ranking, 32 instances at a time in list order
  uint32_t* my_base = s_base + warp * T;
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int k = 0; k < PART_IPL; k++) {
    if (wlo + k * 32 >= whi) break;  // warp-uniform
    const bool valid = key[k] != 0xFFFFFFFFu;
    const uint32_t t = (key[k] >> shift) & mask;
    // lanes holding the same bucket (common.cuh:same_value_lanes; with match.any this kernel took
    // 0.78 ms instead of 0.51 ms)
    uint32_t peers = __ballot_sync(0xFFFFFFFFu, valid);
    if (!valid) peers = ~peers;
    peers = same_value_lanes<BITS>(t, peers);
    const int leader = __ffs(peers) - 1;
    uint32_t prev = 0;
    if (valid && lane == leader) {
      prev = my_base[t];
      my_base[t] = prev + __popc(peers);
    }
    prev = __shfl_sync(0xFFFFFFFFu, prev, leader);
    if (valid) {
      const uint32_t pos = prev + __popc(peers & lt_mask);
      DGE_CHECK(pos < n);
      out[pos] = val[k];
      if (kout) kout[pos] = key[k];
    }
    __syncwarp();
  }
}

// part_scatter for at most 32 buckets (the tile groups of the first level; tiny images): lane b owns
// bucket b. Its count and its running position live in a REGISTER of lane b, the lanes holding a bucket's
// instances are found with one ballot per bucket-index bit (every lane evaluates "who holds MY bucket", a
// lane's peers are then lane bucket's mask), so there are no shared-memory atomics (with few buckets nearly
// all lanes of a warp hit the same counter word) and no shared-memory read-modify-write chain in the
// ranking loop: ~50 instead of ~95 warp instructions per 32 instances.
template <int BITS>
__device__ __forceinline__ uint32_t lanes_of_my_bucket(uint32_t bk, bool valid, const uint32_t (&flip)[BITS]) {
  uint32_t m = __ballot_sync(0xFFFFFFFFu, valid);
#pragma unroll
  for (int j = 0; j < BITS; j++) m &= __ballot_sync(0xFFFFFFFFu, valid && ((bk >> j) & 1u)) ^ flip[j];
  return m;
}

// 3 CTAs per SM = 80 registers: the compiler keeps the masks of pass A for pass C (measured: binning 2.15 ms
// at 1080p / 2 M against 2.21 ms with 64 registers and the masks recomputed)
#ifndef DGE_PART_SMALL_MIN_CTAS
#define DGE_PART_SMALL_MIN_CTAS 3
#endif
template <int BITS>
__global__ void __launch_bounds__(PART_THREADS, DGE_PART_SMALL_MIN_CTAS) part_scatter_small_kernel(
    const uint32_t* __restrict__ tile_ids, const uint32_t* __restrict__ gids,
    const uint32_t* __restrict__ seg_off, int T, const uint32_t* __restrict__ table,
    uint32_t* __restrict__ point_list, uint32_t* __restrict__ keys_out, int S_flat, int shift, uint32_t mask) {
  __shared__ uint32_t s_cnt[PART_WARPS][32];
  PartRunRef pr;
  if (!part_locate(seg_off, S_flat, pr)) return;
  const uint32_t lo = pr.lo, n = pr.n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t* keys = tile_ids + pr.o;
  const uint32_t* vals = gids + pr.o;
  uint32_t* out = point_list + pr.o;
  uint32_t* kout = keys_out ? keys_out + pr.o : nullptr;
  const uint32_t wlo = lo + warp * (uint32_t)PART_WARP_ITEMS;
  const uint32_t whi = min(n, wlo + (uint32_t)PART_WARP_ITEMS);
  uint32_t key[PART_IPL], val[PART_IPL];
#pragma unroll
  for (int k = 0; k < PART_IPL; k++) {
    const uint32_t i = wlo + k * 32 + lane;
    key[k] = i < whi ? keys[i] : 0xFFFFFFFFu;
  }
#pragma unroll
  for (int k = 0; k < PART_IPL; k++) {
    const uint32_t i = wlo + k * 32 + lane;
    val[k] = i < whi ? vals[i] : 0u;
  }
  uint32_t flip[BITS];  // bit j of the bucket this lane owns: clear -> the ballot of bit j is inverted
#pragma unroll
  for (int j = 0; j < BITS; j++) flip[j] = ((lane >> j) & 1) ? 0u : 0xFFFFFFFFu;
  // A: instances of bucket `lane` in this warp's 512
  uint32_t cnt = 0;
#pragma unroll
  for (int k = 0; k < PART_IPL; k++) {
    const bool valid = key[k] != 0xFFFFFFFFu;
    cnt += __popc(lanes_of_my_bucket<BITS>((key[k] >> shift) & mask, valid, flip));
  }
  s_cnt[warp][lane] = cnt;
  __syncthreads();
  // B: position of this warp's first instance of bucket `lane`
  uint32_t base = lane < T ? table[pr.row * (size_t)T + lane] : 0u;
#pragma unroll
  for (int w = 0; w < PART_WARPS; w++)
    if (w < warp) base += s_cnt[w][lane];
  // C: stable ranking, 32 instances at a time in list order
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int k = 0; k < PART_IPL; k++) {
    if (wlo + k * 32 >= whi) break;  // warp-uniform
    const bool valid = key[k] != 0xFFFFFFFFu;
    const uint32_t bk = (key[k] >> shift) & mask;
    const uint32_t mine = lanes_of_my_bucket<BITS>(bk, valid, flip);
    const uint32_t peers = __shfl_sync(0xFFFFFFFFu, mine, bk & 31u);
    const uint32_t prev = __shfl_sync(0xFFFFFFFFu, base, bk & 31u);
    base += __popc(mine);
    if (valid) {
      const uint32_t pos = prev + __popc(peers & lt_mask);
      DGE_CHECK(pos < n);
      out[pos] = val[k];
      if (kout) kout[pos] = key[k];
    }
  }
}

// One level of the partition: the instances of every segment (seg[s] .. seg[s + 1] of the arena, S
// segments) are stably partitioned by bucket = (tile id >> shift) & mask, T buckets per segment.
// flat_rows == 0: grid (runs_max, S), the caller knows an upper bound on a segment's runs; otherwise a
// flat grid of flat_rows count-table rows (part_locate).
static cudaError_t part_level(const uint32_t* seg, int S, int runs_max, size_t flat_rows, int T, int shift,
                              uint32_t mask, const uint32_t* keys, const uint32_t* vals, uint32_t* vals_out,
                              uint32_t* keys_out, uint32_t* table, uint32_t* partial, const PartScanOut& po,
                              cudaStream_t stream) {
  const int tb = (T + 255) / 256;
  const size_t scatter_smem = sizeof(uint32_t) * PART_WARPS * (size_t)(((T + 1) >> 1) + T);
  int id_bits = 1;  // bits that tell buckets 0 .. T-1 apart
  while ((1 << id_bits) < T) id_bits++;
  using scatter_fn = void (*)(const uint32_t*, const uint32_t*, const uint32_t*, int, const uint32_t*, uint32_t*,
                              uint32_t*, int, int, uint32_t);
  static const scatter_fn scatter[12] = {nullptr, part_scatter_kernel<1>, part_scatter_kernel<2>, part_scatter_kernel<3>,
                                         part_scatter_kernel<4>, part_scatter_kernel<5>, part_scatter_kernel<6>,
                                         part_scatter_kernel<7>, part_scatter_kernel<8>, part_scatter_kernel<9>,
                                         part_scatter_kernel<10>, part_scatter_kernel<11>};
  static bool attr_set[12] = {};
  if (id_bits > 11) return cudaErrorInvalidValue;  // T <= PART_MAX_TILES = 2048
  if (!attr_set[id_bits]) {
    cudaError_t e = cudaFuncSetAttribute(scatter[id_bits], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(sizeof(uint32_t) * PART_WARPS * (size_t)(PART_MAX_TILES / 2 + PART_MAX_TILES)));
    if (e != cudaSuccess) return e;
    attr_set[id_bits] = true;
  }
  const dim3 run_grid = flat_rows ? dim3((unsigned)flat_rows) : dim3(runs_max, S);
  const int S_flat = flat_rows ? S : 0;
  static const scatter_fn scatter_small[6] = {nullptr, part_scatter_small_kernel<1>, part_scatter_small_kernel<2>,
                                              part_scatter_small_kernel<3>, part_scatter_small_kernel<4>,
                                              part_scatter_small_kernel<5>};
  static const bool no_small = getenv("DGE_PART_NO_SMALL") != nullptr;  // A/B
  const bool small = T <= 32 && !no_small;
  part_count_kernel<<<run_grid, PART_THREADS, sizeof(uint32_t) * T, stream>>>(keys, seg, T, table, S_flat, shift, mask);
  part_scan_partial_kernel<<<dim3(tb, PART_GROUPS, S), 256, 0, stream>>>(seg, T, table, partial);
  part_scan_tiles_kernel<<<S, 1024, 0, stream>>>(T, partial, po);
  part_scan_runs_kernel<<<dim3(tb, PART_GROUPS, S), 256, 0, stream>>>(seg, T, table, partial);
  if (small)
    scatter_small[id_bits]<<<run_grid, PART_THREADS, 0, stream>>>(keys, vals, seg, T, table, vals_out, keys_out,
                                                                  S_flat, shift, mask);
  else
    scatter[id_bits]<<<run_grid, PART_THREADS, scatter_smem, stream>>>(keys, vals, seg, T, table, vals_out, keys_out,
                                                                       S_flat, shift, mask);
  DGE_LAUNCHED(5);
  return cudaGetLastError();
}

static cudaError_t partition_by_tile(const ViewBatch& vb, int T, int bits, uint32_t R_total, uint32_t R_max,
                                     const uint32_t* keys, const uint32_t* vals, uint32_t* point_list,
                                     uint32_t* ws, size_t ws_bytes, ImgState& img0, cudaStream_t stream) {
  (void)bits;
  if (part_workspace_bytes(R_total, vb.V, T) > ws_bytes) return cudaErrorInvalidValue;
  uint32_t* table = ws;
  uint32_t* partial = ws + part_table_rows(R_total, vb.V) * (size_t)T;
  const int runs = (int)((R_max + PART_RUN - 1) / PART_RUN);
  const PartScanOut po{img0.ranges, vb.img_stride, nullptr, nullptr, nullptr, 1, 0, T};
  return part_level(vb.seg_off, vb.V, runs, 0, T, 0, 0xFFFFFFFFu, keys, vals, point_list, nullptr, table, partial,
                    po, stream);
}

// ---- two-level partition (fit step, T > PART_MAX_TILES) --------------------------------------------
// 1264x832 has 4108 tiles and 1080p 8160: too many for the per-warp tables of part_scatter, and a view's
// list (270 MB at 6 M Gaussians / 1080p) is far larger than the L2 in which the scattered 4-byte stores of a
// single pass have to meet. Two passes of the SAME kernels keep both bounded:
//   level 1  by tile group (tile id >> PART2 shift; G = ceil(T / TS) groups of TS = 256 tiles): few
//            output streams per view, every warp writes long pieces; the tile ids travel with the
//            Gaussian ids. Its scan over the groups is the segment table of level 2.
//   level 2  inside every (view, group) segment by the low bits of the tile id: TS buckets, the
//            scattered stores stay inside the group's piece of the list (a few MB); its scan is `ranges`.
// Both levels are stable, so the result equals the stable sort by tile id (and the reference's list).
// Traffic per instance: 8 B written by expand, 4 + 8 read and 8 written (level 1), 4 + 8 read and 4
// written (level 2) = 44 B — about what the generic path moves (8 + 4 for the histogram + 2 x 16 for the two
// onesweep passes + 4 for identifyTileRanges = 48 B); what it saves is the decoupled look-back (status traffic,
// spinning, 5 warp-instructions per instance): 2.62 -> 1.96 ms per 8 views at 1080p / 2 M Gaussians.
static int part2_shift() {
  static const int s = [] {
    const char* e = getenv("DGE_PART2_SHIFT");
    const int v = e ? atoi(e) : 8;
    return v < 1 ? 1 : (v > 11 ? 11 : v);
  }();
  return s;
}
static int part2_groups(int T) { return (T + (1 << part2_shift()) - 1) >> part2_shift(); }
static size_t align64w(size_t words) { return (words + 63) & ~(size_t)63; }
size_t part2_workspace_bytes(uint32_t R_total, int V, int T) {
  const int G = part2_groups(T), TS = 1 << part2_shift();
  const size_t S = (size_t)V * G;
  const size_t words = align64w(part_table_rows(R_total, V) * (size_t)G) + align64w((size_t)V * PART_GROUPS * G) +
                       align64w(S + 1) + align64w(part_table_rows(R_total, (int)S) * (size_t)TS) +
                       align64w(S * PART_GROUPS * TS);
  return sizeof(uint32_t) * words + 256;
}

static cudaError_t partition_two_level(const ViewBatch& vb, int T, uint32_t R_total, uint32_t R_max,
                                       const uint32_t* keys, const uint32_t* vals, uint32_t* keys_tmp,
                                       uint32_t* vals_tmp, uint32_t* point_list, uint32_t* ws, size_t ws_bytes,
                                       ImgState& img0, cudaStream_t stream) {
  if (part2_workspace_bytes(R_total, vb.V, T) > ws_bytes) return cudaErrorInvalidValue;
  const int shift = part2_shift(), TS = 1 << shift, G = part2_groups(T);
  const int S = vb.V * G;
  if (G > PART_MAX_TILES || S > 65535) return cudaErrorInvalidValue;
  uint32_t* table1 = ws;
  uint32_t* partial1 = table1 + align64w(part_table_rows(R_total, vb.V) * (size_t)G);
  uint32_t* seg2 = partial1 + align64w((size_t)vb.V * PART_GROUPS * G);
  uint32_t* table2 = seg2 + align64w((size_t)S + 1);
  uint32_t* partial2 = table2 + align64w(part_table_rows(R_total, S) * (size_t)TS);
  const int runs = (int)((R_max + PART_RUN - 1) / PART_RUN);
  const PartScanOut po1{nullptr, 0, seg2, vb.seg_off, nullptr, G, shift, T};
  cudaError_t e = part_level(vb.seg_off, vb.V, runs, 0, G, shift, 0xFFFFFFFFu, keys, vals, vals_tmp, keys_tmp,
                             table1, partial1, po1, stream);
  if (e != cudaSuccess) return e;
  const PartScanOut po2{img0.ranges, vb.img_stride, nullptr, vb.seg_off, seg2, G, shift, T};
  return part_level(seg2, S, 0, part_table_rows(R_total, S), TS, 0, (uint32_t)TS - 1u, keys_tmp, vals_tmp,
                    point_list, nullptr, table2, partial2, po2, stream);
}

// T > PART_MAX_TILES goes through the two-level partition. DGE_PART2=0 puts the generic onesweep passes
// back (A/B), DGE_PART2=2 forces the two-level partition for EVERY tile count (tests; with DGE_PART2_SHIFT=4
// even small images have several tile groups).
static int part2_mode() {
  static const int m = [] {
    const char* e = getenv("DGE_PART2");
    return e ? atoi(e) : 1;
  }();
  return m;
}

size_t binning_batched_workspace_bytes(uint32_t R_total, int V, int T) {
  const size_t a = sort_workspace_bytes_segmented(R_total, V);
  const size_t b = T <= PART_MAX_TILES ? part_workspace_bytes(R_total, V, T) : 0;
  const size_t c = part2_mode() ? part2_workspace_bytes(R_total, V, T) : 0;
  return a > b ? (a > c ? a : c) : (b > c ? b : c);
}

cudaError_t launch_binning_batched(const ViewParams& vp, const ViewBatch& vb, uint32_t R_total, uint32_t R_max,
                                   GeomState& g0, BinState& b, ImgState& img0, cudaStream_t stream) {
  const int T = vp.grid_x * vp.grid_y;
  cudaError_t e = cudaMemset2DAsync(img0.ranges, vb.img_stride, 0, sizeof(uint2) * (size_t)T, (size_t)vb.V, stream);
  if (e != cudaSuccess || R_total == 0) return e;
  const int P = vp.P;
  const int blocks = (P + SCAN_THREADS - 1) / SCAN_THREADS;
  const uint32_t* order = g0.sort_val[0];
  scan_reduce_kernel<<<dim3(blocks, vb.V), SCAN_RED_THREADS, 0, stream>>>(P, order, g0.rect, g0.block_sums,
                                                                      vb.geom_stride);
  scan_block_sums_kernel<<<vb.V, 1024, 0, stream>>>(blocks, g0.block_sums, vb.geom_stride);
  const int bits = tile_bits(T);
  const int passes = sort_num_passes(bits);
  static const bool no_part = getenv("DGE_NO_PARTITION") != nullptr;
  const bool two_level = passes > 0 && !no_part && part2_groups(T) <= PART_MAX_TILES &&
                         (size_t)vb.V * part2_groups(T) <= 65535 &&
                         (part2_mode() == 2 || (part2_mode() == 1 && T > PART_MAX_TILES));
  const bool one_pass = passes > 0 && T <= PART_MAX_TILES && !no_part && !two_level;
  uint32_t* keys[2] = {b.tile_ids, b.key_alt};
  uint32_t* vals[2] = {b.point_list, b.val_alt};
  // the single-pass partition reads (key_alt, val_alt) and writes point_list; the two-level one goes
  // (tile_ids, point_list) -> (key_alt, val_alt) -> point_list; the generic sort ping-pongs and must
  // END in (tile_ids, point_list)
  const int src = two_level ? 0 : one_pass ? 1 : (passes & 1);
  expand_kernel<<<dim3(blocks, vb.V), SCAN_THREADS, 0, stream>>>(P, vp.grid_x, order, g0.rect, g0.block_sums,
                                                                 g0.offsets, keys[src], vals[src],
                                                                 vb.geom_stride, vb.seg_off);
  DGE_LAUNCHED(3);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (two_level)
    return partition_two_level(vb, T, R_total, R_max, keys[0], vals[0], keys[1], vals[1], b.point_list, b.sort_ws,
                               b.sort_ws_bytes, img0, stream);
  if (one_pass)  // one deterministic pass; writes point_list and ranges
    return partition_by_tile(vb, T, bits, R_total, R_max, keys[1], vals[1], b.point_list, b.sort_ws,
                             b.sort_ws_bytes, img0, stream);
  if (passes > 0) {
    e = sort_pairs_segmented(keys, vals, R_max, bits, /*iota=*/false, b.sort_ws, b.sort_ws_bytes, vb.V, 0,
                             vb.seg_off, R_total, stream);
    if (e != cudaSuccess) return e;
  }
  tile_ranges_kernel<<<dim3((R_max + 255) / 256, vb.V), 256, 0, stream>>>(R_max, b.tile_ids, img0.ranges,
                                                                          vb.img_stride, vb.seg_off);
  DGE_LAUNCHED(1);
  return cudaGetLastError();
}

cudaError_t launch_debug_keys(const GeomState& g, const BinState& b, int R, uint64_t* keys_out,
                              cudaStream_t stream) {
  if (R == 0) return cudaSuccess;
  debug_keys_kernel<<<(R + 255) / 256, 256, 0, stream>>>((uint32_t)R, b.tile_ids, b.point_list,
                                                         g.rec, keys_out);
  return cudaGetLastError();
}

}  // namespace dge
