// K1 / K13 / K12: per-Gaussian projection, EWA covariance, SH colour, tile rect.
//
// Replaces preprocessCUDA (DGR/cuda_rasterizer/forward.cu:155-256), its clone
// preprocessCUDA_apply_weights (DGR/cuda_rasterizer/apply_weights.cu:148-234) and
// checkFrustum (DGR/cuda_rasterizer/rasterizer_impl.cu:53-63).
//
// Tile assignment and the depth bits that go into the sort key must be BIT-EXACT
// with the reference, so every floating-point operation below is written with
// explicit round-to-nearest intrinsics in the order the reference's own build
// (nvcc 12.9, -fmad=true, sm_100a) contracts them — read off its SASS
// (cuobjdump -sass of oracle/_ref/forward.o; see DESIGN.md "bit-exact preprocess").
// The same sequence is restated on the CPU in oracle/splat_oracle.c.
#include <cstdio>
#include "common.cuh"
#include "math_ref.cuh"

#ifndef DGE_BOX_EXACT_MATH
#define DGE_BOX_EXACT_MATH 0
#endif

namespace dge {

// DGR/cuda_rasterizer/forward.cu:20-71 computeColorFromSH; sh = 3*M floats of this Gaussian.
// Returns result BEFORE the +0.5 in res[3].
template <typename F>
__device__ __forceinline__ void sh_to_rgb_ref(int deg, float dx, float dy, float dz, F sh,
                                              float res[3]) {
  const float len = __fsqrt_rn(FMA(dz, dz, FMA(dx, dx, MUL(dy, dy))));
  const float x = __fdiv_rn(dx, len), y = __fdiv_rn(dy, len), z = __fdiv_rn(dz, len);
#pragma unroll
  for (int c = 0; c < 3; c++) res[c] = MUL(sh(0, c), SH_C0);
  if (deg > 0) {
    const float t1 = MUL(y, SH_C1), t2 = MUL(z, SH_C1), t3 = MUL(x, SH_C1);
#pragma unroll
    for (int c = 0; c < 3; c++) {
      float r = FMA(-t1, sh(1, c), res[c]);
      r = FMA(t2, sh(2, c), r);
      res[c] = FMA(-t3, sh(3, c), r);
    }
    if (deg > 1) {
      const float xx = MUL(x, x), yy = MUL(y, y), zz = MUL(z, z);
      const float xy = MUL(y, x), yz = MUL(z, y), xz = MUL(z, x);
      const float zz2 = ADD(zz, zz);
      const float k4 = MUL(xy, SH_C2_0), k5 = MUL(yz, SH_C2_1);
      const float k6 = MUL(ADD(-yy, ADD(-xx, zz2)), SH_C2_2);
      const float k7 = MUL(xz, SH_C2_3);
      const float xx_m_yy = ADD(xx, -yy);
      const float k8 = MUL(xx_m_yy, SH_C2_4);
#pragma unroll
      for (int c = 0; c < 3; c++) {
        float r = FMA(k4, sh(4, c), res[c]);
        r = FMA(k5, sh(5, c), r);
        r = FMA(k6, sh(6, c), r);
        r = FMA(k7, sh(7, c), r);
        res[c] = FMA(k8, sh(8, c), r);
      }
      if (deg > 2) {
        const float k9 = MUL(MUL(y, SH_C3_0), FMA(xx, 3.0f, -yy));
        const float k10 = MUL(MUL(xy, SH_C3_1), z);
        const float f4 = ADD(-yy, FMA(zz, 4.0f, -xx));  // 4zz - xx - yy
        const float k11 = MUL(MUL(y, SH_C3_2), f4);
        const float k12 = MUL(MUL(z, SH_C3_3), FMA(yy, -3.0f, FMA(xx, -3.0f, zz2)));
        const float k13 = MUL(f4, MUL(x, SH_C3_4));
        const float k14 = MUL(xx_m_yy, MUL(z, SH_C3_5));
        const float k15 = MUL(MUL(x, SH_C3_6), FMA(yy, -3.0f, xx));
#pragma unroll
        for (int c = 0; c < 3; c++) {
          float r = FMA(k9, sh(9, c), res[c]);
          r = FMA(k10, sh(10, c), r);
          r = FMA(k11, sh(11, c), r);
          r = FMA(k12, sh(12, c), r);
          r = FMA(k13, sh(13, c), r);
          r = FMA(k14, sh(14, c), r);
          res[c] = FMA(k15, sh(15, c), r);
        }
      }
    }
  }
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg((const float4*)p); }

// What one view's projection of one Gaussian produces (radius == 0 <=> culled).
struct PreOut {
  int radius;
  uint32_t tiles, key;
  uint8_t clamp_bits;
  ushort4 rect;
  float4 rec[REC_F4];  // blend record (common.cuh:pack_record), valid when radius > 0
};

// The per-Gaussian forward of ONE view (forward.cu:186-255), shared by the per-view kernel and the
// batched one. cov3(c) fills the world-space covariance, color(rgb) the colour BEFORE clamping
// flags are taken (returns the clamp mask), opacity() loads the opacity; all three are only
// called for Gaussians that survive the culls, as in the reference.
template <typename COV3, typename COLOR, typename OPAC>
__device__ __forceinline__ PreOut preprocess_view(const ViewParams& vp, const float* V, const float* Pm,
                                                  float px, float py, float pz, bool prefiltered,
                                                  COV3 cov3f, COLOR colorf, OPAC opacf) {
  PreOut o;
  o.radius = 0;
  o.tiles = 0;
  o.key = 0xFFFFFFFFu;
  o.clamp_bits = 0;
  o.rect = make_ushort4(0, 0, 0, 0);
  // in_frustum (auxiliary.h:139-164): near-plane test only
  const float depth = xform_row(V, 2, px, py, pz);
  if (depth > 0.2f) {
    const float hw = ADD(xform_row(Pm, 3, px, py, pz), 0.0000001f);
    const float p_w = __frcp_rn(hw);
    const float projx = MUL(xform_row(Pm, 0, px, py, pz), p_w);
    const float projy = MUL(xform_row(Pm, 1, px, py, pz), p_w);
    float cov3[6];
    cov3f(cov3);
    const float tx = xform_row(V, 0, px, py, pz), ty = xform_row(V, 1, px, py, pz);
    const float3 cov = cov2d_ref(tx, ty, depth, vp.tan_fovx, vp.tan_fovy, vp.focal_x, vp.focal_y, V, cov3);
    const float det = FMA(cov.x, cov.z, -MUL(cov.y, cov.y));
    if (det != 0.0f) {
      const float inv = __frcp_rn(det);
      const float mid = MUL(ADD(cov.x, cov.z), 0.5f);
      const float s = __fsqrt_rn(fmaxf(FMA(mid, mid, -det), 0.1f));
      const float lam = fmaxf(ADD(mid, s), ADD(mid, -s));
      const int rad = (int)ceilf(MUL(__fsqrt_rn(lam), 3.0f));
      // ndc2Pix in FP64 (auxiliary.h:41-44): ((v + 1.0) * S - 1.0) * 0.5 with one DFMA
      const float pix_x =
          (float)__dmul_rn(__fma_rn(__dadd_rn((double)projx, 1.0), (double)vp.W, -1.0), 0.5);
      const float pix_y =
          (float)__dmul_rn(__fma_rn(__dadd_rn((double)projy, 1.0), (double)vp.H, -1.0), 0.5);
      int x0, y0, x1, y1;
      tile_rect(pix_x, pix_y, rad, vp.grid_x, vp.grid_y, x0, y0, x1, y1);
      const uint32_t cnt = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0);
      if (cnt != 0) {
        float rgb[3];
        o.clamp_bits = colorf(rgb);
        const float con_x = MUL(cov.z, inv), con_y = MUL(cov.y, -inv), con_z = MUL(cov.x, inv);
        const float opac = opacf();
        // Conservative screen-space box of {alpha >= 1/255}: power >= -ln(255 o) bounds the
        // quadratic form, whose extreme |dx|, |dy| are sqrt(2t cz/det), sqrt(2t cx/det). Used by
        // the blend kernels only to SKIP work; every skipped pair fails the exact test too.
        float hx = __int_as_float(0xff800000), hy = hx;  // -inf: never visible
        if (opac > 0.0f) {
          // (fast log / divide / square root: their ~1e-6 relative error disappears in the 0.02 added to the
          // logarithm and the 1.0001 x + 0.01 px margin; the IEEE versions were 6 % of this kernel's instructions)
#if DGE_BOX_EXACT_MATH
          const float k2 = 2.0f * (logf(255.0f * opac) + 0.02f);
#else
          const float k2 = 2.0f * (__logf(255.0f * opac) + 0.02f);
#endif
          const float dc = con_x * con_z - con_y * con_y;
          if (k2 > 0.0f) {
            if (dc > 0.0f) {
#if DGE_BOX_EXACT_MATH
              hx = sqrtf(k2 * con_z / dc) * 1.0001f + 0.01f;
              hy = sqrtf(k2 * con_x / dc) * 1.0001f + 0.01f;
#else
              const float t = __fdividef(k2, dc);
              float sx, sy;
              asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sx) : "f"(t * con_z));
              asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sy) : "f"(t * con_x));
              hx = sx * 1.0001f + 0.01f;
              hy = sy * 1.0001f + 0.01f;
#endif
            } else {
              hx = hy = __int_as_float(0x7f800000);
            }
          }
        }
        pack_record(o.rec, pix_x, pix_y, hx, hy, con_x, con_y, con_z, opac, rgb[0], rgb[1], rgb[2], depth);
        o.radius = rad;
        o.rect = make_ushort4(x0, y0, x1, y1);
        o.key = __float_as_uint(depth);
        o.tiles = cnt;
      }
    }
  } else if (prefiltered) {
    // auxiliary.h:156-160: the reference traps here
    printf("Point is filtered although prefiltered is set. This shouldn't happen!");
    __trap();
  }
  return o;
}

// SH colour of one Gaussian for one camera position + clamp mask (forward.cu:20-71, :241-247)
template <typename SHF>
__device__ __forceinline__ uint8_t sh_color(int D, float px, float py, float pz, float cx, float cy, float cz,
                                            SHF sh, float rgb[3]) {
  const float ddx = ADD(px, -cx), ddy = ADD(py, -cy), ddz = ADD(pz, -cz);
  sh_to_rgb_ref(D, ddx, ddy, ddz, sh, rgb);
  uint8_t cl = 0;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    rgb[c] = ADD(rgb[c], 0.5f);
    if (rgb[c] < 0.0f) cl |= (1u << c);
    rgb[c] = fmaxf(rgb[c], 0.0f);
  }
  return cl;
}

// One thread per Gaussian, 256 per CTA. Loads: xyz/scale 3 coalesced scalar streams,
// quaternion one float4, SH 12 float4 (M==16) issued together only by surviving lanes.
__global__ void __launch_bounds__(256) preprocess_kernel(
    ViewParams vp, const float* __restrict__ means3D, const float* __restrict__ scales,
    const float* __restrict__ rotations, const float* __restrict__ opacities,
    const float* __restrict__ shs, const float* __restrict__ cov3D_precomp,
    const float* __restrict__ colors_precomp, int colors_mode, bool prefiltered,
    int* __restrict__ radii, GeomState g, uint8_t* __restrict__ flags_out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t tiles = 0;
  if (idx < vp.P) {
    float V[16], Pm[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
      V[i] = __ldg(vp.view + i);
      Pm[i] = __ldg(vp.proj + i);
    }
    const float px = __ldg(means3D + 3 * idx), py = __ldg(means3D + 3 * idx + 1),
                pz = __ldg(means3D + 3 * idx + 2);
    const PreOut o = preprocess_view(
        vp, V, Pm, px, py, pz, prefiltered,
        [&](float* cov3) {
          if (cov3D_precomp != nullptr) {
#pragma unroll
            for (int i = 0; i < 6; i++) cov3[i] = __ldg(cov3D_precomp + 6 * (size_t)idx + i);
          } else {
            const float4 q = ldg4(rotations + 4 * (size_t)idx);
            cov3d_from_scale_rot(__ldg(scales + 3 * idx), __ldg(scales + 3 * idx + 1),
                                 __ldg(scales + 3 * idx + 2), vp.scale_modifier, q, cov3);
          }
        },
        [&](float* rgb) -> uint8_t {
          if (colors_mode == 0) {
            const float cx = __ldg(vp.campos), cy = __ldg(vp.campos + 1), cz = __ldg(vp.campos + 2);
            uint8_t cl;
            if (vp.M == 16) {
              float4 v[12];
              const float* base = shs + 48 * (size_t)idx;
#pragma unroll
              for (int i = 0; i < 12; i++) v[i] = ldg4(base + 4 * i);
              const float* f = reinterpret_cast<const float*>(v);
              cl = sh_color(vp.D, px, py, pz, cx, cy, cz, [&](int k, int c) { return f[3 * k + c]; }, rgb);
            } else {
              const float* base = shs + 3 * (size_t)vp.M * idx;
              cl = sh_color(vp.D, px, py, pz, cx, cy, cz, [&](int k, int c) { return __ldg(base + 3 * k + c); }, rgb);
            }
            g.clamped[idx] = cl;
            return cl;
          } else if (colors_mode == 1) {
            // forward.cu:241-247 / rasterizer_impl.cu:274-275: precomputed colours are blended
            // as given; staged into the same record the blend kernels read.
#pragma unroll
            for (int c = 0; c < 3; c++) rgb[c] = __ldg(colors_precomp + 3 * (size_t)idx + c);
          } else {
            rgb[0] = rgb[1] = rgb[2] = 0.0f;  // apply_weights blends no colour
          }
          return 0;
        },
        [&]() { return __ldg(opacities + idx); });
    if (o.radius > 0) {
#pragma unroll
      for (int k = 0; k < REC_F4; k++) g.rec[(size_t)idx * REC_F4 + k] = o.rec[k];
    }
    tiles = o.tiles;
    radii[idx] = o.radius;
    g.rect[idx] = o.rect;
    g.sort_key[0][idx] = o.key;
    // fit step: what the batched per-Gaussian backward needs to know about this view, one byte per Gaussian:
    // bit 0 visible, bits 1-3 the SH clamp mask (instead of per-view radii / clamped arrays)
    if (flags_out != nullptr) flags_out[idx] = o.radius > 0 ? (uint8_t)(1u | ((uint32_t)o.clamp_bits << 1)) : (uint8_t)0;
  }
  // num_rendered = sum of tiles_touched (integer, order-independent)
  uint32_t s = __reduce_add_sync(0xFFFFFFFFu, tiles);
  __shared__ uint32_t warp_sum[8];
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t += warp_sum[i];
    if (t) atomicAdd(g.counters, t);
  }
}

cudaError_t launch_preprocess(const ViewParams& vp, const float* means3D, const float* scales,
                              const float* rotations, const float* opacities, const float* shs,
                              const float* cov3D_precomp, const float* colors_precomp,
                              int colors_mode, bool prefiltered, int* radii, GeomState& g,
                              uint8_t* flags_out, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(g.counters, 0, 64 * sizeof(uint32_t), stream);
  if (e != cudaSuccess) return e;
  const int blocks = (vp.P + 255) / 256;
  preprocess_kernel<<<blocks, 256, 0, stream>>>(vp, means3D, scales, rotations, opacities, shs,
                                                cov3D_precomp, colors_precomp, colors_mode, prefiltered,
                                                radii, g, flags_out);
  DGE_LAUNCHED(1);
  return cudaGetLastError();
}

// ---- fit step: all V views of a step in ONE pass over the Gaussians -------------------------------
// The 236 B of per-Gaussian inputs (192 B of them SH) and cov3D are view-independent: a thread loads
// them once, then projects its Gaussian into every view of the step (cameras in shared memory) and
// writes the per-view records: ~20x less input traffic than V per-view launches. All inputs arrive through
// 16-byte loads: the quaternion and the 12 SH quads per thread, positions and scales per CTA (below). Also produces the
// running max of the radii over the views (max_radii2D statistics) instead of V radii arrays.
constexpr int PRE_B_THREADS = 128;
#ifndef DGE_PRE_B_MIN_CTAS
#define DGE_PRE_B_MIN_CTAS 6
#endif
constexpr int PRE_B_MIN_CTAS = DGE_PRE_B_MIN_CTAS;  // 80 registers: the 48 SH floats live in shared memory, not registers
__global__ void __launch_bounds__(PRE_B_THREADS, PRE_B_MIN_CTAS) preprocess_batched_kernel(
    ViewParams vp, int V, const float* __restrict__ cams, const float* __restrict__ means3D,
    const float* __restrict__ scales, const float* __restrict__ rotations,
    const float* __restrict__ opacities, const float* __restrict__ shs, GeomState g0, size_t geom_stride,
    uint8_t* __restrict__ flags0, size_t flags_stride, int* __restrict__ radii_max, bool prune_lists) {
  extern __shared__ float s_cam[];  // V * 40 floats, V per-view instance counts, then the SH block
  uint32_t* s_tiles = reinterpret_cast<uint32_t*>(s_cam + V * 40);  // [V]
  // this thread's 48 SH floats at s_sh[k * PRE_B_THREADS + tid]: conflict-free, and 48 registers less
  // than keeping them live across the view loop (4 -> 6 CTAs per SM)
  float* s_sh = s_cam + V * 41 + threadIdx.x;
  for (int k = threadIdx.x; k < V * 40; k += blockDim.x) s_cam[k] = cams[k];
  for (int k = threadIdx.x; k < V; k += blockDim.x) s_tiles[k] = 0;
  // The [P,3] inputs (positions, scales) are AoS with a 12-byte stride: the CTA's 128 rows of each are one
  // contiguous, 16-byte aligned block of 96 float4 — loaded as such (coalesced 16-byte loads) and handed out
  // through shared memory at a 3-word stride, which is conflict-free.
  __shared__ __align__(16) float s_in[2][3 * PRE_B_THREADS];
  {
    const size_t first = (size_t)blockIdx.x * PRE_B_THREADS;
    const int nflt = 3 * min(PRE_B_THREADS, vp.P - (int)first);
    const float* src[2] = {means3D + 3 * first, scales + 3 * first};
#pragma unroll
    for (int a = 0; a < 2; a++) {
      const bool aligned = (reinterpret_cast<uintptr_t>(src[a]) & 15) == 0;
      for (int j = threadIdx.x; j < 3 * PRE_B_THREADS / 4; j += blockDim.x) {
        if (aligned && 4 * j + 3 < nflt) {
          reinterpret_cast<float4*>(s_in[a])[j] = ldg4(src[a] + 4 * j);
        } else {
          for (int e = 0; e < 4; e++)
            if (4 * j + e < nflt) s_in[a][4 * j + e] = __ldg(src[a] + 4 * j + e);
        }
      }
    }
  }
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = idx < vp.P;
  float px = 0.f, py = 0.f, pz = 0.f, opac = 0.f;
  float cov3[6];
  if (live) {
    px = s_in[0][3 * threadIdx.x];
    py = s_in[0][3 * threadIdx.x + 1];
    pz = s_in[0][3 * threadIdx.x + 2];
    const float4 q = ldg4(rotations + 4 * (size_t)idx);
    cov3d_from_scale_rot(s_in[1][3 * threadIdx.x], s_in[1][3 * threadIdx.x + 1], s_in[1][3 * threadIdx.x + 2],
                         vp.scale_modifier, q, cov3);
    opac = __ldg(opacities + idx);
    if (shs != nullptr) {  // nullptr: no colour (mask back-projection, K13 of the reference)
      const float* base = shs + 48 * (size_t)idx;
      float4 v[12];
#pragma unroll
      for (int i = 0; i < 12; i++) v[i] = ldg4(base + 4 * i);
      const float* f = reinterpret_cast<const float*>(v);
#pragma unroll
      for (int k = 0; k < 48; k++) s_sh[k * PRE_B_THREADS] = f[k];
    }
  }
  int rmax = 0;
  for (int view = 0; view < V; view++) {
    const float* cam = s_cam + view * 40;
    uint32_t tiles = 0;
    if (live) {
      ViewParams w = vp;
      w.tan_fovx = cam[35];
      w.tan_fovy = cam[36];
      // rasterizer_impl.cu:190-191 (host float arithmetic in the reference; same operations here)
      w.focal_y = __fdiv_rn((float)vp.H, MUL(2.0f, w.tan_fovy));
      w.focal_x = __fdiv_rn((float)vp.W, MUL(2.0f, w.tan_fovx));
      PreOut o = preprocess_view(
          w, cam, cam + 16, px, py, pz, false,
          [&](float* c) {
#pragma unroll
            for (int i = 0; i < 6; i++) c[i] = cov3[i];
          },
          [&](float* rgb) -> uint8_t {
            if (shs == nullptr) {
              rgb[0] = rgb[1] = rgb[2] = 0.0f;
              return 0;
            }
            return sh_color(vp.D, px, py, pz, cam[32], cam[33], cam[34],
                            [&](int k, int c) { return s_sh[(3 * k + c) * PRE_B_THREADS]; }, rgb);
          },
          [&]() { return opac; });
      if (prune_lists && o.radius > 0) {
        // Instance-list pruning (per-step family only): the reference lists a Gaussian in every tile
        // of the square of its 3-sigma radius; only the tiles that also meet the box outside which alpha
        // < 1/255 for certain can ever blend it (the same box the blend kernels cull with), so the
        // others are not emitted — 27 % fewer instances at config 2, identical images and gradients.
        // The rect only ever SHRINKS (the 3-sigma square still clips, as in the reference); radii,
        // visibility flags and statistics are untouched.
        const float x = o.rec[0].x, y = o.rec[0].y, hx = o.rec[1].w, hy = o.rec[3].x;
        // (float -> int conversions saturate; the min() keeps the + 1 from overflowing for hx = +inf)
        const int bx0 = max(0, (int)floorf((x - hx) * 0.0625f)), bx1 = min((int)floorf((x + hx) * 0.0625f), 1 << 20) + 1;
        const int by0 = max(0, (int)floorf((y - hy) * 0.0625f)), by1 = min((int)floorf((y + hy) * 0.0625f), 1 << 20) + 1;
        const int x0 = max((int)o.rect.x, bx0), y0 = max((int)o.rect.y, by0);
        const int x1 = max(x0, min((int)o.rect.z, bx1)), y1 = max(y0, min((int)o.rect.w, by1));
        // (hx, hy = -inf for a Gaussian that can never reach 1/255: floorf(-inf) + 1 clamps to an empty rect)
        o.rect = make_ushort4(x0, y0, x1, y1);
        o.tiles = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0);
        if (o.tiles == 0) o.rect = make_ushort4(0, 0, 0, 0);
      }
      const size_t sh_ = (size_t)view * geom_stride;
      if (o.radius > 0) {
        float4* dst = shift_ptr(g0.rec, sh_) + (size_t)idx * REC_F4;
#pragma unroll
        for (int k = 0; k < REC_F4; k++) dst[k] = o.rec[k];
      }
      shift_ptr(g0.rect, sh_)[idx] = o.rect;
      shift_ptr(g0.sort_key[0], sh_)[idx] = o.key;
      // one flag byte per (view, Gaussian) for the per-Gaussian backward: bit 0 visible, bits 1-3 SH clamp mask.
      // (The view's rows of blend-stage sums are zeroed by a memset that runs beside the forward blend,
      // api.cu: written from here they were 0.96 of this kernel's 2.24 GB of stores at config 2.)
      if (flags0 != nullptr)
        flags0[(size_t)view * flags_stride + idx] = o.radius > 0 ? (uint8_t)(1u | ((uint32_t)o.clamp_bits << 1)) : (uint8_t)0;
      rmax = max(rmax, o.radius);
      tiles = o.tiles;
    }
    const uint32_t s = __reduce_add_sync(0xFFFFFFFFu, tiles);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(&s_tiles[view], s);
  }
  if (live && radii_max != nullptr) radii_max[idx] = rmax;
  __syncthreads();
  for (int k = threadIdx.x; k < V; k += blockDim.x)
    if (s_tiles[k]) atomicAdd(shift_ptr(g0.counters, (size_t)k * geom_stride), s_tiles[k]);
}

cudaError_t launch_preprocess_batched(const ViewParams& vp, const ViewBatch& vb, const float* means3D,
                                      const float* scales, const float* rotations, const float* opacities,
                                      const float* shs, GeomState& g0, uint8_t* flags, size_t flags_stride,
                                      int* radii_max, bool prune_lists, cudaStream_t stream) {
  if (shs != nullptr && vp.M != 16) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemset2DAsync(g0.counters, vb.geom_stride, 0, 64 * sizeof(uint32_t), (size_t)vb.V, stream);
  if (e != cudaSuccess) return e;
  const int blocks = (vp.P + PRE_B_THREADS - 1) / PRE_B_THREADS;
  const size_t smem = (size_t)vb.V * (40 * sizeof(float) + sizeof(uint32_t)) + 48 * PRE_B_THREADS * sizeof(float);
  preprocess_batched_kernel<<<blocks, PRE_B_THREADS, smem, stream>>>(
      vp, vb.V, vb.cams, means3D, scales, rotations, opacities, shs, g0, vb.geom_stride,
      flags, flags_stride, radii_max, prune_lists);
  DGE_LAUNCHED(1);
  return cudaGetLastError();
}

// ---- colours of a geometry-only batched preprocess -------------------------------------------------
// Multi-GPU fit: the next step's projection, depth sort and binning do not depend on the SH coefficients, so
// they run (preprocess_batched_kernel with shs == nullptr) while the all-reduce of the f_rest gradient — three
// quarters of the step's bytes — is still on the wire; once f_rest has been stepped, this kernel fills in
// what was left out: rgb of every visible (view, Gaussian) pair into its blend record and the SH clamp bits
// into its flag byte. Same sh_color as the fused path: records and flags end up bit-identical.
__global__ void __launch_bounds__(PRE_B_THREADS, PRE_B_MIN_CTAS) colour_batched_kernel(
    int P, int D, int V, const float* __restrict__ cams, const float* __restrict__ means3D,
    const float* __restrict__ shs, float4* __restrict__ rec0, size_t geom_stride, uint8_t* __restrict__ flags,
    size_t flags_stride) {
  extern __shared__ float s_cam[];  // V * 40 floats, then the SH block
  float* s_sh = s_cam + V * 40 + threadIdx.x;
  for (int k = threadIdx.x; k < V * 40; k += blockDim.x) s_cam[k] = cams[k];
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const float px = __ldg(means3D + 3 * idx), py = __ldg(means3D + 3 * idx + 1), pz = __ldg(means3D + 3 * idx + 2);
  bool loaded = false;
  // four views at a time: their flag bytes are loaded together (one coalesced byte per lane and view) before
  // any is looked at
  for (int v0 = 0; v0 < V; v0 += 4) {
    uint32_t fl[4];
#pragma unroll
    for (int k = 0; k < 4; k++) fl[k] = v0 + k < V ? (uint32_t)flags[(size_t)(v0 + k) * flags_stride + idx] : 0u;
    if (!((fl[0] | fl[1] | fl[2] | fl[3]) & 1u)) continue;
    if (!loaded) {  // Gaussians no view sees never touch their 192 bytes of SH
      const float* base = shs + 48 * (size_t)idx;
      float4 v[12];
#pragma unroll
      for (int i = 0; i < 12; i++) v[i] = ldg4(base + 4 * i);
      const float* f = reinterpret_cast<const float*>(v);
#pragma unroll
      for (int k = 0; k < 48; k++) s_sh[k * PRE_B_THREADS] = f[k];
      loaded = true;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (!(fl[k] & 1u)) continue;
      const int view = v0 + k;
      const float* cam = s_cam + view * 40;
      float rgb[3];
      const uint8_t cl = sh_color(D, px, py, pz, cam[32], cam[33], cam[34],
                                  [&](int kk, int c) { return s_sh[(3 * kk + c) * PRE_B_THREADS]; }, rgb);
      // the record's third quarter as a whole (r, g, b, view-space depth): a full 16-byte store, nothing read back
      const float depth = xform_row(cam, 2, px, py, pz);
      shift_ptr(rec0, (size_t)view * geom_stride)[(size_t)idx * REC_F4 + 2] = make_float4(rgb[0], rgb[1], rgb[2], depth);
      flags[(size_t)view * flags_stride + idx] = (uint8_t)(1u | ((uint32_t)cl << 1));
    }
  }
}

cudaError_t launch_colour_batched(int P, int D, int V, const float* cams, const float* means3D, const float* shs,
                                  GeomState& g0, size_t geom_stride, uint8_t* flags, size_t flags_stride,
                                  cudaStream_t stream) {
  const int blocks = (P + PRE_B_THREADS - 1) / PRE_B_THREADS;
  const size_t smem = (size_t)V * 40 * sizeof(float) + 48 * PRE_B_THREADS * sizeof(float);
  colour_batched_kernel<<<blocks, PRE_B_THREADS, smem, stream>>>(P, D, V, cams, means3D, shs, g0.rec, geom_stride, flags,
                                                                  flags_stride);
  DGE_LAUNCHED(1);
  return cudaGetLastError();
}

// K12 checkFrustum (DGR/cuda_rasterizer/rasterizer_impl.cu:53-63)
__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D,
                                    const float* __restrict__ view,
                                    uint8_t* __restrict__ present) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  float V[16];
#pragma unroll
  for (int i = 0; i < 16; i++) V[i] = __ldg(view + i);
  const float px = __ldg(means3D + 3 * idx), py = __ldg(means3D + 3 * idx + 1),
              pz = __ldg(means3D + 3 * idx + 2);
  present[idx] = xform_row(V, 2, px, py, pz) > 0.2f ? 1 : 0;
}

cudaError_t launch_mark_visible(int P, const float* means3D, const float* view, const float* proj,
                                uint8_t* present, cudaStream_t stream) {
  (void)proj;  // the reference's side-frustum test is commented out (auxiliary.h:154)
  mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, view, present);
  DGE_LAUNCHED(1);
  return cudaGetLastError();
}

}  // namespace dge
