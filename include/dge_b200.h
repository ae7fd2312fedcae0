/*
 * dge_b200 — C-ABI of the B200-native differentiable Gaussian-splatting
 * rasterizer that replaces DGE's `diff_gaussian_rasterization` hot path.
 *
 * Every entry point below is the drop-in for one function of the reference's
 * torch-free CUDA library (`CudaRasterizer::Rasterizer`,
 * gaussiansplatting/submodules/diff-gaussian-rasterization/cuda_rasterizer/
 * rasterizer.h, abbreviated DGR/ below) or of its torch glue
 * (DGR/rasterize_points.cu). Plain pointers and sizes only: all pointers are
 * DEVICE pointers unless the name says `host`; the library allocates nothing
 * persistent — scratch comes from the caller through the three allocator
 * callbacks exactly like the reference's std::function<char*(size_t)> trio.
 *
 * Return value: >= 0 on success (dge_rasterize_forward returns num_rendered),
 * < 0 on failure; dge_last_error() then returns a thread-local message.
 * Streams are passed as void* (a cudaStream_t); NULL is the legacy stream.
 */
#ifndef DGE_B200_H_
#define DGE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Scratch allocator: must return a device pointer to at least `bytes` bytes,
 * 256-byte aligned, valid until the matching backward has run.
 * Replaces std::function<char*(size_t)> (DGR/cuda_rasterizer/rasterizer.h:32-34,
 * bound to torch tensors by resizeFunctional, DGR/rasterize_points.cu:27-33). */
typedef char* (*dge_alloc_fn)(void* ctx, size_t bytes);

const char* dge_last_error(void);
/* ABI version of this header; bumped on any signature change. */
int dge_abi_version(void);

/* Replaces CudaRasterizer::Rasterizer::forward (DGR/cuda_rasterizer/rasterizer.h:31-56,
 * rasterizer_impl.cu:179-285). Same argument meaning and order; additions:
 * `alloc_ctx` (passed back to the three callbacks) and `stream`.
 * NULL / empty semantics as the reference: shs==NULL <=> colors_precomp given,
 * scales/rotations==NULL <=> cov3D_precomp given.
 * out_color [3,H,W], out_depth [1,H,W], radii [P] (int32) are fully written. */
int dge_rasterize_forward(dge_alloc_fn geometryBuffer, dge_alloc_fn binningBuffer,
                          dge_alloc_fn imageBuffer, void* alloc_ctx,
                          int P, int D, int M, const float* background,
                          int width, int height, const float* means3D,
                          const float* shs, const float* colors_precomp,
                          const float* opacities, const float* scales,
                          float scale_modifier, const float* rotations,
                          const float* cov3D_precomp, const float* viewmatrix,
                          const float* projmatrix, const float* cam_pos,
                          float tan_fovx, float tan_fovy, int prefiltered,
                          float* out_color, float* out_depth, int* radii,
                          int debug, void* stream);

/* Replaces CudaRasterizer::Rasterizer::backward (DGR/cuda_rasterizer/rasterizer.h:58-87,
 * rasterizer_impl.cu:289-341) TOGETHER WITH the nine torch::zeros fills of
 * RasterizeGaussiansBackwardCUDA (DGR/rasterize_points.cu:120-128): every output
 * below is fully written by the call (zeros for culled Gaussians), so the
 * caller passes uninitialised memory. dL_dconic [P,4] and dL_dcolor [P,3] are
 * the reference's intermediates; any output pointer except dL_dmean2D,
 * dL_dmean3D may be NULL when the caller does not need it.
 * dL_dmean2D is [P,3] with .z = 0 (DGR/rasterize_points.cu:121).
 * `scratchBuffer` provides dge_backward_scratch_bytes(P) bytes that only need to
 * live until the call's work on `stream` has finished (the blend-stage sums).
 * accumulate != 0: the outputs are running sums — rows of visible Gaussians are ADDED to,
 * rows of culled Gaussians are left alone (what autograd's AccumulateGrad does with the
 * reference's per-view tensors, DGE.py:170-239, without the per-view tensors). */
int dge_rasterize_backward(dge_alloc_fn scratchBuffer, void* alloc_ctx,
                           int P, int D, int M, int R, const float* background,
                           int width, int height, const float* means3D,
                           const float* shs, const float* colors_precomp,
                           const float* scales, float scale_modifier,
                           const float* rotations, const float* cov3D_precomp,
                           const float* viewmatrix, const float* projmatrix,
                           const float* campos, float tan_fovx, float tan_fovy,
                           const int* radii, char* geom_buffer,
                           char* binning_buffer, char* image_buffer,
                           const float* dL_dpix, float* dL_dmean2D,
                           float* dL_dconic, float* dL_dopacity,
                           float* dL_dcolor, float* dL_dmean3D,
                           float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                           float* dL_drot, int accumulate, int debug, void* stream);

/* Replaces CudaRasterizer::Rasterizer::apply_weights (DGR/cuda_rasterizer/rasterizer.h:89-112,
 * rasterizer_impl.cu:343-447): DGE's mask back-projection. `weights` [P,CH] f32
 * and `cnt` [P] i32 are accumulated IN PLACE; image_weights is [CH,H,W],
 * CH = num_channels in {1,2,3}. */
int dge_apply_weights(dge_alloc_fn geometryBuffer, dge_alloc_fn binningBuffer,
                      dge_alloc_fn imageBuffer, void* alloc_ctx, int P, int D,
                      int M, const float* background, int width, int height,
                      const float* means3D, const float* shs, float* weights,
                      const float* opacities, const float* scales,
                      float scale_modifier, const float* rotations,
                      const float* cov3D_precomp, const float* viewmatrix,
                      const float* projmatrix, const float* cam_pos,
                      float tan_fovx, float tan_fovy, int prefiltered,
                      const float* image_weights, int* radii, int* cnt,
                      int num_channels, int debug, void* stream);

/* Replaces CudaRasterizer::Rasterizer::markVisible (DGR/cuda_rasterizer/rasterizer.h:24-29,
 * rasterizer_impl.cu:128-133). present is uint8[P] (bool). */
int dge_mark_visible(int P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, uint8_t* present, void* stream);

/* Scratch sizes (the reference's required<GeometryState/BinningState/ImageState>,
 * DGR/cuda_rasterizer/rasterizer_impl.h:66-72). The layouts are ours. */
size_t dge_geom_bytes(int P);
size_t dge_binning_bytes(int R, int width, int height);
size_t dge_image_bytes(int width, int height);
size_t dge_backward_scratch_bytes(int P);

/* Inspection of the opaque scratch blobs — used by the parity tests to compare
 * every intermediate with the reference's (DGR/cuda_rasterizer/rasterizer_impl.cu:135-175).
 * geom: out[0]=records float4[P][4]: one 64-byte blend record per Gaussian — (x, y, conic.x, conic.y |
 *       conic.z, power threshold, opacity, hx | r, g, b, view-space depth | hy, cull constants x3);
 *       out[1]=out[2]=NULL (reserved); out[3]=rect
 *       ushort4[P] (min.x,min.y,max.x,max.y), out[4]=clamped uint8[P] (bit ch),
 *       out[5]=depth_order uint32[P] (Gaussian ids sorted by depth bits),
 *       out[6]=point_offsets uint32[P] (inclusive, in depth order),
 *       out[7]=num_rendered uint32[1]
 * binning: out[0]=point_list uint32[R], out[1]=tile ids (sorted) uint32[R]
 * image: out[0]=final_T float[N], out[1]=n_contrib uint32[N], out[2]=ranges uint2[T] */
void dge_geom_pointers(char* chunk, int P, void** out);
void dge_binning_pointers(char* chunk, int R, int width, int height, void** out);
void dge_image_pointers(char* chunk, int width, int height, void** out);

/* Rebuilds the reference's sorted 64-bit key list ((tile<<32)|depth bits,
 * DGR/cuda_rasterizer/rasterizer_impl.cu:88-93) from our binning state, for
 * bit-exact comparison. keys_out is uint64[R]. */
int dge_debug_sorted_keys(char* geom_buffer, char* binning_buffer, int P, int R,
                          int width, int height, uint64_t* keys_out,
                          void* stream);

/* ---- measurement hooks (bench.py) ----
 * Kernels launched by this library so far (all entry points, this process). */
unsigned long long dge_launch_count(void);
/* Per-stage CUDA-event timing on the launching stream. Bit s of stage_mask enables stage s:
 * 0 preprocess, 1 depth sort, 2 binning (scan+expand+tile sort+ranges), 3 blend forward,
 * 4 blend backward, 5 per-Gaussian backward, 6 apply_weights blend. */
#define DGE_NUM_STAGES 7
/* out: device, 2*256 uint64, zero-filled by the caller. Every SM that runs a block of the probe stores
 * (clock64, globaltimer ns) at out[2*smid]. Two probes on one stream, before and after a timed region,
 * give the average SM clock of that region (delta cycles / delta ns) with no driver query in between. */
int dge_clock_probe(unsigned long long* out, void* stream);
void dge_profile_enable(unsigned stage_mask);
int dge_profile_read(float* ms_out, int* count_out);

/* ---- fit step: the views of one optimisation step --------------------
 * `cam` is a device record of 40 floats: viewmatrix[16] | projmatrix[16] | campos[3] | tan_fovx |
 * tan_fovy | pad[3]. `acc` is this view's [P][12] row block of blend-stage sums (moments of dL/dG * G over the
 * Gaussian's pixels, dL/dopacity, dL/dcolour: ACC_* in csrc/common.cuh), `flags` its [P] bytes (bit 0: the
 * Gaussian is visible in the view, bits 1-3: SH colour channel clamped at 0).
 *
 * dge_fit_forward = dge_rasterize_forward (SH colours, scale/rotation covariances) that also
 * zeroes `acc` and writes `flags`.
 * dge_fit_backward_blend = the blend backward of the view into `acc` (K7 only);
 * background_is_black != 0 is the caller's promise that background == (0,0,0) (DGE.py:87), which
 * removes the background term of dL/dalpha at compile time.
 * dge_fit_backward_geom = K8 + K9 of rasterizer_impl.cu:324-340 for ALL V views of the step in one
 * pass over the Gaussians (acc of view v at acc + v*acc_stride_floats, its flags at flags + v*flags_stride);
 * with accumulate != 0 the six
 * outputs are added to, otherwise every row is written. */
int dge_fit_forward(dge_alloc_fn geometryBuffer, dge_alloc_fn binningBuffer, dge_alloc_fn imageBuffer,
                    void* alloc_ctx, int P, int D, int M, const float* background, int width,
                    int height, const float* means3D, const float* shs, const float* opacities,
                    const float* scales, float scale_modifier, const float* rotations,
                    const float* cam, float tan_fovx, float tan_fovy, float* out_color,
                    float* out_depth, int* radii, float* acc, uint8_t* flags, void* stream);
int dge_fit_backward_blend(int P, int R, const float* background, int background_is_black,
                           int width, int height, char* geom_buffer, char* binning_buffer,
                           char* image_buffer, const float* dL_dpix, float* acc, void* stream);
int dge_fit_backward_geom(int P, int D, int M, int V, const float* cams, int width, int height,
                          float scale_modifier, const float* acc, size_t acc_stride_floats,
                          const uint8_t* flags, size_t flags_stride, const float* means3D, const float* shs, const float* scales,
                          const float* rotations, float* dL_dmean3D, float* dL_dmean2D,
                          float* dL_dsh, float* dL_dopacity, float* dL_dscale, float* dL_drot,
                          int accumulate, void* stream);

/* The same three stages for ALL V views of the step at once (1 <= V <= 64): every stage is ONE launch
 * whose grid carries the view as a dimension, so a 512x512 view's 1024 tiles no longer leave most of
 * the 148 SMs idle, the 236 B of per-Gaussian inputs are read once per step instead of once per
 * view, and the host waits for the instance counts once per step. `cams` is [V][40], out_color
 * [V,3,H,W], out_depth [V,1,H,W], acc [V][acc_stride_floats] (acc_stride_floats >= 12*P, multiple
 * of 4; zeroed by a memset that the library runs beside the forward blend and that
 * dge_fit_views_backward_blend of the same acc waits for), flags [V][flags_stride] (flags_stride >= P),
 * radii_max [P] = max over the views of the reference's per-view radii. geometryBuffer is
 * asked for V*dge_geom_bytes(P) bytes, imageBuffer for V*dge_image_bytes(W,H), binningBuffer for
 * dge_fit_binning_bytes(R_total, V, W, H) once the counts are known. num_rendered_host (host, [V], may be
 * NULL) receives the per-view num_rendered. Returns R_total = their sum.
 * acc + flags (both or neither) and radii_max may be NULL (forward-only rendering, e.g. DGE's render_all_view).
 * extra [P] / out_extra [V,3,H,W] (both or neither): one more per-Gaussian scalar blended like a colour
 * channel, background added per channel — bit-identical to a second forward of the same view with
 * colors_precomp = extra repeated three times, which is how DGE.forward renders its "semantic" map of
 * the edit mask for every view of every step (threestudio/systems/DGE.py:198-204).
 * prune_lists == 0: the instance lists and tile ranges of every view are bit-identical to the
 * reference's (and to dge_rasterize_forward's). prune_lists != 0: a Gaussian is only listed in the tiles
 * of its 3-sigma square that also meet the conservative box outside which its alpha is < 1/255 — the
 * instances dropped can never blend, so images, depth, gradients, radii and back-projected weights are
 * identical, with ~27 % fewer instances to emit, partition and stage. */
int dge_fit_views_forward(dge_alloc_fn geometryBuffer, dge_alloc_fn binningBuffer, dge_alloc_fn imageBuffer,
                          void* alloc_ctx, int P, int D, int M, int V, const float* background, int width,
                          int height, const float* means3D, const float* shs, const float* opacities,
                          const float* scales, float scale_modifier, const float* rotations,
                          const float* cams, float* out_color, float* out_depth, int* radii_max, float* acc,
                          size_t acc_stride_floats, uint8_t* flags, size_t flags_stride, int* num_rendered_host,
                          const float* extra, float* out_extra, int prune_lists, void* stream);
/* The same call in three pieces, for callers that want to start a step's projection, depth sort and binning before
 * its SH coefficients are final (multi-GPU fit: they run under the all-reduce of the f_rest gradient, fit.py):
 *   dge_fit_views_front  = batched preprocess + segment offsets + depth sort + binning; shs == NULL leaves the
 *                          colours out of the blend records (flags then carry the visibility bit only). Returns
 *                          R_total.
 *   dge_fit_views_colour = SH -> rgb + clamp bits of every visible (view, Gaussian) pair into the records and
 *                          flags of a front half that ran with shs == NULL (bit-identical to the fused path).
 *   dge_fit_views_blend  = the forward blend of a front half (and the zeroing of acc beside it).
 * dge_fit_views_forward == dge_fit_views_front (with shs) + dge_fit_views_blend. */
int dge_fit_views_front(dge_alloc_fn geometryBuffer, dge_alloc_fn binningBuffer, dge_alloc_fn imageBuffer,
                        void* alloc_ctx, int P, int D, int M, int V, int width, int height, const float* means3D,
                        const float* shs, const float* opacities, const float* scales, float scale_modifier,
                        const float* rotations, const float* cams, int* radii_max, uint8_t* flags,
                        size_t flags_stride, int* num_rendered_host, int prune_lists, void* stream);
int dge_fit_views_colour(int P, int D, int M, int V, const float* means3D, const float* shs, const float* cams,
                         char* geom_buffer, uint8_t* flags, size_t flags_stride, void* stream);
int dge_fit_views_blend(int P, int V, int R_total, const float* background, int width, int height, char* geom_buffer,
                        char* binning_buffer, char* image_buffer, float* out_color, float* out_depth, float* acc,
                        size_t acc_stride_floats, const float* extra, float* out_extra, void* stream);
/* dL_dpix is [V,3,H,W]; the three blobs are the ones dge_fit_views_forward filled. */
int dge_fit_views_backward_blend(int P, int V, int R_total, const float* background, int background_is_black,
                                 int width, int height, char* geom_buffer, char* binning_buffer,
                                 char* image_buffer, const float* dL_dpix, float* acc, size_t acc_stride_floats,
                                 void* stream);
size_t dge_fit_binning_bytes(int R_total, int V, int width, int height);
/* DGE.update_mask (threestudio/systems/DGE.py:101-165) in one call: dge_apply_weights for V views at once,
 * view v with its own mask image image_weights[v] ([V,CH,H,W]) and camera cams[v], all accumulating
 * into the same weights [P,CH] / cnt [P] (in place, as the reference's per-view calls do). Scratch as
 * for dge_fit_views_forward. Returns the total number of instances. */
int dge_fit_views_apply_weights(dge_alloc_fn geometryBuffer, dge_alloc_fn binningBuffer, dge_alloc_fn imageBuffer,
                                void* alloc_ctx, int P, int V, int width, int height, const float* means3D,
                                const float* opacities, const float* scales, float scale_modifier,
                                const float* rotations, const float* cams, const float* image_weights,
                                int num_channels, float* weights, int* cnt, int* num_rendered_host,
                                int prune_lists, void* stream);

/* GaussianModel's activations (gaussiansplatting/scene/gaussian_model.py:221-258)
 * for the whole model in one pass — shs[P,16,3] = cat(f_dc[P,1,3], f_rest[P,15,3]), opacities =
 * sigmoid, scales = exp, rotations = normalize — and dge_fit_backward_geom with their backward in
 * its epilogue: the seven outputs are gradients w.r.t. the RAW parameters, every row written. */
/* (dge_fit_activate: f_dc == f_rest == NULL skips the features, opacity_raw == NULL the other three.) */
int dge_fit_activate(int P, const float* f_dc, const float* f_rest, const float* opacity_raw,
                     const float* scaling_raw, const float* rotation_raw, float* shs,
                     float* opacities, float* scales, float* rotations, void* stream);
int dge_fit_backward_geom_raw(int P, int D, int V, const float* cams, int width, int height,
                              float scale_modifier, const float* acc, size_t acc_stride_floats,
                              const uint8_t* flags, size_t flags_stride, const float* means3D, const float* shs, const float* opacities,
                              const float* scales, const float* rotations,
                              const float* rotation_raw, float* d_xyz, float* d_means2D,
                              float* d_f_dc, float* d_f_rest, float* d_opacity_raw,
                              float* d_scaling_raw, float* d_rotation_raw, void* stream);

/* ---- fit-step helpers ----
 * L1 loss of one rendered view and its gradient (threestudio/systems/DGE.py:672):
 * grad[i] = scale * sign(image[i] - target[i]); *loss_accum += scale * sum |image - target|. */
int dge_l1_loss_grad(const float* image, const float* target, size_t n, float scale,
                     float* grad, float* loss_accum, void* stream);
/* Densification statistics of one step (threestudio/systems/DGE.py:266-284, GaussianModel.add_densification_stats
 * gaussiansplatting/scene/gaussian_model.py:811-815) in one pass over the model: where radii_max[i] > 0 (some view of
 * the step saw Gaussian i; radii_max = max over the step's views, after the MAX all-reduce on several GPUs):
 * max_radii2D[i] = max(max_radii2D[i], radii_max[i]); xyz_gradient_accum[i] += |means2D_grad[i].xy| (the summed
 * screen-space gradient, [P,3] floats); denom[i] += 1. In place. */
int dge_fit_update_stats(int P, const int* radii_max, const float* means2D_grad, int* max_radii2D,
                         float* xyz_gradient_accum, float* denom, void* stream);
/*
 * Fused Adam over one flat fp32 parameter block (torch.optim.Adam semantics,
 * gaussiansplatting/scene/gaussian_model.py:374: eps=1e-15, no weight decay,
 * no amsgrad), with the optional per-Gaussian grad mask of
 * gaussian_model.py:837-856 (mask uint8[n/stride] or NULL). In place. */
int dge_fused_adam(float* param, const float* grad, float* exp_avg,
                   float* exp_avg_sq, size_t n, float lr, float beta1,
                   float beta2, float eps, int step, const uint8_t* mask,
                   int stride, void* stream);

/* ---- densification on the flat fit buffers ----
 * GaussianModel.densify_and_prune (gaussiansplatting/scene/gaussian_model.py:543-807) as two passes.
 * dge_densify_select: per Gaussian the clone / split / prune decisions from its (masked, thresholded) accumulated
 *   gradient `grad` [P], raw scaling [P,3], raw opacity [P] and the edit mask (uint8 [P] or NULL): clone if
 *   |grad| >= max_grad and max(scale) <= size_threshold (= percent_dense * extent), split if grad >= max_grad and
 *   max(scale) > size_threshold; prune (inside the mask) if sigmoid(opacity) < min_opacity or, with prune_size >= 0
 *   (= 0.1 * extent), max(scale) > prune_size — the children with their scale / (0.8 N). keep [4][P] int32: the
 *   original stays | its clone stays | its children stay | split-selected; sel [P]: bit 0 cloned, bit 1 split.
 * dge_densify_gather: with scan = exclusive scans of keep's four rows and their totals K_*, writes the new
 *   parameter / exp_avg / exp_avg_sq blocks of the six groups (src / dst: [3 buffers][6 groups] pointers, widths[6]
 *   floats per Gaussian; g_* = indices of the xyz, scaling, rotation groups) in the reference's row order
 *   [kept originals | kept clones | children copy 0 .. N-1], the children at R(q) sample + xyz with
 *   samples [N][K_split][3] (the draw of :685-687) and log(scale / (0.8 N)), new rows with zero Adam moments
 *   (cat_tensors_to_optimizer :609-640), and the new edit mask. */
int dge_densify_select(int P, const float* grad, const float* scaling_raw, const float* opacity_raw,
                       const uint8_t* mask, float max_grad, float size_threshold, float min_opacity,
                       float prune_size, int N, int* keep, uint8_t* sel, void* stream);
int dge_densify_gather(int P, int N, const int* keep, const int* scan, int K_orig, int K_clone, int K_child,
                       int K_split, const float* const* src, float* const* dst, const int* widths, int g_xyz,
                       int g_scaling, int g_rotation, const float* samples, const uint8_t* mask_in,
                       uint8_t* mask_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DGE_B200_H_ */
