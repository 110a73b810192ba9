/* gags_b200.h — C-ABI of the B200-native GAGS feature rasteriser.
 *
 * Drop-in boundary.  The reference reaches all of this arithmetic through one
 * Python call, gsplat.rasterization(...), at /root/reference/gaussian_renderer/__init__.py:56-70
 * (gsplat is an external pip dependency, /root/reference/environment.yml:26).  Each entry point
 * below replaces one CUDA op that call dispatches to; the op it replaces is named in the comment.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless marked "host";
 *   - the caller allocates and owns every buffer (incl. workspaces); nothing is retained;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no host sync inside
 *     unless stated;
 *   - return 0 on success, a positive cudaError_t, or a negative GAGS_E* code; never throws;
 *   - float = IEEE fp32, row-major, densely packed unless a stride is given.
 */
#ifndef GAGS_B200_H
#define GAGS_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GAGS_TILE 16

#define GAGS_EINVAL  (-1)   /* bad argument (null pointer, bad size, unsupported D)        */
#define GAGS_EALIGN  (-2)   /* pointer/stride not aligned as required (16 B for rows)       */
#define GAGS_ESMALL  (-3)   /* workspace too small                                         */
#define GAGS_ERANGE  (-4)   /* value out of supported range (e.g. tile_bits + 32 > 64)      */

/* flags for gags_camera_t.flags */
#define GAGS_F_LOG_SCALES      1  /* scales are log-scales: apply exp() then * scaling_modifier
                                     (GaussianModel.get_scaling, scene/gaussian_model.py:116-118) */
#define GAGS_F_LOGIT_OPACITY   2  /* opacities are logits: apply sigmoid()
                                     (GaussianModel.get_opacity, scene/gaussian_model.py:133-134) */

/* Host struct describing one pinhole view: what render() derives from a Camera at
 * gaussian_renderer/__init__.py:27-38 (K from FoV) and :55 (viewmat = world_view_transform^T). */
typedef struct gags_camera {
  float viewmat[16];       /* row-major 4x4 world -> camera (host copy)                  */
  const float *viewmat_dev;/* optional DEVICE pointer to the same 16 floats; when non-NULL the
                              kernels read it instead of viewmat[] (no host round trip)    */
  float fx, fy, cx, cy;    /* intrinsics in pixels                                       */
  int32_t width, height;   /* image size in pixels                                       */
  float eps2d;             /* 0.3  : added to the 2-D covariance diagonal                */
  float near_plane;        /* 0.01                                                      */
  float far_plane;         /* 1e10                                                      */
  float radius_clip;       /* 0                                                         */
  float scaling_modifier;  /* render(..., scaling_modifier) :42                           */
  int32_t flags;           /* GAGS_F_*                                                   */
} gags_camera_t;

/* library / build identification (static strings) */
const char *gags_version(void);
const char *gags_build_arch(void);           /* "sm_100a" */
const char *gags_error_string(int code);

/* Implementation selector for the wide (D > 32) blend kernels: 0 = auto (tcgen05 tensor-core path
 * whenever D % 16 == 0, else SIMT), 1 = SIMT only, 2 = tensor-core required.  Process-wide. */
int gags_set_blend_impl(int32_t impl);
int gags_get_blend_impl(void);

/* ---------------------------------------------------------------------------------------------
 * K1  projection + frustum cull  (replaces gsplat fully_fused_projection_fwd)
 * In : means[N,3] quats[N,4] (w,x,y,z; normalised in-kernel) scales[N,3] opacities[N]
 * Out: radii[N] (0 = culled) means2d[N,2] depths[N] conics[N,3]
 *      opac_out[N]      activated opacity (may be NULL)
 *      tiles_touched[N] number of 16x16 tiles overlapped = pass 1 of isect_tiles (may be NULL)
 *      geom[N,8]        packed blend record {mx,my,a,b,c,opacity,depth,radius} (may be NULL)
 * tile_w/tile_h are only used for tiles_touched.
 */
int gags_project_fwd(const float *means, const float *quats, const float *scales,
                     const float *opacities, int64_t N, const gags_camera_t *cam /*host*/,
                     int32_t tile_w, int32_t tile_h, int32_t *radii, float *means2d,
                     float *depths, float *conics, float *opac_out, int32_t *tiles_touched,
                     float *geom, void *stream);

/* K2  VJP of K1 (replaces fully_fused_projection_bwd; App. A.7), chained through the fused
 * activations selected by cam->flags.  Gradient outputs may be NULL individually.
 * In : forward inputs + radii + conics, v_means2d[N,2] v_depths[N] (NULL = 0) v_conics[N,3]
 * Out: v_means[N,3] v_quats[N,4] v_scales[N,3]   (w.r.t. the RAW inputs given to the forward) */
int gags_project_bwd(const float *means, const float *quats, const float *scales, int64_t N,
                     const gags_camera_t *cam /*host*/, const int32_t *radii,
                     const float *conics, const float *v_means2d, const float *v_depths,
                     const float *v_conics, float *v_means, float *v_quats, float *v_scales,
                     void *stream);

/* sigmoid VJP for the fused opacity activation: v_logit = v_opac * o * (1 - o). */
int gags_opacity_bwd(const float *opac_act, const float *v_opac, int64_t N, float *v_logit,
                     void *stream);

/* K3  spherical harmonics -> RGB (replaces gsplat spherical_harmonics fwd; App. A.2; basis ==
 * /root/reference/utils/sh_utils.py:57-112).  colors = max(SH(dir) + 0.5, 0) for radii > 0.
 * In : means[N,3], campos[3] (DEVICE), coeffs[N,K,3] (K >= (deg+1)^2), radii[N] (NULL = all)
 * Out: colors[N, out_stride] first 3 entries of each row written (out_stride >= 3)           */
int gags_sh_fwd(int32_t degree, const float *means, const float *campos /*device*/,
                const float *coeffs, int32_t K, const int32_t *radii, int64_t N, float *colors,
                int32_t out_stride, void *stream);

/* K3 bwd: v_coeffs[N,K,3] (rows of culled Gaussians and bases above `degree` are zeroed),
 * v_means[N,3] (+= gradient through the view direction; may be NULL).                         */
int gags_sh_bwd(int32_t degree, const float *means, const float *campos /*device*/,
                const float *coeffs, int32_t K, const int32_t *radii, int64_t N,
                const float *v_colors, int32_t v_stride, float *v_coeffs, float *v_means,
                void *stream);

/* ---------------------------------------------------------------------------------------------
 * K4  tile intersection (replaces gsplat isect_tiles; App. A.3) — integer stage, exact.
 * gags_tile_count : tiles_touched[N] from (means2d, radii)            (pass 1)
 * gags_tile_scan  : cum_tiles[N] = inclusive prefix sum (int32), *n_isects_dev = total
 * gags_tile_emit  : isect_ids[n] = (tile << 32) | depth_bits, flatten_ids[n] = gaussian index,
 *                   emitted in ascending Gaussian index, row-major tiles   (pass 2)
 */
int gags_tile_count(const float *means2d, const int32_t *radii, int64_t N, int32_t tile_w,
                    int32_t tile_h, int32_t *tiles_touched, void *stream);
size_t gags_tile_scan_workspace_bytes(int64_t N);
int gags_tile_scan(const int32_t *tiles_touched, int64_t N, int32_t *cum_tiles,
                   int32_t *n_isects_dev, void *workspace, size_t workspace_bytes, void *stream);
int gags_tile_emit(const float *means2d, const int32_t *radii, const float *depths,
                   const int32_t *cum_tiles, int64_t N, int32_t tile_w, int32_t tile_h,
                   int64_t *isect_ids, int32_t *flatten_ids, void *stream);

/* K4-K6 as a tile-bucketed segmented sort (csrc/tile_buckets.cu): same isect_ids / flatten_ids /
 * offsets as the count-scan-emit-sort-offsets sequence above, bit for bit, with about a sixth of
 * its HBM traffic.  gags_tile_bucket_count: scatter every (Gaussian, tile) pair into the tile's slab
 * of bucket[n_tiles * gags_tile_bucket_max()] (8 B entries; count[n_tiles] scratch) + exclusive
 * scan -> offsets[n_tiles+1] (= isect_offsets), stats_dev = {n_isects, largest bucket}.
 * gags_tile_bucket_sort: sort every bucket by (depth bits, Gaussian index) in shared memory.
 * Returns GAGS_ERANGE when max_bucket > gags_tile_bucket_max(): use the global path then.
 * isect_ids may be NULL.                                                                        */
int32_t gags_tile_bucket_max(void);
int gags_tile_bucket_count(const float *means2d, const int32_t *radii, const float *depths,
                           int64_t N, int32_t tile_w, int32_t tile_h, int32_t *count, void *bucket,
                           int32_t *offsets, int32_t *stats_dev, void *stream);
int gags_tile_bucket_sort(const void *bucket, int32_t tile_w, int32_t tile_h, const int32_t *offsets,
                          int32_t max_bucket, int64_t *isect_ids, int32_t *flatten_ids,
                          void *stream);
/* gags_tile_bucket_sort launched before the host has read stats_dev (the device sorts while the
 * host waits for the counts).  capacity = elements isect_ids / flatten_ids hold; tiles that do not
 * fit are skipped on the device and the caller discards the result when stats_dev says so.      */
int gags_tile_bucket_sort_guarded(const void *bucket, int32_t tile_w, int32_t tile_h,
                                  const int32_t *offsets, int64_t capacity, int64_t *isect_ids,
                                  int32_t *flatten_ids, void *stream);

/* K5  stable LSD radix sort of (int64 key, int32 value) on bits [0, end_bit)
 * (replaces gsplat's cub::DeviceRadixSort::SortPairs call).  Ping-pong buffers: the result is in
 * (keys_a, vals_a) if *selector_host == 0 else (keys_b, vals_b).                              */
size_t gags_sort_pairs_workspace_bytes(int64_t n);
int gags_sort_pairs(int64_t *keys_a, int64_t *keys_b, int32_t *vals_a, int32_t *vals_b, int64_t n,
                    int32_t end_bit, void *workspace, size_t workspace_bytes,
                    int32_t *selector_host /*host*/, void *stream);

/* K6  tile offsets (replaces gsplat isect_offset_encode; App. A.4).
 * offsets[n_tiles + 1]: offsets[t] = first sorted index with tile >= t; offsets[n_tiles] = n.  */
int gags_tile_offsets(const int64_t *isect_ids_sorted, int64_t n, int32_t n_tiles,
                      int32_t *offsets, void *stream);

/* ---------------------------------------------------------------------------------------------
 * K7  alpha-blend forward (replaces gsplat rasterize_to_pixels_fwd and the channel_chunk loop
 * + torch.cat around it; App. A.5).  One launch for any D (D % 4 == 0, rows 16-B aligned).
 * In : geom[N,8] (from gags_project_fwd), colors[N,D] (row stride = D), background[D] or NULL,
 *      offsets[n_tiles+1], flatten_ids[n_isects]
 * Out: render[H,W,D] alphas[H,W] last_ids[H,W] (index into flatten_ids of the last contributor)
 */
int gags_blend_fwd(const float *geom, const float *colors, int32_t D, const float *background,
                   int32_t width, int32_t height, const int32_t *offsets,
                   const int32_t *flatten_ids, float *render, float *alphas, int32_t *last_ids,
                   void *stream);

/* K8a feature-only backward (frozen geometry; the only gradient train.py consumes,
 * /root/reference/scene/gaussian_model.py:192-206):  v_colors[g,:] += sum_px w(g,px) v_render[px,:]
 * with w identical to the forward weight.  v_colors[N,D] must be zero-initialised by the caller. */
int gags_blend_bwd_features(const float *geom, int32_t D, int32_t width, int32_t height,
                            const int32_t *offsets, const int32_t *flatten_ids,
                            const float *v_render, float *v_colors, void *stream);

/* K7 + K8a with a weight-tile cache (training with frozen geometry, D > 32, D % 16 == 0).
 * gags_blend_fwd_cached is gags_blend_fwd that additionally saves, for every batch of 32 Gaussians
 * it blends into a 16x8 half tile, the 128 x 32 blend-weight tile (bf16 hi/lo, 16 KB) and the 32
 * Gaussian ids; gags_blend_bwd_features_cached then computes v_colors as a streaming tensor-core
 * GEMM over those tiles instead of re-walking the tile lists (same result as
 * gags_blend_bwd_features).  Caller-owned buffers, `slots = gags_blend_cache_slots(n_isects,
 * n_tiles)`:  wcache[slots * 16384] bytes (16-B aligned), wmeta[slots * 32], wlist[slots],
 * wcount[tile_w * ceil(height / 8) + 1] (the extra int is the backward's job counter).  None needs
 * initialising.                                                                                */
int gags_blend_cache_supported(int32_t D);
/* 1 when gags_blend_fwd / gags_blend_fwd_cached accept last_ids == NULL for this D (the wide
 * tensor-core kernel then skips the last-contributor tracking and the per-step selects of the
 * transmittance chain; render_alphas = sum of the blend weights instead of 1 - T, equal to a few
 * 1e-7).  last_ids is only read by gags_blend_bwd_full.                                           */
int gags_blend_last_ids_optional(int32_t D);     /* 1 if the cached pair handles this D          */
int64_t gags_blend_cache_slots(int64_t n_isects, int32_t n_tiles);
int gags_blend_fwd_cached(const float *geom, const float *colors, int32_t D,
                          const float *background, int32_t width, int32_t height,
                          const int32_t *offsets, const int32_t *flatten_ids, float *render,
                          float *alphas, int32_t *last_ids, void *wcache, int32_t *wmeta,
                          int32_t *wlist, int32_t *wcount, void *stream);
/* The cached forward in two passes, split at the weight-tile cache (same buffers as
 * gags_blend_fwd_cached; together they produce exactly its outputs):
 *   gags_blend_fwd_weights    — the geometry-only half (tile walk, cull, alpha, transmittance):
 *                               alphas, last_ids (may be NULL), weight tiles + batch lists; reads no
 *                               feature.  After it, gags_blend_cache_mark_rows names the feature rows
 *                               the view will read.
 *   gags_blend_fwd_from_cache — render = cached weights x colors (+ (1 - alpha) * background).      */
int gags_blend_fwd_weights(const float *geom, int32_t width, int32_t height, const int32_t *offsets,
                           const int32_t *flatten_ids, float *alphas, int32_t *last_ids,
                           void *wcache, int32_t *wmeta, int32_t *wlist, int32_t *wcount,
                           void *stream);
int gags_blend_fwd_from_cache(const float *colors, int32_t D, const float *background, int32_t width,
                              int32_t height, const int32_t *offsets, const void *wcache,
                              const int32_t *wmeta, const int32_t *wlist, const int32_t *wcount,
                              const float *alphas, float *render, void *stream);
int gags_blend_bwd_features_cached(int32_t D, int32_t width, int32_t height,
                                   const int32_t *offsets, const void *wcache,
                                   const int32_t *wmeta, const int32_t *wlist,
                                   int32_t *wcount, const float *v_render, float *v_colors,
                                   void *stream);
/* The same backward with the masked L1 loss of train.py:162-163 fused into it: takes the RENDER
 * instead of v_render, forms v_render = grad_scale * m * sign(render - emb[seg]) on the fly (the
 * semantics of gags_l1_loss_segmap below; seg < 0 or >= n_seg = pixel without a target; mask may be
 * NULL) and adds sum m |render - emb[seg]| to *loss_out (caller zeroes it and divides by H W D).
 * Replaces the pair gags_l1_loss_segmap + gags_blend_bwd_features_cached: the [H,W,D] gradient map
 * is never written to or read from HBM.                                                         */
int gags_blend_bwd_features_cached_l1(int32_t D, int32_t width, int32_t height,
                                      const int32_t *offsets, const void *wcache,
                                      const int32_t *wmeta, const int32_t *wlist, int32_t *wcount,
                                      const float *render, const int32_t *seg, const float *emb,
                                      const float *mask, int32_t n_seg, float grad_scale,
                                      float *loss_out, float *v_colors, void *stream);

/* K8b full backward (replaces rasterize_to_pixels_bwd; App. A.6).  Outputs must be
 * zero-initialised; v_colors may be NULL (skip), v_alphas may be NULL (= 0).  Any D gsplat accepts
 * (wide D runs in channel blocks of 256 that accumulate into the same geometry gradients).
 * Out: v_means2d[N,2] v_conics[N,3] v_opacities[N] v_colors[N,D]                              */
int gags_blend_bwd_full(const float *geom, const float *colors, int32_t D,
                        const float *background, int32_t width, int32_t height,
                        const int32_t *offsets, const int32_t *flatten_ids,
                        const float *render_alphas, const int32_t *last_ids,
                        const float *v_render, const float *v_alphas, float *v_means2d,
                        float *v_conics, float *v_opacities, float *v_colors, void *stream);

/* ---------------------------------------------------------------------------------------------
 * §8f-1  fused distillation loss: loss = mean(|render - target| * mask) and its gradient in one
 * pass (replaces l1_loss(feature_map*mask, gt*mask), /root/reference/utils/loss_utils.py:20-21 and
 * train.py:162-163).  render/target/v_render are channel-last [HW, D]; mask[HW] or NULL.
 * loss_out[1] must be zero-initialised; v_render = sign(r - t) * mask^2... see DESIGN.md.       */
int gags_l1_loss_fused(const float *render, const float *target, const float *mask, int64_t HW,
                       int32_t D, float grad_scale, float *loss_out, float *v_render,
                       void *stream);

/* The same loss with the target in the reference's compact per-view form — seg[HW] int32 segment
 * ids (< 0 = no target) and emb[n_seg, D] per-segment embeddings, i.e. the inputs that
 * read_sam_clip_feature (/root/reference/scene/dataset_readers.py:54-121, train.py:162) gathers into
 * a dense map every iteration: target[p, :] = emb[seg[p], :].  No dense target is materialised.  */
int gags_l1_loss_segmap(const float *render, const int32_t *seg, const float *emb,
                        const float *mask, int64_t HW, int32_t D, int32_t n_seg, float grad_scale,
                        float *loss_out, float *v_render, void *stream);

/* The same loss against the reference's FULL target (read_sam_clip_feature, default mode,
 * /root/reference/scene/dataset_readers.py:54-121 called at /root/reference/train.py:162-166): three
 * SAM levels of segment ids seg3[3][HW] (-1 = none), one table emb[n_seg, D] and the per-pixel level
 * weights scale_map3[3][HW] (the scale decoder's output at image size):
 *   target[p,:] = sum_l scale_map3[l,p] * emb[seg3[l,p],:],  valid(p) = all three ids in [0, n_seg).
 * v_scale_map (optional, [3][HW], zeroed by the caller, needs D % 128 == 0) receives d loss /
 * d scale_map3 — the path through which train.py trains the scale decoder (:149 -> :162).        */
int gags_l1_loss_sam(const float *render, const int32_t *seg3, const float *emb,
                     const float *scale_map3, int64_t HW, int32_t D, int32_t n_seg,
                     float grad_scale, float *loss_out, float *v_render, float *v_scale_map,
                     void *stream);
/* ... and fused into the cached feature backward (cf. gags_blend_bwd_features_cached_l1).        */
int gags_blend_bwd_features_cached_sam(int32_t D, int32_t width, int32_t height,
                                       const int32_t *offsets, const void *wcache,
                                       const int32_t *wmeta, const int32_t *wlist, int32_t *wcount,
                                       const float *render, const int32_t *seg3, const float *emb,
                                       const float *scale_map3, int32_t n_seg, float grad_scale,
                                       float *loss_out, float *v_scale_map, float *v_colors,
                                       void *stream);

/* Per-pixel losses over the channel dimension of channel-last rows a, b [HW, D] (D % 4 == 0):
 *   mode 0 = l1_loss_map (/root/reference/utils/loss_utils.py:23-24): out[p] = mean_c |a - b|;
 *   mode 1 = the per-pixel cosine similarity of cos_loss (:29-30, eps 1e-8); stats[HW][2] = |a|, |b|.
 * gags_pixel_loss_bwd gives v_a = g[p] * d out[p] / d a (b is the target).                         */
int gags_pixel_loss_fwd(int32_t mode, const float *a, const float *b, int64_t HW, int32_t D,
                        float *out, float *stats, void *stream);
int gags_pixel_loss_bwd(int32_t mode, const float *a, const float *b, const float *g,
                        const float *out, const float *stats, int64_t HW, int32_t D, float *v_a,
                        void *stream);

/* v[0..numel) *= *scale_dev (a DEVICE scalar), a no-op pass when the scalar is exactly 1: chains the
 * fused loss's stored gradient with autograd's incoming grad_output without a host sync and, in the
 * usual loss.backward() case, without touching the 2 GB buffer.  numel % 4 == 0.                */
int gags_scale_inplace(float *v, const float *scale_dev, int64_t numel, void *stream);

/* §8f-2  fused Adam on the per-Gaussian feature table (replaces torch.optim.Adam(lr, eps=1e-15)
 * on _semantic_feature, /root/reference/scene/gaussian_model.py:199,208; train.py:222-223).
 * Updates param/m/v in place; if zero_grad != 0 the gradient is zeroed in the same pass.       */
int gags_adam_step(float *param, float *grad, float *exp_avg, float *exp_avg_sq, int64_t numel,
                   double lr, double beta1, double beta2, double eps, int32_t step,
                   int32_t zero_grad, void *stream);

/* Row-sparse form for a [rows, D] table (D % 4 == 0) whose gradient is mostly zero rows — one view
 * touches a small part of the Gaussians, yet torch.optim.Adam (the reference's optimiser) reads and
 * the training loop re-zeroes the whole [N, D] gradient every step ("dense
 * even though only visible rows have grad").  row_flags[r] == 0 PROMISES gradient row r is all zero:
 * the row takes the g = 0 update (m, v decay, p moves by its momentum — bit-identical to the dense
 * pass) without its gradient being read.  Flagged rows are applied AND zeroed, and the flags are
 * cleared, so afterwards the gradient buffer is all zero again with no separate fill.            */
int gags_adam_step_rows(float *param, float *grad, float *exp_avg, float *exp_avg_sq,
                        uint8_t *row_flags, int64_t rows, int32_t D, double lr, double beta1,
                        double beta2, double eps, int32_t step, void *stream);
/* Lazily evaluated form of the same update.  With g = 0 a row's step depends only on the row's own
 * (p, m, v) and on two scalars of the step, so a row no view touches for k steps takes those k steps
 * later, in registers, in one visit: the same fp32 operations in the same order (bit-identical to
 * the dense pass), for 1/k of the memory traffic.  last_step[r] = the optimiser step row r is current
 * to; step_consts[2 s], [2 s + 1] = the pair gags_adam_step_consts gives for step s (one entry per
 * step taken, written by the caller).  Every row selected by row_flags (NULL = all rows: a flush)
 * takes the zero-gradient steps last_step[r]+1 .. t_to, then — if t_apply = t_to + 1 (0 = none) —
 * step t_apply with its gradient row, which is re-zeroed; last_step[r] is advanced.              */
int gags_adam_step_consts(double lr, double beta1, double beta2, int32_t step, float *out2_host);
int gags_adam_lazy_rows(float *param, float *grad, float *exp_avg, float *exp_avg_sq,
                        uint8_t *row_flags, int32_t *last_step, const float *step_consts,
                        int64_t rows, int32_t D, int32_t t_to, int32_t t_apply, double beta1,
                        double beta2, double eps, int32_t clear_flags, void *stream);
/* Sets row_flags[g] = 1 for every Gaussian of every batch gags_blend_fwd_cached kept for this view
 * (the rows the cached feature backward reduces into; never clears a flag).                      */
int gags_blend_cache_mark_rows(int32_t width, int32_t height, const int32_t *offsets,
                               const int32_t *wmeta, const int32_t *wlist, const int32_t *wcount,
                               uint8_t *row_flags, void *stream);

/* Multi-GPU form of gags_adam_step: gradient all-reduce + Adam + parameter all-gather in ONE kernel
 * over NVLink peer memory (view-parallel training).  grad_ptrs[q] / param_ptrs[q]
 * (host arrays of `world` device addresses, q = rank) are every rank's full [numel] gradient and
 * parameter buffer, mapped into this process (CUDA IPC / symmetric memory).  This rank owns
 * elements [start, start + count) (multiples of 4): it sums that range of all gradients in rank
 * order, updates its slice of the moments (exp_avg_shard / exp_avg_sq_shard hold `count` floats)
 * and writes the new parameters into every rank's buffer.  The caller provides an inter-rank
 * barrier before (all gradients final) and after (all replicas written, gradients consumed).   */
int gags_adam_step_peer(int32_t world, int32_t rank, const uint64_t *grad_ptrs,
                        const uint64_t *param_ptrs, float *exp_avg_shard, float *exp_avg_sq_shard,
                        int64_t start, int64_t count, double lr, double beta1, double beta2,
                        double eps, int32_t step, void *stream);

/* gags_adam_step_peer through the NVSwitch: mc_grad / mc_param are the MULTICAST addresses of the
 * symmetric gradient / parameter buffers (NVLS).  The gradient sum is formed inside the switch
 * (multimem.ld_reduce) and one multimem.st updates every replica, so each rank moves 1/world of
 * the table per direction.  param_local = this rank's own (unicast) parameter buffer.  Same
 * barriers around the call as gags_adam_step_peer.                                              */
int gags_adam_step_multicast(const float *mc_grad, float *mc_param, const float *param_local,
                             float *exp_avg_shard, float *exp_avg_sq_shard, int64_t start,
                             int64_t count, double lr, double beta1, double beta2, double eps,
                             int32_t step, void *stream);

/* Row-sparse gradient all-reduce for view-parallel training: only the rows some
 * rank's backward flagged (gags_blend_cache_mark_rows) are summed over the ranks, in place, into
 * every replica's gradient buffer; each rank keeps the full optimiser state and follows with its own
 * gags_adam_step_rows driven by `union_flags`.  grad_ptrs[q] / flag_ptrs[q] (host arrays of `world`
 * device addresses): every rank's [rows, D] gradient buffer and its uint8 row flags (4 * ceil(rows/4)
 * bytes, zero padded), mapped into this process; mc_grad / mc_flags: their NVLS multicast addresses
 * (multimem.ld_reduce / multimem.st), or both NULL for plain peer loads / stores.  union_flags
 * (local, same size as the flags) receives the OR over ranks.  Barriers before and after the call
 * are the caller's, as for gags_adam_step_peer.                                                   */
int gags_grad_allreduce_rows(int32_t world, int32_t rank, const uint64_t *grad_ptrs,
                             const uint64_t *flag_ptrs, float *mc_grad, const uint8_t *mc_flags,
                             uint8_t *union_flags, int64_t rows, int32_t D, void *stream);

/* Tuning hook: CTAs per SM of the two exchange kernels' grids (0 = built-in: 4 unicast, 2
 * multicast).  Process-wide; used by tools/peer_rate.py.                                         */
int gags_set_peer_grid(int32_t ctas_per_sm);
/* Tuning hook: switch reductions in flight per thread of the multicast exchange (2, 4 or 8).      */
int gags_set_peer_unroll(int32_t unroll);

/* Tuning hook: wide forward blend variant.  3 (default) = alpha evaluation (lane = Gaussian) and
 * transmittance chain (lane = pixel) in two warp groups; 2 = one thread per pixel does both (the
 * round-1 kernel); add 10 (12 / 13) for per-lane streaming stores instead of the TMA tensor-store
 * epilogue.  Bit-identical outputs; process-wide; used by the parity tests and bench.             */
int gags_set_fwd_variant(int32_t variant);

/* Tuning hook: 1 (default) = gags_blend_fwd_from_cache as ONE persistent CTA per SM (double-buffered
 * TMEM accumulators: the epilogue of a half tile overlaps the MMAs of the next), 0 = one CTA per half
 * tile.  Bit-identical outputs; process-wide.                                                     */
int gags_set_blend_pass(int32_t persistent);

/* Tuning hook: 1 (default) = gags_blend_bwd_features_cached_l1 without a mask stages the exact sign
 * operand (one bf16 part, scale applied to the accumulator); 0 = the hi / lo split of scale * sign.  */
int gags_set_bwd_sign_operand(int32_t on);

/* Zero-fill with a small grid (a quarter of the thread slots), meant to run on a second stream
 * beside latency-bound kernels; ptr 16-B aligned, bytes % 16 == 0.                               */
int gags_zero_fill(void *ptr, int64_t bytes, void *stream);

/* cudaMemsetAsync(ptr, 0, bytes) on `stream` (driver memset; keeps the 2 GB gradient-buffer fill off
 * the SMs that the neighbouring stream's kernels need). */
int gags_memset_zero(void *ptr, size_t bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GAGS_B200_H */
