/* bds.h - C ABI of libbds_b200.so: the B200-native (sm_100a) render + bilateral hot path.
 *
 * Nothing like this exists in the reference (it is pure Python over pip packages); each entry
 * point names the reference interface whose arithmetic it replaces (paths relative to
 * /root/reference/project).  INTEGRATION.md shows the ctypes / torch binding a maintainer adds.
 *
 * Conventions (all entry points):
 *   - return 0 on success, <0 on error; bds_last_error() gives the thread-local message;
 *   - every pointer is a DEVICE pointer owned by the caller unless marked host_; the library
 *     never allocates or frees device memory - scratch comes from the caller, sized by the
 *     matching *_workspace_bytes query;
 *   - every call is asynchronous on the given cudaStream_t (passed as void*), re-entrant across
 *     streams and devices, and keeps no mutable global state beyond lazily set function attributes;
 *   - all tensors are dense, row-major, fp32 unless noted; ids are int32.
 */
#ifndef BDS_H_
#define BDS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BDS_ABI_VERSION 5
#define BDS_MAX_LEVELS 4
#define BDS_COUNTERS_LEN 4096  /* int32 entries of the projection counters buffer */
#define BDS_TILE 16
#define BDS_SPLAT_FLOATS 12 /* one packed splat record = 48 bytes */

typedef void* bds_stream_t; /* cudaStream_t */

const char* bds_last_error(void);
int bds_abi_version(void);
/* returns the compute capability major*10+minor of the current device, or <0 */
int bds_device_arch(void);
/* number of kernels this library has launched so far in this process (statistics for bench.py) */
unsigned long long bds_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Bilateral grids.  Replaces bilateral/lib_bilagrid.py:171-230 (slice), :317-368
 * (BilateralGrid.forward -> F.grid_sample 5-D), models/modules.py:494-504 (get_sample_grid),
 * :409-420 (fill_matrix_res), :505-584 (MultiScaleBilateralAffineTransform.forward) and the
 * sequential 3x4 apply of models/trainers/scene_graph.py:112-117.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int n_levels;
  int L[BDS_MAX_LEVELS];      /* guidance (luma) resolution   (grid_W in the reference ctor) */
  int GY[BDS_MAX_LEVELS];     /* grid height                  (grid_Y) */
  int GX[BDS_MAX_LEVELS];     /* grid width                   (grid_X) */
  int factor[BDS_MAX_LEVELS]; /* guidance down-sampling factor of modules.py:505; 0 or 1 = the
                                 guidance_factor=None branch (full-resolution guidance) */
} bds_bilateral_desc;

size_t bds_bilateral_workspace_bytes(const bds_bilateral_desc* d, int H, int W);

/* rgb_in [H,W,3]; grids[l] -> the image's grid slot, reference layout [12,L,GY,GX];
 * rgb_out [H,W,3]; affine_out[l] (optional, host array may be NULL or hold NULLs) [H,W,12]
 * = the "rgb_affine_mats" the reference module returns. */
int bds_bilateral_fwd(const bds_bilateral_desc* d, int H, int W, const float* rgb_in,
                      const float* const* host_grids, float* rgb_out, float* const* host_affine_out,
                      void* workspace, bds_stream_t stream);

/* v_rgb_out [H,W,3]; v_affine[l] optional extra cotangent on affine_out[l];
 * v_rgb_in [H,W,3] (written); v_grids[l] [12,L,GY,GX] (ACCUMULATED into: caller zero-fills). */
int bds_bilateral_bwd(const bds_bilateral_desc* d, int H, int W, const float* rgb_in,
                      const float* const* host_grids, const float* v_rgb_out,
                      const float* const* host_v_affine, float* v_rgb_in, float* const* host_v_grids,
                      void* workspace, bds_stream_t stream);

/* Generic per-sample slice (BilateralGrid.forward with arbitrary xy, lib_bilagrid.py:317-368):
 * xy [n,2] in [0,1], rgb [n,3] -> affine [n,12]. */
int bds_bilagrid_slice_fwd(const float* grid, int L, int GY, int GX, int n, const float* xy,
                           const float* rgb, float* affine, bds_stream_t stream);
int bds_bilagrid_slice_bwd(const float* grid, int L, int GY, int GX, int n, const float* xy,
                           const float* rgb, const float* v_affine, float* v_grid /*accumulated*/,
                           float* v_rgb /*written*/, bds_stream_t stream);

/* Total-variation loss over all image slots (lib_bilagrid.py:152-168, modules.py:466-472).
 * grids [N,12,L,GY,GX]; loss (1 float, ACCUMULATED: weight * tv); v_grids ACCUMULATED with
 * v_loss * weight * dtv/dgrid when v_grids != NULL. */
int bds_tv_fwd_bwd(const float* grids, int N, int L, int GY, int GX, float weight, float v_loss,
                   float* loss, float* v_grids, bds_stream_t stream);

/* The same for ALL levels of a multi-scale module in ONE launch (modules.py:466-472 loops over the levels):
 * loss += sum_l weights[l] * tv(grids[l]); host_grids / host_v_grids: host arrays of n_levels device pointers
 * (v_grids may be NULL = loss only, or hold NULLs); N, L, GY, GX, weights: host arrays of n_levels entries. */
int bds_tv_levels_fwd_bwd(int n_levels, const float* const* host_grids, const int* N, const int* L, const int* GY,
                          const int* GX, const float* weights, float v_loss, float* loss, float* const* host_v_grids,
                          bds_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Spherical harmonics.  Replaces gsplat.cuda._wrapper.spherical_harmonics as called at
 * models/gaussians/vanilla.py:383-389 (import seam models/gaussians/basics.py:15).
 * dirs [n,3] (normalised inside), coeffs [n,K,3], out [n,3]; bands above `degree` ignored.
 * ------------------------------------------------------------------------------------------ */
int bds_sh_fwd(int n, int degree, int K, const float* dirs, const float* coeffs, float* out,
               bds_stream_t stream);
/* v_coeffs [n,K,3] fully written (zeros for inactive bands); v_dirs [n,3] optional (NULL = skip) */
int bds_sh_bwd(int n, int degree, int K, const float* dirs, const float* coeffs, const float* v_out,
               float* v_coeffs, float* v_dirs, bds_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Render.  Replaces gsplat.rendering.rasterization as called at models/trainers/base.py:393-408
 * (fully_fused_projection -> isect_tiles -> radix sort -> isect_offset_encode ->
 * rasterize_to_pixels), the glue at base.py:414-417 / scene_graph.py:287-294, and
 * (fused epilogue) the bilateral chain above.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int n_gauss;       /* N */
  int n_cams;        /* C */
  int width, height; /* pixels */
  float near_plane, far_plane, radius_clip, eps2d;
  int antialiased;   /* rasterize_mode == "antialiased" */
  /* vertical band of tile rows rendered by this call (multi-GPU tile sharding):
     cameras [cam_begin, cam_end) x tile rows; band = global tile-row range over the
     concatenation of all cameras' tile rows: [row_begin, row_end) in [0, C*tile_rows]. */
  int row_begin, row_end;
  /* optional fused activation + SH fast path (vanilla.py:122-146, :383-389) */
  int raw_params;    /* 1: scales are log-scales, quats un-normalised (always normalised anyway),
                           opacities are logits */
  int sh_degree;     /* >=0: colours come from SH coefficients (features_dc/rest), evaluated for
                        visible Gaussians only, +0.5 and clamp[0,1] applied; <0: colours given */
  int sh_K;          /* number of SH bases held per Gaussian (1 + rest) */
} bds_render_desc;

/* Projection (+ optional activations and SH).  Writes gsplat-shaped per-(cam,gauss) outputs
 * radii [C,N] int32, means2d [C,N,2], depths [C,N], conics [C,N,3], (compensations [C,N] or NULL),
 * tiles_touched [C,N] int32 (exact number of (tile) records the Gaussian will emit inside the
 * band), tile_counts [band tiles + 1] int32 (records per band tile, last entry 0; zero-filled by the
 * call; band tile t = (global_row - row_begin)*tile_w + tx), and compacts the visible splats into
 * packed 48-byte records:
 *   splats [cap,12] = {x, y, a', b', c', opacity, r, g, b, depth, bits(flat id c*N+n), log2(opacity)}
 * with (a',b',c') = log2(e) * (a/2, b, c/2) of the conic (alpha = exp2(log2(opacity) - sigma'));
 * slot_of [C,N] int32 = record index or -1 (optional, may be NULL); counters = BDS_COUNTERS_LEN device int32,
 * zero-filled by the caller: [0] = number of records, [1] is set to 1 on capacity overflow, [2] = length of the
 * queue of very large splats (slots in [4..]) whose tiles are counted by a follow-up launch of the same call.
 * camera position for the SH view direction is taken from viewmats. */
int bds_project_fwd(const bds_render_desc* d, const float* means, const float* quats,
                    const float* scales, const float* opacities, const float* colors /*[N,3] or [C,N,3] or NULL*/,
                    int colors_per_cam, const float* features_dc /*[N,3] or NULL*/,
                    const float* features_rest /*[N,K-1,3] or NULL*/, const float* viewmats,
                    const float* Ks, int32_t* radii, float* means2d, float* depths, float* conics,
                    float* compensations, int32_t* tiles_touched, int32_t* tile_counts, float* splats,
                    int32_t splat_cap, int32_t* slot_of, int32_t* counters, bds_stream_t stream);

/* Binning: per-tile start offsets from the per-tile counts, emission of every (tile, splat) pair into its
 * tile's segment (atomic cursor per tile), ONE CTA per tile sorts its segment by (depth bits, Gaussian
 * id) - shared memory up to 4096 records, in place in global memory beyond - and gathers the packed
 * records in that order.  Two steps because the record count is data dependent:
 *   bds_bin_count   -> tile_offsets [band tiles + 1] int32 = exclusive scan of tile_counts;
 *                      n_isect_dev (device int64) = total number of records
 *   (caller reads n_isect, allocates)   bds_bin_sort -> sorted records. */
size_t bds_bin_count_workspace_bytes(const bds_render_desc* d);
int bds_bin_count(const bds_render_desc* d, const int32_t* tile_counts /*[band tiles + 1]*/,
                  int32_t* tile_offsets /*[band tiles + 1]*/, int64_t* n_isect_dev, void* workspace,
                  bds_stream_t stream);
size_t bds_bin_sort_workspace_bytes(const bds_render_desc* d, int64_t n_isect);
/* sorted_splats [n_isect,12] (field 10 = splat slot); sorted_slots [n_isect] int32 or NULL */
int bds_bin_sort(const bds_render_desc* d, int64_t n_isect, int32_t n_slots /* = counters[0] */,
                 const int32_t* radii, const float* splats, const int32_t* tile_offsets,
                 float* sorted_splats, int32_t* sorted_slots, void* workspace, bds_stream_t stream);

/* Epilogue description for the fused composite kernel. */
typedef struct {
  int mode;          /* 0: plain gsplat outputs (render_colors [.,.,D], alphas); 1: reference glue
                        fused: clamp(rgb,max=1), expected depth, sky composite; 2: mode 1 + the
                        multi-scale bilateral chain with full-resolution guidance */
  int channels;      /* mode 0: D in {3,4}; channel 3 = depth */
  int expected_depth;/* mode 0: divide depth channel by max(alpha,1e-10) (ED / RGB+ED) */
  bds_bilateral_desc bil; /* mode 2 */
} bds_epilogue_desc;

/* Fused front-to-back composite (+ epilogue).  One CTA per band tile.
 * mode 0 outputs: render [P,D], alpha [P];                        (P = band pixels, row-major
 * mode 1/2 outputs: rgb [P,3], rgb_gauss [P,3], depth [P], alpha [P]   per camera then y then x)
 * always: last_ids [P] int32 (index into the sorted records of the last contributing one, -1 none)
 * backgrounds: mode 0 optional [C,D]; sky: modes 1/2 [P,3] (may be NULL = black).
 * grids (mode 2): host array of C*n_levels device pointers, grids[c*n_levels+l] -> [12,L,GY,GX]. */
int bds_composite_fwd(const bds_render_desc* d, const bds_epilogue_desc* e, const float* sorted_splats,
                      const int32_t* tile_offsets, const float* backgrounds, const float* sky,
                      const float* const* host_grids, float* out_rgb, float* out_rgb_gauss,
                      float* out_depth, float* out_alpha, int32_t* last_ids, void* workspace,
                      bds_stream_t stream);
size_t bds_composite_workspace_bytes(const bds_render_desc* d, const bds_epilogue_desc* e);

/* Masked re-render over the SAME sorted lists (no projection / binning / sort): replaces the reference's
 * render_fn(opacity_mask) = a second full gsplat rasterization with opacities * mask
 * (models/trainers/base.py:392-419, called per class at models/trainers/scene_graph.py:296-313).
 * A splat whose keep flag is 0 behaves exactly like a splat of opacity 0 (it fails alpha >= 1/255 on
 * every pixel), so the images equal the reference's re-rasterization bit for bit.
 * bds_slot_keep: slot_keep[slot] = gaussian_keep[Gaussian of that splat]  (gaussian_keep [N] bytes,
 * slot_keep [n_slots] bytes; splats / counters as written by bds_project_fwd). */
int bds_slot_keep(const bds_render_desc* d, const float* splats, const int32_t* counters, int32_t n_slots,
                  const uint8_t* gaussian_keep, uint8_t* slot_keep, bds_stream_t stream);
int bds_composite_fwd_masked(const bds_render_desc* d, const bds_epilogue_desc* e, const float* sorted_splats,
                             const int32_t* tile_offsets, const uint8_t* slot_keep, const float* backgrounds,
                             const float* sky, const float* const* host_grids, float* out_rgb,
                             float* out_rgb_gauss, float* out_depth, float* out_alpha, int32_t* last_ids,
                             void* workspace, bds_stream_t stream);

/* Backward of the above.  Cotangents: v_rgb [P,D or 3], v_rgb_gauss [P,3] (optional), v_depth [P]
 * (optional), v_alpha [P] (optional).  Outputs: v_splats [n_slots,12] ACCUMULATED (caller zeroes; 16-byte aligned):
 * {m_x, m_y, m_xx, m_xy, m_yy, m_0, v_r, v_g, v_b, v_depth, sum|w g_x|, sum|w g_y|} - the pixel moments of
 * w = alpha * v_alpha about the splat's mean (every 2-D gradient is linear in them; bds_project_bwd converts);
 * v_sky [P,3] optional (written); v_grids: host array like host_grids, ACCUMULATED;
 * v_backgrounds [C,D] optional ACCUMULATED. */
int bds_composite_bwd(const bds_render_desc* d, const bds_epilogue_desc* e, const float* sorted_splats,
                      const int32_t* sorted_slots, const int32_t* tile_offsets, const float* backgrounds,
                      const float* sky, const float* const* host_grids, const float* out_rgb_gauss,
                      const float* out_depth, const float* out_alpha, const int32_t* last_ids,
                      const float* v_rgb, const float* v_rgb_gauss, const float* v_depth,
                      const float* v_alpha, float* v_splats, float* v_sky, float* const* host_v_grids,
                      float* v_backgrounds, void* workspace, bds_stream_t stream);

/* Projection backward: consumes the per-splat cotangents (plus optional dense extras
 * v_means2d_extra [C,N,2], v_depths_extra [C,N], v_conics_extra [C,N,3] from users of the gsplat
 * info tensors) and ACCUMULATES into v_means [N,3], v_quats [N,4], v_scales [N,3],
 * v_opacities [N], and either v_colors ([N,3] / [C,N,3]) or v_features_dc/v_features_rest;
 * v_viewmats [C,4,4] optional (ACCUMULATED).  Also writes the dense densification taps
 * v_means2d [C,N,2] and absgrad [C,N,2] when non-NULL (caller zero-fills). */
int bds_project_bwd(const bds_render_desc* d, const float* means, const float* quats,
                    const float* scales, const float* opacities, const float* colors, int colors_per_cam,
                    const float* features_dc, const float* features_rest, const float* viewmats,
                    const float* Ks, const float* splats, const int32_t* counters,
                    const float* v_splats, const float* v_means2d_extra, const float* v_depths_extra,
                    const float* v_conics_extra, float* v_means, float* v_quats, float* v_scales,
                    float* v_opacities, float* v_colors, float* v_features_dc, float* v_features_rest,
                    float* v_viewmats, float* v_means2d, float* absgrad, bds_stream_t stream);

/* Multi-GPU gradient exchange with the SH gradient in COMPACT form.  The gradient of the 3 K SH coefficients of a
 * Gaussian is, per camera, the outer product of the K basis values of its view direction and ONE 3-vector: the colour
 * cotangent.  bds_project_bwd_compact_sh is bds_project_bwd for the SH fast path (sh_degree >= 0) writing that 3-vector
 * per (camera, Gaussian) into v_sh_color [C,N,3] (caller zero-fills) instead of accumulating v_features_dc / _rest; the
 * ranks then all-reduce 11 + 3 C floats per Gaussian instead of 59, and bds_sh_expand_bwd rebuilds v_features_dc [N,3]
 * and v_features_rest [N,K-1,3] (fully WRITTEN) from the reduced v_sh_color. */
int bds_project_bwd_compact_sh(const bds_render_desc* d, const float* means, const float* quats, const float* scales,
                               const float* opacities, const float* viewmats, const float* Ks, const float* splats,
                               const int32_t* counters, const float* v_splats, float* v_means, float* v_quats,
                               float* v_scales, float* v_opacities, float* v_sh_color, float* v_viewmats,
                               bds_stream_t stream);
int bds_sh_expand_bwd(const bds_render_desc* d, const float* means, const float* viewmats, const float* v_sh_color,
                      float* v_features_dc, float* v_features_rest, bds_stream_t stream);

/* Companion of bds_project_bwd for the optional dense extras only: a Gaussian that gsplat calls visible
 * (radii > 0) but whose alpha >= 1/255 footprint misses every tile of the band owns no splat record here
 * (slot_of < 0), so the record walk of bds_project_bwd never sees a cotangent a user of the gsplat info tensors
 * put on its means2d / depths / conics.  This entry adds exactly those contributions (ACCUMULATED into the same
 * outputs).  radii, slot_of [C,N] int32 as written by bds_project_fwd.  No-op when all three extras are NULL. */
int bds_project_bwd_extras(const bds_render_desc* d, const float* means, const float* quats, const float* scales,
                           const float* viewmats, const float* Ks, const int32_t* radii, const int32_t* slot_of,
                           const float* v_means2d_extra, const float* v_depths_extra, const float* v_conics_extra,
                           float* v_means, float* v_quats, float* v_scales, float* v_viewmats, bds_stream_t stream);

/* Densification statistics of one step in one launch.  Replaces VanillaGaussians.after_train
 * (models/gaussians/vanilla.py:163-191) as called by BasicTrainer.postprocess_per_train_step (base.py:279-297):
 * radii [n] int32 (info["radii"] of the step's camera), xys_grad [n,2] (info["means2d"].absgrad or .grad; scaled here
 * by scale_x / scale_y - pass 1 when the caller already applied base.py:285-286), last_size = max(W, H).
 * first != 0 (the statistics were reset to None, vanilla.py:172-175): xys_grad_norm = |xys_grad| for EVERY Gaussian,
 * vis_counts = 1; else for visible Gaussians (radii > 0): xys_grad_norm += |xys_grad|, vis_counts += 1.
 * Always for visible Gaussians: max_2dsize = max(max_2dsize, radii / last_size).  All three [n] fp32, updated in place. */
int bds_densify_stats(int64_t n, const int32_t* radii, const float* xys_grad, float scale_x, float scale_y,
                      float last_size, int first, float* xys_grad_norm, float* vis_counts, float* max_2dsize,
                      bds_stream_t stream);

/* Fused photometric loss used by the benchmark step: mean((rgb-gt)^2) +
 * lambda_d*mean(depth) + lambda_a*mean(alpha); writes the cotangents and ACCUMULATES the loss. */
int bds_loss_fwd_bwd(int64_t n_pix, const float* rgb, const float* gt, const float* depth,
                     const float* alpha, float lambda_d, float lambda_a, float inv_count,
                     float* loss, float* v_rgb, float* v_depth, float* v_alpha, bds_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BDS_H_ */
