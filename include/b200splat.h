/*
 * b200splat.h -- C ABI of libb200splat.so: the B200 (sm_100a) rasterizer hot path.
 *
 * This is the drop-in boundary below the gsplat Python operators.  Each entry point
 * replaces one function of the reference's pybind/libtorch module
 * (/root/reference/gsplat/gsplat/cuda/csrc/ext.cpp:4-18, declared in bindings.h:19-225);
 * the reference interface it replaces is cited per function.
 *
 * Conventions (all functions):
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless
 *     marked "host"; outputs are caller-allocated and need NOT be zero-initialised
 *     (the kernels write every element, including the zeros the reference obtains
 *     from its torch::zeros allocations);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy
 *     default stream) of the CURRENT device and the call returns without synchronising;
 *   - return value 0 = success, negative = error (B200_ERR_*); b200_last_error()
 *     gives a thread-local message.  No exceptions cross the boundary; reentrant.
 *   - float = IEEE fp32, ids/offsets int32, sort keys int64 -- as in the reference.
 */
#ifndef B200SPLAT_H
#define B200SPLAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200_API __attribute__((visibility("default")))
#else
#define B200_API
#endif

#define B200_ABI_VERSION 1
#define B200_MAX_BLUR_SAMPLES 10 /* helpers.cuh:222 */
#define B200_OK 0
#define B200_ERR_INVALID (-1) /* bad argument (the reference raises TORCH_CHECK / AT_ERROR) */
#define B200_ERR_CUDA (-2)    /* a CUDA runtime call or launch failed */

#define B200_SH_POLY 0 /* sh.cuh:268-340 */
#define B200_SH_FAST 1 /* sh.cuh:54-156  */

/* project_gaussians_backward flags */
#define B200_PROJ_EXACT 1u /* clamp-aware J and full dL/dviewmat: the gradients of the reference's
                              torch path (project_gaussians.py:81-112); 0 = reference CUDA path
                              (backward.cu:454-532 ignores the fov clamp; project_gaussians.py:272-307
                              approximates dL/dR) */

B200_API int b200_abi_version(void);
B200_API const char *b200_last_error(void);
/* Number of kernels this library has launched in this process (hand-written kernels 1 each, a cub
 * primitive counts its documented passes); bench.py reports the per-step delta as gpu_launches. */
B200_API long long b200_launch_count(void);

/* Size of one packed per-Gaussian blend record (bytes); scratch for the blend = N * this. */
B200_API size_t b200_packed_record_bytes(void);

/* ---- projection ------------------------------------------------------------------------
 * replaces project_gaussians_forward_tensor (bindings.h:43-72, bindings.cu:154-257; kernel
 * forward.cu:13-112).  lin_vel / ang_vel are DEVICE float[3] (NULL = zero) so the caller never
 * needs the .tolist() D2H sync of project_gaussians.py:178-179.  viewmat: >= 12 floats row-major.
 * Outputs: cov3d (N,6) xys (N,2) depths (N) pix_vels (N,2) radii (N) i32 conics (N,3)
 * compensation (N) num_tiles_hit (N) i32.  quat_norm_flag (optional DEVICE int32, caller-zeroed) is OR-ed with 1 if
 * any quaternion violates the reference's `quats.norm(dim=-1) - 1 < 1e-6` assert (project_gaussians.py:69), so the
 * caller can raise it at its next host sync instead of paying a separate reduction + sync. */
B200_API int b200_project_gaussians_forward(int num_points, const float *means3d, const float *scales, float glob_scale,
                                   const float *quats, const float *lin_vel, const float *ang_vel,
                                   float rolling_shutter_time, float exposure_time, const float *viewmat, float fx,
                                   float fy, float cx, float cy, unsigned img_height, unsigned img_width,
                                   unsigned block_width, float clip_thresh, float *cov3d, float *xys, float *depths,
                                   float *pix_vels, int32_t *radii, float *conics, float *compensation,
                                   int32_t *num_tiles_hit, int32_t *quat_norm_flag, void *stream);

/* replaces project_gaussians_backward_tensor (bindings.h:74-108, bindings.cu:259-358; kernel
 * backward.cu:371-451) AND the Python-side v_viewmat block (project_gaussians.py:272-307) AND adds
 * the camera-velocity gradients the reference only has through its torch path.
 * v_cov2d (N,3) / v_cov3d (N,6) are optional scratch outputs (NULL = not materialised).
 * v_mean3d (N,3) v_scale (N,3) v_quat (N,4) are written for every Gaussian (zeros where radii<=0).
 * v_lin_vel[3], v_ang_vel[3], v_viewmat[12] are optional; when non-NULL they are OVERWRITTEN with
 * the sums over Gaussians (block-reduced in the kernel, one atomic per block and component). */
B200_API int b200_project_gaussians_backward(int num_points, const float *means3d, const float *scales, float glob_scale,
                                    const float *quats, const float *lin_vel, const float *ang_vel,
                                    float rolling_shutter_time, float exposure_time, const float *viewmat, float fx,
                                    float fy, float cx, float cy, unsigned img_height, unsigned img_width,
                                    const float *cov3d, const int32_t *radii, const float *conics,
                                    const float *compensation, const float *v_xy, const float *v_depth,
                                    const float *v_pix_vel, const float *v_conic, const float *v_compensation,
                                    unsigned flags, float *v_cov2d, float *v_cov3d, float *v_mean3d, float *v_scale,
                                    float *v_quat, float *v_lin_vel, float *v_ang_vel, float *v_viewmat,
                                    void *stream);

/* replaces compute_cov2d_bounds_tensor (bindings.h:19-22, bindings.cu:19-60): conics (N,3), radii (N) f32 */
B200_API int b200_compute_cov2d_bounds(int num_pts, const float *cov2d, float *conics, float *radii, void *stream);

/* ---- spherical harmonics ----------------------------------------------------------------
 * replaces compute_sh_forward_tensor / compute_sh_backward_tensor (bindings.h:24-41,
 * bindings.cu:62-151; kernels sh.cuh:434-498).  coeffs (N,K,3), K = (degree+1)^2, 3 channels;
 * viewdirs (N,3) un-normalised; v_coeffs rows >= (degrees_to_use+1)^2 are written as zeros. */
B200_API int b200_compute_sh_forward(int method, int num_points, int degree, int degrees_to_use, const float *viewdirs,
                            const float *coeffs, float *colors, void *stream);
B200_API int b200_compute_sh_backward(int method, int num_points, int degree, int degrees_to_use, const float *viewdirs,
                             const float *v_colors, float *v_coeffs, void *stream);

/* ---- tile binning -----------------------------------------------------------------------
 * b200_cumulative_intersects replaces torch.cumsum in compute_cumulative_intersects
 * (gsplat/utils.py:106-125): inclusive int32 scan.  `total_host_pinned` (HOST, optional) receives
 * cum[N-1] by an async copy on `stream` (the caller synchronises before reading it); when `flag_dev` is given its
 * int32 is copied to total_host_pinned[1] in the same way (deferred quat-norm check, see above). */
B200_API size_t b200_scan_temp_bytes(int num_points);
B200_API int b200_cumulative_intersects(int num_points, const int32_t *num_tiles_hit, int32_t *cum_tiles_hit, void *temp,
                                        size_t temp_bytes, int32_t *total_host_pinned, const int32_t *flag_dev,
                                        void *stream);

/* replaces map_gaussian_to_intersects_tensor (bindings.h:199-209, bindings.cu:360-402; kernel
 * forward.cu:116-153).  Slots reserved by cum_tiles_hit but not emitted (the reference's "phantom"
 * entries, which keep their zeros init) are written as key 0 / id 0 here too. */
B200_API int b200_map_gaussian_to_intersects(int num_points, int num_intersects, const float *xys, const float *depths,
                                    const int32_t *radii, const int32_t *cum_tiles_hit, unsigned tiles_x,
                                    unsigned tiles_y, unsigned block_width, int64_t *isect_ids,
                                    int32_t *gaussian_ids, void *stream);

/* replaces torch.sort + torch.gather in bin_and_sort_gaussians (gsplat/utils.py:179-180): stable LSD
 * radix sort of (key, id) pairs on the bits that can be set (32 depth bits + ceil(log2(num_tiles))). */
B200_API size_t b200_sort_temp_bytes(int num_intersects);
B200_API int b200_sort_intersects(int num_intersects, int num_tiles, const int64_t *isect_ids, const int32_t *gaussian_ids,
                         int64_t *isect_ids_sorted, int32_t *gaussian_ids_sorted, void *temp, size_t temp_bytes,
                         void *stream);

/* replaces get_tile_bin_edges_tensor (bindings.h:211-215, bindings.cu:404-422; kernel forward.cu:158-180):
 * tile_bins (num_tiles,2) i32, (0,0) for empty tiles. */
B200_API int b200_get_tile_bin_edges(int num_intersects, int num_tiles, const int64_t *isect_ids_sorted, int32_t *tile_bins,
                            void *stream);

/* Fused fast path of bin_and_sort_gaussians for callers that only need what the blend consumes
 * (rasterize.py:146-163 discards the key arrays): gaussian_ids_sorted (I) and tile_bins (num_tiles,2), produced by a
 * two-level sort (Gaussians by depth, then pairs by tile id) that is order-identical to the stable sort of the
 * reference's 64-bit keys, phantom zero-key slots included.  `ws`: 256-byte aligned scratch of
 * b200_bin_tiles_ws_bytes(N, I) bytes.  num_intersects must equal sum(num_tiles_hit). */
B200_API size_t b200_bin_tiles_ws_bytes(int num_points, int num_intersects);
B200_API int b200_bin_tiles(int num_points, int num_intersects, const float *xys, const float *depths,
                            const int32_t *radii, const int32_t *num_tiles_hit, unsigned tiles_x, unsigned tiles_y,
                            unsigned block_width, void *ws, size_t ws_bytes, int32_t *gaussian_ids_sorted,
                            int32_t *tile_bins, void *stream);

/* Culled binning -- the path gsplat.rasterize_gaussians takes.  Same two-level sort, but a (tile, Gaussian) pair of
 * the reference's bbox (forward.cu:94-102: the square around a 3-sigma circle, inflated by the blur length) is kept only
 * if the Gaussian can reach alpha >= 1/255 inside that tile for some blur sample; dropped pairs cannot change any pixel,
 * survivors keep the reference's order, and the reference's phantom copies of Gaussian 0 survive iff it can touch
 * tile 0.  Two calls around ONE host sync:
 *   b200_bin_cull_count: per-Gaussian survivor counts (the 32-bit survival mask of every 32-tile chunk is kept for the
 *       emit pass), depth sort on a helper stream, offsets; totals_host_pinned[5] (HOST, pinned, async) =
 *       {sum(num_tiles_hit) i.e. the reference's num_intersects, phantom slots, Gaussian-0-touches-tile-0, culled
 *       entry count M, *flag_dev (0 if flag_dev is null: the deferred input-check word of
 *       b200_project_gaussians_forward rides along with the totals)};
 *   b200_bin_cull_emit: after the caller synchronised and allocated M ids -> gaussian_ids_sorted (M), tile_bins.
 * `packed` = records from b200_pack_records; ws_g (b200_bin_cull_ws_bytes(N) bytes) must stay alive and untouched
 * between the two calls; ws_e has b200_bin_cull_emit_ws_bytes(M) bytes; both 256-byte aligned. */
B200_API size_t b200_bin_cull_ws_bytes(int num_points);
B200_API size_t b200_bin_cull_emit_ws_bytes(int num_entries);
B200_API int b200_bin_cull_count(int num_points, const void *packed, const float *depths, const int32_t *radii,
                                 const int32_t *num_tiles_hit, unsigned img_height, unsigned img_width,
                                 unsigned block_width, unsigned n_blur_samples, float rolling_shutter_time,
                                 float exposure_time, void *ws_g, size_t ws_g_bytes, int32_t *totals_host_pinned,
                                 const int32_t *flag_dev, void *stream);
B200_API int b200_bin_cull_emit(int num_points, int num_entries, const void *packed, const int32_t *radii,
                                const int32_t *num_tiles_hit, unsigned img_height, unsigned img_width,
                                unsigned block_width, unsigned n_blur_samples, float rolling_shutter_time,
                                float exposure_time, const void *ws_g, void *ws_e, size_t ws_e_bytes,
                                int32_t *gaussian_ids_sorted, int32_t *tile_bins, void *stream);

/* Packed-record variants of the blend: b200_rasterize_forward/backward = b200_pack_records + these.  Lets a caller pack
 * once per render and share the records between culled binning, forward and backward.
 * out_alpha (H,W) may be null; when given it receives 1 - mean_s final_Ts, the alpha channel gsplat/rasterize.py:161-163
 * builds from final_Ts with two more passes.  v_output_alpha may be null (= a zero cotangent for that channel).
 * outputs_are_zero != 0: the six gradient arrays were already cleared by the caller (e.g. carved from one zero-filled
 * allocation), so the six memsets that otherwise precede the kernel are skipped. */
B200_API int b200_pack_records(int num_points, const float *xys, const float *pix_vels, const float *conics,
                               const float *colors, const float *opacities, void *packed, void *stream);
B200_API int b200_blend_forward_packed(unsigned img_height, unsigned img_width, unsigned block_width,
                                       unsigned n_blur_samples, const int32_t *gaussian_ids_sorted,
                                       const int32_t *tile_bins, const void *packed, float rolling_shutter_time,
                                       float exposure_time, const float *background, float *out_img, float *final_Ts,
                                       int32_t *final_idx, float *out_alpha, void *stream);
B200_API int b200_blend_backward_packed(int num_points, unsigned img_height, unsigned img_width, unsigned block_width,
                                        unsigned n_blur_samples, const int32_t *gaussian_ids_sorted,
                                        const int32_t *tile_bins, const void *packed, float rolling_shutter_time,
                                        float exposure_time, const float *background, const float *final_Ts,
                                        const int32_t *final_idx, const float *v_output, const float *v_output_alpha,
                                        float *v_xy, float *v_xy_abs, float *v_pix_vels, float *v_conic,
                                        float *v_colors, float *v_opacity, int outputs_are_zero, void *stream);

/* ---- blend (blur + rolling shutter), 3 channels ------------------------------------------
 * replaces rasterize_forward_tensor (bindings.h:110-127, bindings.cu:424-503; kernel forward.cu:306-456).
 * packed_ws: scratch of num_points * b200_packed_record_bytes() bytes, 16-byte aligned (filled here).
 * out_img (H,W,3), final_Ts (H,W,S) f32, final_idx (H,W,S) i32.  background: DEVICE float[3]. */
B200_API int b200_rasterize_forward(int num_points, unsigned img_height, unsigned img_width, unsigned block_width,
                           unsigned n_blur_samples, const int32_t *gaussian_ids_sorted, const int32_t *tile_bins,
                           const float *xys, const float *pix_vels, float rolling_shutter_time, float exposure_time,
                           const float *conics, const float *colors, const float *opacities, const float *background,
                           void *packed_ws, float *out_img, float *final_Ts, int32_t *final_idx, void *stream);

/* replaces rasterize_backward_tensor (bindings.h:168-197, bindings.cu:684-773; kernel backward.cu:143-369).
 * v_xy, v_xy_abs, v_pix_vels (N,2); v_conic, v_colors (N,3); v_opacity (N): overwritten (zeroed inside).
 * packed_ws as above (re-filled here, so forward and backward may use different scratch). */
B200_API int b200_rasterize_backward(int num_points, unsigned img_height, unsigned img_width, unsigned block_width,
                            unsigned n_blur_samples, const int32_t *gaussian_ids_sorted, const int32_t *tile_bins,
                            const float *xys, const float *pix_vels, float rolling_shutter_time, float exposure_time,
                            const float *conics, const float *colors, const float *opacities, const float *background,
                            const float *final_Ts, const int32_t *final_idx, const float *v_output,
                            const float *v_output_alpha, void *packed_ws, float *v_xy, float *v_xy_abs,
                            float *v_pix_vels, float *v_conic, float *v_colors, float *v_opacity, void *stream);

/* ---- fused pre/post-processing on RAW Splatfacto parameters ("next" row f-1 of SURVEY.md section 8) -----------------
 * Folds the caller-side glue of nerfstudio/models/splatfacto.py:816-856 into one kernel each way:
 *   forward : exp(log_scales), quats/|quats|, projection (b200_project_gaussians_forward semantics), "fast" SH colour
 *             + clamp(rgb + 0.5, 0) for Gaussians that survive the bbox test, sigmoid(opacity_logit) * compensation,
 *             and the packed blend record  ->  packed (N * b200_packed_record_bytes()), depths, radii, num_tiles_hit
 *             (exactly the inputs of b200_bin_cull_count / b200_blend_*_packed).
 *   backward: blend gradients (v_xy, v_pix_vel (N,2); v_conic, v_colors (N,3); v_opacity (N)) -> gradients of the raw
 *             parameters, every row written (zeros for culled Gaussians), plus optional camera gradients
 *             (g_lin_vel[3], g_ang_vel[3], g_viewmat[12]; overwritten).  Projection gradients are the clamp-aware
 *             "exact" ones (B200_PROJ_EXACT).
 * sh_dc (N,3) and sh_rest (N,K-1,3) are the two Splatfacto tensors (no torch.cat needed); sh_bases = K. */
B200_API int b200_fused_preprocess_forward(int num_points, const float *means, const float *log_scales,
                                           const float *quats, const float *opacity_logit, const float *sh_dc,
                                           const float *sh_rest, int sh_bases, int degrees_to_use, const float *viewmat,
                                           const float *cam_pos, const float *lin_vel, const float *ang_vel,
                                           float rolling_shutter_time, float exposure_time, float fx, float fy, float cx,
                                           float cy, unsigned img_height, unsigned img_width, unsigned block_width,
                                           float clip_thresh, void *packed, float *depths, int32_t *radii,
                                           int32_t *num_tiles_hit, void *stream);
/* The forward in two halves, for a trainer that projects and bins image k+1 while the SH coefficients of step k are still
 * being exchanged / updated (gsplat.dp.PipelinedTrainer): b200_fused_geometry_forward reads the geometry parameters only
 * and leaves the records' colours 0; b200_fused_colors_forward fills clamp(SH colour + 0.5, 0) into the records of the
 * Gaussians with radii > 0 afterwards (the binning does not read colours).  Together they write what
 * b200_fused_preprocess_forward writes; b200_fused_preprocess_backward serves both. */
B200_API int b200_fused_geometry_forward(int num_points, const float *means, const float *log_scales, const float *quats,
                                         const float *opacity_logit, const float *viewmat, const float *lin_vel,
                                         const float *ang_vel, float rolling_shutter_time, float exposure_time, float fx,
                                         float fy, float cx, float cy, unsigned img_height, unsigned img_width,
                                         unsigned block_width, float clip_thresh, void *packed, float *depths,
                                         int32_t *radii, int32_t *num_tiles_hit, void *stream);
B200_API int b200_fused_colors_forward(int num_points, const float *means, const float *sh_dc, const float *sh_rest,
                                       int sh_bases, int degrees_to_use, const float *cam_pos, const int32_t *radii,
                                       void *packed, void *stream);
B200_API int b200_fused_preprocess_backward(int num_points, const float *means, const float *log_scales,
                                            const float *quats, const float *opacity_logit, const float *sh_dc,
                                            const float *sh_rest, int sh_bases, int degrees_to_use, const float *viewmat,
                                            const float *cam_pos, const float *lin_vel, const float *ang_vel,
                                            float rolling_shutter_time, float exposure_time, float fx, float fy, float cx,
                                            float cy, unsigned img_height, unsigned img_width, unsigned block_width,
                                            float clip_thresh, const void *packed, const int32_t *radii,
                                            const float *v_xy, const float *v_pix_vel, const float *v_conic,
                                            const float *v_colors, const float *v_opacity, float *g_means,
                                            float *g_log_scales, float *g_quats, float *g_opacity_logit, float *g_sh_dc,
                                            float *g_sh_rest, float *g_lin_vel, float *g_ang_vel, float *g_viewmat,
                                            void *stream);

/* ---- N-channel blend (no blur), fp16 accumulators like the reference ----------------------
 * replaces nd_rasterize_forward_tensor / nd_rasterize_backward_tensor (bindings.h:129-166,
 * bindings.cu:506-682; kernels forward.cu:185-304, backward.cu:22-141). */
B200_API int b200_nd_rasterize_forward(int num_points, unsigned img_height, unsigned img_width, unsigned block_width,
                              unsigned channels, const int32_t *gaussian_ids_sorted, const int32_t *tile_bins,
                              const float *xys, const float *conics, const float *colors, const float *opacities,
                              const float *background, float *out_img, float *final_Ts, int32_t *final_idx,
                              void *stream);
B200_API int b200_nd_rasterize_backward(int num_points, unsigned img_height, unsigned img_width, unsigned block_width,
                               unsigned channels, const int32_t *gaussian_ids_sorted, const int32_t *tile_bins,
                               const float *xys, const float *conics, const float *colors, const float *opacities,
                               const float *background, const float *final_Ts, const int32_t *final_idx,
                               const float *v_output, const float *v_output_alpha, float *v_xy, float *v_xy_abs,
                               float *v_conic, float *v_colors, float *v_opacity, void *stream);

/* ---- photometric loss ("next" row f-3 of SURVEY section 8) -----------------------------------
 * Fused L1: loss[0] = mean |pred - target| over numel floats and, when grad is not null, grad = sign(pred - target) /
 * numel -- the cotangent rasterize_backward consumes -- in ONE pass; replaces the L1 term of the caller's loss
 * (nerfstudio/models/splatfacto.py:957, `torch.abs(gt_img - pred_img).mean()`) and its autograd chain.
 * ws: b200_l1_loss_ws_bytes() bytes of device scratch, 16-byte aligned, one per concurrent stream; pass
 * ws_is_zeroed = 1 if its first 4 bytes are zero (they are left zero on return), else they are cleared here.
 * Deterministic (block partial sums are added in block order). */
B200_API size_t b200_l1_loss_ws_bytes(void);
B200_API int b200_l1_loss(long long numel, const float *pred, const float *target, float *loss, float *grad, void *ws,
                          int ws_is_zeroed, void *stream);
/* The same with the caller's gamma step folded in (splatfacto.py:879-880 `clamp(rgb, max=1) ** (1 / gamma)` before the loss of
 * :957): pred_linear is the LINEAR render, loss[0] = mean |min(pred, 1)^(1/gamma) - target|, grad_linear = the cotangent w.r.t.
 * the linear image (torch's clamp / pow backward: (1/gamma) x^(1/gamma - 1) up to and including x = 1, 0 above, infinite at 0). */
B200_API int b200_l1_loss_gamma(long long numel, const float *pred_linear, const float *target, float gamma, float *loss,
                                float *grad_linear, void *ws, int ws_is_zeroed, void *stream);
/* The same against the dataset's uint8 image with the caller's ground-truth preparation folded in: uint8 -> float / 255
 * (splatfacto.py:900-910 get_gt_img), RGBA composited over `background` (3 DEVICE floats; :912-923 composite_with_background,
 * alpha * rgb + (1 - alpha) * background, unfused like torch), clamp(min = min_level) with min_level = min_rgb_level / 255
 * (:952-953; <= 0: off), optional mask (n_pixels floats: gt * mask, pred * mask, :957-964), L1 mean over 3 * n_pixels (:966).
 * channels = 3 or 4; gamma > 0 applies the gamma step to pred as in b200_l1_loss_gamma, <= 0: pred is used as it is;
 * target_out (3 * n_pixels floats or null) receives the prepared float target (what the SSIM term is taken against). */
B200_API int b200_l1_loss_u8(long long n_pixels, int channels, const float *pred, const unsigned char *target_u8,
                             const float *background, float min_level, float gamma, const float *mask, float *loss,
                             float *grad, float *target_out, void *ws, int ws_is_zeroed, void *stream);

/* SSIM term of the photometric loss, (1 - lambda) * L1 + lambda * (1 - SSIM) (nerfstudio/models/splatfacto.py:957-975;
 * SSIM = pytorch_msssim.SSIM(data_range=1, size_average=True, channel=C): 11-tap Gaussian window sigma 1.5, separable,
 * valid padding, K = (0.01, 0.03), mean over channels and the (H-10) x (W-10) valid positions).  Images are (H, W, C).
 *   b200_ssim_forward: ssim_out[0] = mean SSIM; if loss_out: loss_out[0] = (1 - lambda) * l1_loss[0] + lambda * (1 - SSIM)
 *       (l1_loss null = 0); if maps (b200_ssim_maps_bytes bytes): the three partial-derivative maps for the backward.
 *   b200_ssim_backward: grad = [v_scale[0] *] (add_scale * add_in + scale * dSSIM/dpred); add_in, v_scale optional.
 *       With add_in = the L1 cotangent, add_scale = 1 - lambda, scale = -lambda this is the photometric cotangent.
 * window11: optional HOST float[11], the normalised window taps (null = exp(-x^2 / 4.5) / sum evaluated in float here);
 *       SSIM of smooth images is sensitive to the last bits of the taps, so a binding that must reproduce
 *       pytorch_msssim passes the taps that library builds.
 * ws: b200_ssim_ws_bytes bytes, 16-byte aligned, ws_is_zeroed as for b200_l1_loss.  Deterministic. */
B200_API size_t b200_ssim_ws_bytes(unsigned img_height, unsigned img_width, unsigned channels);
B200_API size_t b200_ssim_maps_bytes(unsigned img_height, unsigned img_width, unsigned channels);
B200_API int b200_ssim_forward(unsigned img_height, unsigned img_width, unsigned channels, const float *pred,
                               const float *target, float *maps, float *ssim_out, float *loss_out, const float *l1_loss,
                               float ssim_lambda, const float *window11, void *ws, int ws_is_zeroed, void *stream);
B200_API int b200_ssim_backward(unsigned img_height, unsigned img_width, unsigned channels, const float *pred,
                                const float *target, const float *maps, float scale, const float *add_in,
                                float add_scale, const float *v_scale, const float *window11, float *grad,
                                void *stream);

/* ---- optimizer ("next" row f-2 of SURVEY section 8) ---------------------------------------------
 * One Adam update (no weight decay, no amsgrad: what nerfstudio/engine/optimizers.py:158-171 builds for every Splatfacto
 * group, eps = 1e-15) of `numel` consecutive floats of the flat {param, grad, exp_avg, exp_avg_sq} buffers:
 *   g' = grad_scale * grad;  m += (1 - beta1)(g' - m);  v = beta2 v + (1 - beta2) g'^2;
 *   param -= lr / (1 - beta1^step) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps);   grad = 0 if zero_grad.
 * `step` counts from 1; the hyper-parameters are doubles so that 1 - beta, the bias corrections and lr / (1 - beta1^step)
 * are formed in double exactly like the Python optimizer forms them.  The four pointers may point anywhere into equally aligned buffers (same offset each), so a
 * trainer can update slice by slice as gradient exchanges complete. */
B200_API int b200_adam_step(long long numel, float *param, float *grad, float *exp_avg, float *exp_avg_sq, int step,
                            double lr, double beta1, double beta2, double eps, double grad_scale, int zero_grad,
                            void *stream);

/* Device-resident optimizer state: the step count, the bias corrections and a per-step VETO live on the device, so a
 * trainer that never synchronises with the host (CUDA-graph replays, capacity-mode binning below) can still skip the
 * update of a step whose render turned out incomplete.  state: DEVICE, b200_adam_state_bytes() bytes, zeroed by the caller
 * before the first step; as int32 words: [2] optimizer steps applied, [3] current step vetoed (0/1), [4] steps prepared,
 * [5] steps vetoed, [6..15] indices (0-based, in `steps prepared` numbering) of the last 10 vetoed steps, slot = count % 10.
 *   b200_adam_prepare: once per step before the first update.  veto_flag: DEVICE int32 or null; non-zero = veto this step
 *       (a data-parallel caller reduces it with MAX across ranks first); it is cleared here.
 *   b200_adam_step_state: b200_adam_step with step / bias corrections / veto read from `state`; a vetoed step leaves
 *       param and moments untouched and still clears the gradient slice.  The slice must start 16-byte aligned. */
B200_API size_t b200_adam_state_bytes(void);
B200_API int b200_adam_prepare(void *state, int32_t *veto_flag, double lr, double beta1, double beta2, void *stream);
B200_API int b200_adam_step_state(long long numel, float *param, float *grad, float *exp_avg, float *exp_avg_sq,
                                  const void *state, double beta1, double beta2, double eps, double grad_scale,
                                  int zero_grad, void *stream);
/* The same update as a background kernel for a caller that runs it on a second stream beside latency-bound work: at most
 * ctas_per_sm (1..32) resident CTAs of 256 threads per SM loop over the slice, so kernels launched beside it always find
 * free thread slots instead of queueing behind a grid of thousands of short CTAs.  Same arithmetic, same result. */
B200_API int b200_adam_step_state_background(long long numel, float *param, float *grad, float *exp_avg, float *exp_avg_sq,
                                             const void *state, double beta1, double beta2, double eps, double grad_scale,
                                             int zero_grad, int ctas_per_sm, void *stream);

/* ---- rasterization without a host sync ---------------------------------------------------------------------------
 * The reference reads the intersection count back to size its lists (gsplat/utils.py:123-124 `.item()`); the two-phase
 * culled binning above reads the culled count.  In CAPACITY mode the caller sizes the lists from a running high-water
 * mark instead: b200_bin_cull_count with totals_host_pinned = NULL, then b200_bin_cull_emit_capacity.  The id list has
 * `capacity` slots: the real entries in the reference's order, then padding that belongs to no tile.
 * status: DEVICE int32[4]: [0] |= 1 if this image's entries did not fit (lists incomplete: discard the render -- veto the
 * optimizer step with b200_adam_prepare, grow the capacity, repeat the image), [1] entries of this image, [2] running
 * maximum, [3] the reference's num_intersects of this image.  b200_blend_forward_packed_status = b200_blend_forward_packed
 * plus the reference's empty-render branch (rasterize.py:136-144: background image, final_Ts = 0, alpha = 1) taken on
 * the device when status[3] < 1.  b200_set_record_colors patches the colours of packed records (everything the binning
 * reads is colour-independent, so a caller may bin before it has shaded). */
B200_API int b200_bin_cull_emit_capacity(int num_points, int capacity, const void *packed, const int32_t *radii,
                                         const int32_t *num_tiles_hit, unsigned img_height, unsigned img_width,
                                         unsigned block_width, unsigned n_blur_samples, float rolling_shutter_time,
                                         float exposure_time, const void *ws_g, void *ws_e, size_t ws_e_bytes,
                                         int32_t *gaussian_ids_sorted, int32_t *tile_bins, int32_t *status, void *stream);
B200_API int b200_blend_forward_packed_status(unsigned img_height, unsigned img_width, unsigned block_width,
                                              unsigned n_blur_samples, const int32_t *gaussian_ids_sorted,
                                              const int32_t *tile_bins, const void *packed, float rolling_shutter_time,
                                              float exposure_time, const float *background, const int32_t *status,
                                              float *out_img, float *final_Ts, int32_t *final_idx, float *out_alpha,
                                              void *stream);
B200_API int b200_set_record_colors(int num_points, const float *colors, void *packed, void *stream);

/* ---- adaptive density control on the flat buffers ("next" row f-2 of SURVEY section 8) -----------------------------------
 * replaces the Python mask / torch.cat / optimizer-state surgery of nerfstudio/models/splatfacto.py:352-622:
 *   b200_densify_accumulate  after_train (:408-434): grad_norm += |absgrad|, vis_counts += 1, max_2d = max(., radius /
 *       max_dim), max_dim = max(H, W), for the Gaussians visible in this image (radii > 0); first != 0: the first image since the last
 *       refinement initialises all three for EVERY Gaussian (norm, 1, 0) like the reference does.
 *   b200_densify_plan        refinement_after + cull_gaussians (:443-566): every split / duplicate / cull decision and,
 *       by two prefix sums, the final row of every survivor in the reference's order (kept originals, the children of
 *       split sample 0, 1, ..., the duplicates).  Negative split_screen_size / cull_scale_thresh / cull_screen_size switch
 *       that criterion off (the reference's step schedule, :462, :545, :548); do_densify = 0 culls only.
 *       counts (DEVICE int32[4]): new row count, splits, duplicates, culled originals -- the caller reads it once to size
 *       the new buffers.  ws: b200_densify_ws_bytes(num_points, n_split_samples) bytes, 256-byte aligned, kept for the gathers.
 *   b200_densify_gather      writes one field of the new buffers: field 0 copy; 1 means (split children are placed at
 *       mean + R(q/|q|) (exp(log_scale) * z), z = row sample * n_splits + rank of `randn`, the caller's
 *       (n_split_samples * n_splits, 3) normal draw, :574-582); 2 log-scales (split children shrink by 1.6, :596);
 *       3 optimizer moment (children start at zero, :384-399).  src (N, width) -> dst (new count, width). */
B200_API int b200_densify_accumulate(int num_points, const float *absgrad, const int32_t *radii, float max_dim, int first,
                                     float *grad_norm, float *vis_counts, float *max_2d, void *stream);
B200_API size_t b200_densify_ws_bytes(int num_points, int n_split_samples);
B200_API int b200_densify_plan(int num_points, const float *log_scales, const float *opacity_logit, const float *grad_norm,
                               const float *vis_counts, const float *max_2d, float half_max_dim, float densify_grad_thresh,
                               float densify_size_thresh, float split_screen_size, int n_split_samples,
                               float cull_alpha_thresh, float cull_scale_thresh, float cull_screen_size, int do_densify,
                               void *ws, size_t ws_bytes, int32_t *counts, void *stream);
B200_API int b200_densify_gather(int num_points, int n_split_samples, int field, int width, const float *src, float *dst,
                                 const void *ws, const int32_t *counts, const float *log_scales, const float *quats,
                                 const float *randn, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SPLAT_H */
