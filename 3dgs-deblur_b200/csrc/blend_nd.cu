// N-channel blend (no blur), forward + backward, with the reference's fp16 accumulators.
// Semantics of /root/reference/gsplat/gsplat/cuda/csrc/forward.cu:185-304 and backward.cu:22-141.
// Not used by Splatfacto (both of its calls pass 3 channels, splatfacto.py:867,890); kept for API
// completeness of rasterize_gaussians(colors.shape[-1] != 3) and deliberately simple: one thread per
// pixel, per-warp rectangle cull, records read straight from L2.
#include <cuda_fp16.h>

#include "blend_common.cuh"

namespace b200 {

struct NdParams {
    BlendGeom g;
    int C;
    const int32_t *ids_sorted;
    const int2 *tile_bins;
    const float2 *xys;
    const float *conics, *colors, *opac, *background;
};

__global__ void __launch_bounds__(BLEND_THREADS) nd_forward_kernel(NdParams p, float *__restrict__ out_img,
                                                                   float *__restrict__ final_Ts,
                                                                   int32_t *__restrict__ final_idx) {
    extern __shared__ __half s_acc[];
    const int tid = threadIdx.x;
    const int tile = blockIdx.x, tile_x = tile % p.g.tbx, tile_y = tile / p.g.tbx;
    int lx, ly;
    bool has_pixel;
    tile_pixel(p.g.bw, tid, lx, ly, has_pixel);
    const int j = tile_x * p.g.bw + lx, i = tile_y * p.g.bw + ly;
    const bool inside = has_pixel && i < p.g.H && j < p.g.W;
    if (!inside) return;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    __half *acc = s_acc + (size_t)tid * p.C;
    for (int c = 0; c < p.C; ++c) acc[c] = __float2half(0.f);
    const int2 range = p.tile_bins[tile];
    float T = 1.f;
    int last = 0;
    for (int k = range.x; k < range.y; ++k) {
        const int g = p.ids_sorted[k];
        const float2 xy = p.xys[g];
        const float ca = p.conics[3 * (size_t)g], cb = p.conics[3 * (size_t)g + 1], cc = p.conics[3 * (size_t)g + 2];
        const float dx = xy.x - px, dy = xy.y - py;
        const float sigma = 0.5f * (ca * dx * dx + cc * dy * dy) + cb * dx * dy;
        const float alpha = fminf(0.999f, p.opac[g] * __expf(-sigma));
        if (sigma < 0.f || alpha < 1.f / 255.f) continue;
        const float next_T = T * (1.f - alpha);
        if (next_T <= 1e-4f) break;
        const float vis = alpha * T;
        for (int c = 0; c < p.C; ++c) acc[c] = __hadd(acc[c], __float2half(p.colors[(size_t)p.C * g + c] * vis));
        T = next_T;
        last = k;
    }
    const size_t pix = (size_t)i * p.g.W + j;
    final_Ts[pix] = T;
    final_idx[pix] = last;
    for (int c = 0; c < p.C; ++c) out_img[pix * p.C + c] = __half2float(acc[c]) + T * p.background[c];
}

__global__ void __launch_bounds__(BLEND_THREADS) nd_backward_kernel(NdParams p, const float *__restrict__ final_Ts,
                                                                    const int32_t *__restrict__ final_idx,
                                                                    const float *__restrict__ v_out,
                                                                    const float *__restrict__ v_out_alpha,
                                                                    float *v_xy, float *v_xy_abs, float *v_conic,
                                                                    float *v_rgb, float *v_opac) {
    extern __shared__ __half s_acc[];
    const int tid = threadIdx.x;
    const int tile = blockIdx.x, tile_x = tile % p.g.tbx, tile_y = tile / p.g.tbx;
    int lx, ly;
    bool has_pixel;
    tile_pixel(p.g.bw, tid, lx, ly, has_pixel);
    const int j = tile_x * p.g.bw + lx, i = tile_y * p.g.bw + ly;
    const bool inside = has_pixel && i < p.g.H && j < p.g.W;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const size_t pix = inside ? (size_t)i * p.g.W + j : 0;
    const int2 range = p.tile_bins[tile];
    const float *vo = v_out + pix * p.C;
    const float voa = inside ? v_out_alpha[pix] : 0.f;
    const float T_final = inside ? final_Ts[pix] : 1.f;
    float T = T_final;
    __half *Sb = s_acc + (size_t)tid * p.C;
    for (int c = 0; c < p.C; ++c) Sb[c] = __float2half(0.f);
    const int bin_final = inside ? final_idx[pix] : 0;
    const int wmax = __reduce_max_sync(0xffffffffu, bin_final);
    for (int k = min(wmax, range.y) - 1; k >= range.x; --k) {  // strictly below bin_final (backward.cu:75-76)
        bool valid = inside && k < bin_final;
        const int g = p.ids_sorted[k];
        const float2 xy = p.xys[g];
        const float ca = p.conics[3 * (size_t)g], cb = p.conics[3 * (size_t)g + 1], cc = p.conics[3 * (size_t)g + 2];
        const float dx = xy.x - px, dy = xy.y - py;
        const float sigma = 0.5f * (ca * dx * dx + cc * dy * dy) + cb * dx * dy;
        const float opac = p.opac[g];
        const float vis = __expf(-sigma);
        const float alpha = fminf(0.99f, opac * vis);
        valid = valid && sigma >= 0.f && alpha >= 1.f / 255.f;
        if (!__any_sync(0xffffffffu, valid)) continue;
        float gx = 0.f, gy = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, vop = 0.f;
        if (valid) {
            const float ra = 1.f / (1.f - alpha);
            T *= ra;
            const float fac = alpha * T;
            float v_alpha = 0.f;
            for (int c = 0; c < p.C; ++c) {
                const float col = p.colors[(size_t)p.C * g + c];
                atomicAdd(v_rgb + (size_t)p.C * g + c, fac * vo[c]);  // per-lane, like backward.cu:105
                v_alpha += (col * T - __half2float(Sb[c]) * ra) * vo[c];
                v_alpha += -T_final * ra * p.background[c] * vo[c];
                Sb[c] = __hadd(Sb[c], __float2half(col * fac));
            }
            v_alpha += T_final * ra * voa;
            const float v_sigma = -opac * vis * v_alpha;
            c0 = 0.5f * v_sigma * dx * dx; c1 = v_sigma * dx * dy; c2 = 0.5f * v_sigma * dy * dy;
            gx = v_sigma * (ca * dx + cb * dy); gy = v_sigma * (cb * dx + cc * dy);
            vop = vis * v_alpha;
        }
        const float r0 = warp_sum(c0), r1 = warp_sum(c1), r2 = warp_sum(c2), r3 = warp_sum(gx), r4 = warp_sum(gy),
                    r5 = warp_sum(fabsf(gx)), r6 = warp_sum(fabsf(gy)), r7 = warp_sum(vop);
        if ((tid & 31) == 0) {
            atomicAdd(v_conic + 3 * (size_t)g, r0); atomicAdd(v_conic + 3 * (size_t)g + 1, r1); atomicAdd(v_conic + 3 * (size_t)g + 2, r2);
            atomicAdd(v_xy + 2 * (size_t)g, r3); atomicAdd(v_xy + 2 * (size_t)g + 1, r4);
            atomicAdd(v_xy_abs + 2 * (size_t)g, r5); atomicAdd(v_xy_abs + 2 * (size_t)g + 1, r6);
            atomicAdd(v_opac + g, r7);
        }
    }
}

static int nd_check(int n, unsigned H, unsigned W, unsigned bw, unsigned C) {
    B200_REQUIRE(n >= 1, "num_points must be >= 1");
    B200_REQUIRE(bw > 1 && bw <= 16, "block_width must be between 2 and 16");
    B200_REQUIRE(H > 0 && W > 0, "image size must be positive");
    B200_REQUIRE(C >= 1 && (size_t)C * BLEND_THREADS * sizeof(__half) <= 200 * 1024, "unsupported channel count %u", C);
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_nd_rasterize_forward(int num_points, unsigned img_height, unsigned img_width,
                                         unsigned block_width, unsigned channels, const int32_t *gaussian_ids_sorted,
                                         const int32_t *tile_bins, const float *xys, const float *conics,
                                         const float *colors, const float *opacities, const float *background,
                                         float *out_img, float *final_Ts, int32_t *final_idx, void *stream) {
    int rc = nd_check(num_points, img_height, img_width, block_width, channels);
    if (rc) return rc;
    B200_REQUIRE(gaussian_ids_sorted && tile_bins && xys && conics && colors && opacities && background && out_img &&
                     final_Ts && final_idx, "null pointer");
    NdParams p;
    p.g = BlendGeom{(int)img_height, (int)img_width, (int)block_width,
                    (int)((img_width + block_width - 1) / block_width),
                    (int)((img_height + block_width - 1) / block_width), 0.f, 0.f};
    p.C = (int)channels; p.ids_sorted = gaussian_ids_sorted; p.tile_bins = reinterpret_cast<const int2 *>(tile_bins);
    p.xys = reinterpret_cast<const float2 *>(xys); p.conics = conics; p.colors = colors; p.opac = opacities;
    p.background = background;
    const size_t smem = (size_t)channels * BLEND_THREADS * sizeof(__half);
    if (smem > 48 * 1024) B200_CUDA(cudaFuncSetAttribute(nd_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nd_forward_kernel<<<p.g.tbx * p.g.tby, BLEND_THREADS, smem, as_stream(stream)>>>(p, out_img, final_Ts, final_idx);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_nd_rasterize_backward(int num_points, unsigned img_height, unsigned img_width,
                                          unsigned block_width, unsigned channels,
                                          const int32_t *gaussian_ids_sorted, const int32_t *tile_bins,
                                          const float *xys, const float *conics, const float *colors,
                                          const float *opacities, const float *background, const float *final_Ts,
                                          const int32_t *final_idx, const float *v_output,
                                          const float *v_output_alpha, float *v_xy, float *v_xy_abs, float *v_conic,
                                          float *v_colors, float *v_opacity, void *stream) {
    int rc = nd_check(num_points, img_height, img_width, block_width, channels);
    if (rc) return rc;
    B200_REQUIRE(gaussian_ids_sorted && tile_bins && xys && conics && colors && opacities && background && final_Ts &&
                     final_idx && v_output && v_output_alpha && v_xy && v_xy_abs && v_conic && v_colors && v_opacity,
                 "null pointer");
    cudaStream_t st = as_stream(stream);
    const size_t n = (size_t)num_points;
    B200_CUDA(cudaMemsetAsync(v_xy, 0, n * 2 * sizeof(float), st));
    B200_CUDA(cudaMemsetAsync(v_xy_abs, 0, n * 2 * sizeof(float), st));
    B200_CUDA(cudaMemsetAsync(v_conic, 0, n * 3 * sizeof(float), st));
    B200_CUDA(cudaMemsetAsync(v_colors, 0, n * channels * sizeof(float), st));
    B200_CUDA(cudaMemsetAsync(v_opacity, 0, n * sizeof(float), st));
    NdParams p;
    p.g = BlendGeom{(int)img_height, (int)img_width, (int)block_width,
                    (int)((img_width + block_width - 1) / block_width),
                    (int)((img_height + block_width - 1) / block_width), 0.f, 0.f};
    p.C = (int)channels; p.ids_sorted = gaussian_ids_sorted; p.tile_bins = reinterpret_cast<const int2 *>(tile_bins);
    p.xys = reinterpret_cast<const float2 *>(xys); p.conics = conics; p.colors = colors; p.opac = opacities;
    p.background = background;
    const size_t smem = (size_t)channels * BLEND_THREADS * sizeof(__half);
    if (smem > 48 * 1024) B200_CUDA(cudaFuncSetAttribute(nd_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nd_backward_kernel<<<p.g.tbx * p.g.tby, BLEND_THREADS, smem, st>>>(p, final_Ts, final_idx, v_output, v_output_alpha,
                                                                     v_xy, v_xy_abs, v_conic, v_colors, v_opacity);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
