// Blur / rolling-shutter alpha-blend, backward.  Semantics of
// /root/reference/gsplat/gsplat/cuda/csrc/backward.cu:143-369, including the
// reference's 0.99 alpha clamp (backward.cu:275; the forward uses 0.999) and abs-grad per
// pixel-sample contribution (backward.cu:328-329).
//
// How it differs from the reference kernel:
//   * one back-to-front walk of the tile list for all S samples (per-sample T, running colour dot
//     product and last-contributor index live in registers) instead of S walks;
//   * per-Gaussian gradients are accumulated over the S samples in registers BEFORE the warp
//     reduction: S x fewer shuffles and atomics than backward.cu:334-365;
//   * the 13 per-Gaussian sums of a visit are reduced over the warp through a shared-memory
//     transposition (packed kernel: 13 STS + 4 LDS.128 + 7 packed adds + 1 shuffle) or a 16-shuffle
//     transposing butterfly (generic kernel) instead of 13 x 5 shuffles; either way each sum ends on
//     its own lane, so the 13 atomics issue as ONE predicated RED instruction instead of 13 serial
//     ones from lane 0;
//   * the same per-warp rectangle cull and TMA-bulk staged 64-byte records as the forward.
// Bound: FP32 issue, MUFU, LSU and L2 atomic throughput -- not HBM; see DESIGN.md section 4.
#include "blend_common.cuh"

namespace b200 {

struct BlendBwdParams {
    BlendGeom g;
    const int32_t *ids_sorted;
    const int2 *tile_bins;
    const PackedGaussian *packed;
    const float *background;
    const float *final_Ts;
    const int32_t *final_idx;
    const float *v_out;        // (H,W,3)
    const float *v_out_alpha;  // (H,W), or null for a zero cotangent
    float *v_xy, *v_xy_abs, *v_pix_vel, *v_conic, *v_rgb, *v_opac;
};

// Sum 16 per-lane values across the warp; lane L returns the total of value (L >> 1) & 15.
__device__ __forceinline__ float butterfly16(const float (&v)[16], int lane) {
    float w8[8], w4[4], w2[2], w1;
    bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
        w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    up = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = up ? w8[i] : w8[i + 4], keep = up ? w8[i + 4] : w8[i];
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    up = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = up ? w4[i] : w4[i + 2], keep = up ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    up = lane & 2;
    {
        const float send = up ? w2[0] : w2[1], keep = up ? w2[1] : w2[0];
        w1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
    return w1;
}

template <int S, int PPL>
__global__ void __launch_bounds__(BLEND_THREADS / PPL, PPL == 2 ? 5 : (PPL == 4 ? 6 : 3)) blend_backward_kernel(const BlendBwdParams p) {
    constexpr int NT = BLEND_THREADS / PPL;
    __shared__ __align__(128) PackedGaussian s_rec[BLEND_STAGES][BLEND_BATCH];
    __shared__ __align__(8) uint64_t s_bar[BLEND_STAGES];
    __shared__ int s_max[NT / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int tile_x = tile % p.g.tbx, tile_y = tile / p.g.tbx;
    const int2 range = p.tile_bins[tile];
    const float inv_s = 1.0f / (float)S;
    const float bg0 = __ldg(p.background), bg1 = __ldg(p.background + 1), bg2 = __ldg(p.background + 2);

    bool inside[PPL];
    float px[PPL], py[PPL], roll[PPL];
    // per-pixel cotangents and per-sample state (backward.cu:198-217)
    float vo[PPL][3];
    float Tm[PPL][S], D[PPL][S];  // Tm = T / S;  D = Tfinal/S * (v_alpha_out - bg.v_out) - sum_behind(fac * rgb.v_out)
    int bin_final[PPL][S];
    int my_max = -1;
#pragma unroll
    for (int q = 0; q < PPL; ++q) {
        int lx, ly;
        bool has_pixel;
        tile_pixel_ppl<PPL>(p.g.bw, tid, q, lx, ly, has_pixel);
        const int j = tile_x * p.g.bw + lx, i = tile_y * p.g.bw + ly;
        inside[q] = has_pixel && i < p.g.H && j < p.g.W;
        px[q] = (float)j + 0.5f; py[q] = (float)i + 0.5f;
        roll[q] = (float)((double)p.g.rs_time * ((double)(py[q] / (float)p.g.H) - 0.5));
        const size_t pix = inside[q] ? (size_t)i * p.g.W + j : 0;
        float voa = 0.f;
        vo[q][0] = vo[q][1] = vo[q][2] = 0.f;
        if (inside[q]) {
            vo[q][0] = p.v_out[3 * pix]; vo[q][1] = p.v_out[3 * pix + 1]; vo[q][2] = p.v_out[3 * pix + 2];
            voa = p.v_out_alpha ? p.v_out_alpha[pix] : 0.f;
        }
        const float bgdot = bg0 * vo[q][0] + bg1 * vo[q][1] + bg2 * vo[q][2];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            Tm[q][s] = inv_s; D[q][s] = 0.f; bin_final[q][s] = -1;
            if (inside[q]) {
                const float Tf = p.final_Ts[pix * S + s];
                Tm[q][s] = Tf * inv_s;
                D[q][s] = Tf * inv_s * (voa - bgdot);
                bin_final[q][s] = min(p.final_idx[pix * S + s], range.y - 1);  // batches only cover [range.x, range.y)
                my_max = max(my_max, bin_final[q][s]);
            }
        }
    }
    float blur[S];
#pragma unroll
    for (int s = 0; s < S; ++s) blur[s] = blur_offset<S>(s, p.g.exposure);

    const int wmax = __reduce_max_sync(0xffffffffu, my_max);  // last contributor over the warp's pixel-samples
    if (lane == 0) s_max[warp] = wmax;
    const WarpWindow win = warp_window<PPL>(inside, px, py, roll);

    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < BLEND_STAGES; ++st) mbar_init(&s_bar[st], 1);
        fence_mbar_init();
    }
    __syncthreads();
    int hi = -1;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) hi = max(hi, s_max[w]);
    const int total = hi - range.x + 1;  // entries hi, hi-1, ..., range.x
    const int nb = total > 0 ? (total + BLEND_BATCH - 1) / BLEND_BATCH : 0;

    // which output array / component this lane's butterfly result goes to
    float *my_dst = nullptr;
    int my_stride = 0;
    {
        const int k = lane >> 1;
        if ((lane & 1) == 0) {
            if (k < 3) { my_dst = p.v_rgb + k; my_stride = 3; }
            else if (k < 6) { my_dst = p.v_conic + (k - 3); my_stride = 3; }
            else if (k < 8) { my_dst = p.v_xy + (k - 6); my_stride = 2; }
            else if (k < 10) { my_dst = p.v_xy_abs + (k - 8); my_stride = 2; }
            else if (k < 12) { my_dst = p.v_pix_vel + (k - 10); my_stride = 2; }
            else if (k == 12) { my_dst = p.v_opac; my_stride = 1; }
        }
    }

    // stage entry t of batch b  <->  list index hi - b*BATCH - t  (t ascending = back to front)
    auto issue = [&](int b) {
        const int st = b & 1;
        const int top = hi - b * BLEND_BATCH;
        const int cnt = min(BLEND_BATCH, top - range.x + 1);
        if (tid == 0) mbar_arrive_expect_tx(&s_bar[st], (uint32_t)cnt * (uint32_t)sizeof(PackedGaussian));
#pragma unroll
        for (int r = 0; r < PPL; ++r) {
            const int e = tid + r * NT;
            if (e < cnt) {
                const int g = __ldg(p.ids_sorted + top - e);
                tma_bulk_g2s(&s_rec[st][e], p.packed + g, (uint32_t)sizeof(PackedGaussian), &s_bar[st]);
            }
        }
    };

    B200_COUNT_DECL;
    if (nb > 0) issue(0);
    for (int b = 0; b < nb; ++b) {
        const int st = b & 1;
        if (b + 1 < nb) issue(b + 1);
        mbar_wait(&s_bar[st], (uint32_t)((b >> 1) & 1));
        const int top = hi - b * BLEND_BATCH;
        const int cnt = min(BLEND_BATCH, top - range.x + 1);

        if (wmax >= range.x) {
            for (int c0 = 0; c0 < cnt; c0 += 32) {
                const int e = c0 + lane;
                const unsigned my_mask =
                    ((e < cnt) && (top - e <= wmax)) ? sample_mask_exact<S>(s_rec[st][e], win, p.g.exposure) : 0u;
                unsigned m = __ballot_sync(0xffffffffu, my_mask != 0u);
                if ((e < cnt) && (top - e <= wmax)) B200_COUNT(0, 1);
                while (m) {
                    const int src = __ffs(m) - 1;
                    const int k = c0 + src;
                    m &= m - 1;
                    const unsigned smask = __shfl_sync(0xffffffffu, my_mask, src);
                    if (lane == 0) { B200_COUNT(1, 1); B200_COUNT(2, __popc(smask)); }
                    const int idx = top - k;
                    const float4 A = *reinterpret_cast<const float4 *>(&s_rec[st][k].x);    // x y vx vy
                    const float4 Bq = *reinterpret_cast<const float4 *>(&s_rec[st][k].ca);  // a b c opac
                    const float4 C = *reinterpret_cast<const float4 *>(&s_rec[st][k].r);    // r g b thr
                    const float cut = C.w + 1e-4f;
                    float sxx = 0.f, sxy = 0.f, syy = 0.f, gxs = 0.f, gys = 0.f, gxa = 0.f, gya = 0.f, pvx = 0.f, pvy = 0.f,
                          vop = 0.f;
                    float cdot[PPL], facsum[PPL], dx0[PPL], dy0[PPL];
                    const SigmaEntry se = sigma_entry(A, Bq);
#pragma unroll
                    for (int q = 0; q < PPL; ++q) {
                        cdot[q] = C.x * vo[q][0] + C.y * vo[q][1] + C.z * vo[q][2];
                        facsum[q] = 0.f;
                        dx0[q] = A.x - px[q]; dy0[q] = A.y - py[q];
                    }
                    bool any = false;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        if (!(smask & (1u << s))) continue;  // warp-uniform
                        // The lane's PPL pixels are evaluated side by side in straight-line code (independent
                        // dependency chains for the scheduler to interleave); invalid pixels are masked, not branched.
                        float dx[PPL], dy[PPL], sigma[PPL], tau[PPL];
                        bool ok[PPL], some = false;
#pragma unroll
                        for (int q = 0; q < PPL; ++q) {
                            tau[q] = blur[s] + roll[q];
                            sigma_eval(se, px[q], py[q], dx0[q], dy0[q], tau[q], dx[q], dy[q], sigma[q]);
                            ok[q] = (idx <= bin_final[q][s]) && !(sigma[q] > cut || sigma[q] < 0.f);  // backward.cu:252-254,276
                            some |= ok[q];
                            if (inside[q]) B200_COUNT(3, 1);
                        }
                        if (!some) continue;
                        float vis[PPL], ov[PPL], alpha[PPL];
                        some = false;
#pragma unroll
                        for (int q = 0; q < PPL; ++q) {
                            vis[q] = exp_neg_approx(ok[q] ? sigma[q] : 0.f);
                            ov[q] = Bq.w * vis[q];
                            alpha[q] = fminf(0.99f, ov[q]);
                            ok[q] = ok[q] && !(alpha[q] < 1.f / 255.f);
                            some |= ok[q];
                            if (ok[q]) B200_COUNT(4, 1);
                        }
                        if (!some) continue;
                        any = true;
#pragma unroll
                        for (int q = 0; q < PPL; ++q) {
                            const float ra = rcp_approx(1.f - alpha[q]);
                            const float Tn = Tm[q][s] * ra;  // T / S of backward.cu:294-296
                            const float fac = alpha[q] * Tn;
                            const float v_alpha = Tn * cdot[q] + ra * D[q][s];  // backward.cu:303-311
                            // no zeroing when the clamp is active (backward.cu:317); masked pixels contribute exactly 0
                            const float v_sigma = ok[q] ? -ov[q] * v_alpha : 0.f;
                            Tm[q][s] = ok[q] ? Tn : Tm[q][s];
                            D[q][s] = ok[q] ? D[q][s] - fac * cdot[q] : D[q][s];  // running buffer, :313-315
                            facsum[q] += ok[q] ? fac : 0.f;
                            const float u = v_sigma * dx[q], w = v_sigma * dy[q];
                            sxx += u * dx[q]; sxy += u * dy[q]; syy += w * dy[q];
                            const float gx = Bq.x * u + Bq.y * w;
                            const float gy = Bq.y * u + Bq.z * w;
                            gxs += gx; gys += gy; gxa += fabsf(gx); gya += fabsf(gy);
                            pvx += gx * tau[q]; pvy += gy * tau[q];
                            vop += ok[q] ? vis[q] * v_alpha : 0.f;
                        }
                    }
                    float vr = 0.f, vg = 0.f, vb = 0.f;
#pragma unroll
                    for (int q = 0; q < PPL; ++q) {
                        vr += facsum[q] * vo[q][0]; vg += facsum[q] * vo[q][1]; vb += facsum[q] * vo[q][2];
                    }
                    if (!__any_sync(0xffffffffu, any)) continue;  // backward.cu:281-283
                    if (lane == 0) B200_COUNT(5, 1);
                    const float v[16] = {vr, vg, vb, 0.5f * sxx, sxy, 0.5f * syy, gxs, gys, gxa, gya, pvx, pvy, vop,
                                         0.f, 0.f, 0.f};
                    const float tot = butterfly16(v, lane);
                    if (my_dst && tot != 0.f) {
                        const int gid = s_rec[st][k].id;
                        atomicAdd(my_dst + (unsigned)gid * (unsigned)my_stride, tot);  // N * 3 < 2^32
                    }
                }
            }
        }
        __syncthreads();  // stage `st` is refilled two batches from now
    }
    B200_COUNT_FLUSH(8);
}


// ---- two pixels per lane, packed (PPL = 2, 16x16 tiles): the kernel the train step runs ------------------------------
// Same walk, staging, cull and reduction as blend_backward_kernel<S, 2>; the lane's two pixels (rows r and r + 4 of the
// warp's 8x8 block) live in the two halves of packed float pairs (f2, blend_common.cuh), so every FMA-pipe instruction of
// the sigma evaluation and of the gradient algebra issues once for both.  Per-sample structure:
//   * warp-uniform skips first: the exact cull's sample bit and idx <= (warp max of that sample's last contributor);
//   * sigma, exp, alpha for both pixels in straight-line code; pixels that fail a test are masked by zeroing `vis` and
//     `alpha` (then fac, v_sigma and every accumulated term are exactly 0 and T is kept by a select) -- one vote decides
//     whether the gradient algebra runs at all.
// Visit epilogue of the packed kernel: the 13 per-Gaussian sums are reduced over the warp through a shared-memory
// transposition (B200_BWD_SMEM_REDUCE=1, default) or with the 16-shuffle transposing butterfly above (=0, A/B build).
#ifndef B200_BWD_SMEM_REDUCE
#define B200_BWD_SMEM_REDUCE 1
#endif
// B200_BWD_T_SELECT=1 (A/B build): keep a masked pixel's transmittance with two selects instead of relying on
// rcp.approx(1) == 1 (true on sm_100a: tools/micro/rcp_check.cu; c2 backward 617 -> 610 us without them)
#ifndef B200_BWD_T_SELECT
#define B200_BWD_T_SELECT 0
#endif
constexpr int RED_VALUES = 13;  // rgb 3, conic 3, xy 2, |xy| 2, pixel velocity 2, opacity 1
constexpr int RED_STRIDE = 36;  // floats per row: 32 lanes + 4 of padding (rows stay 16-byte aligned, halves hit distinct banks)
#ifndef B200_BWD_MIN_CTAS
#define B200_BWD_MIN_CTAS 5  // 96 registers; A/B of 4 / 6 in DESIGN.md section 9
#endif
template <int S, bool VOTE>
__global__ void __launch_bounds__(128, B200_BWD_MIN_CTAS) blend_backward_kernel2(const BlendBwdParams p) {
    constexpr int NT = 128;
    __shared__ __align__(128) PackedGaussian s_rec[BLEND_STAGES][BLEND_BATCH];
    __shared__ __align__(8) uint64_t s_bar[BLEND_STAGES];
    __shared__ int s_max[NT / 32];
#if B200_BWD_SMEM_REDUCE
    __shared__ __align__(16) float s_red[NT / 32][RED_VALUES * RED_STRIDE];  // per-warp transposition buffer (visit epilogue)
#endif

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int tile_x = tile % p.g.tbx, tile_y = tile / p.g.tbx;
    const int2 range = p.tile_bins[tile];
    const float inv_s = 1.0f / (float)S;
    const float bg0 = __ldg(p.background), bg1 = __ldg(p.background + 1), bg2 = __ldg(p.background + 2);

    bool inside[2];
    float px[2], py[2], roll[2];
    float vo[2][3];
    float Tm_[2][S], D_[2][S];
    int bin_final[2][S];
    int smax[S];  // per-sample last contributor over the lane's pixels, then over the warp
#pragma unroll
    for (int s = 0; s < S; ++s) smax[s] = -1;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        int lx, ly;
        bool has_pixel;
        tile_pixel_ppl<2>(p.g.bw, tid, q, lx, ly, has_pixel);
        const int j = tile_x * p.g.bw + lx, i = tile_y * p.g.bw + ly;
        inside[q] = has_pixel && i < p.g.H && j < p.g.W;
        px[q] = (float)j + 0.5f; py[q] = (float)i + 0.5f;
        roll[q] = (float)((double)p.g.rs_time * ((double)(py[q] / (float)p.g.H) - 0.5));
        const size_t pix = inside[q] ? (size_t)i * p.g.W + j : 0;
        float voa = 0.f;
        vo[q][0] = vo[q][1] = vo[q][2] = 0.f;
        if (inside[q]) {
            vo[q][0] = p.v_out[3 * pix]; vo[q][1] = p.v_out[3 * pix + 1]; vo[q][2] = p.v_out[3 * pix + 2];
            voa = p.v_out_alpha ? p.v_out_alpha[pix] : 0.f;
        }
        const float bgdot = bg0 * vo[q][0] + bg1 * vo[q][1] + bg2 * vo[q][2];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            Tm_[q][s] = inv_s; D_[q][s] = 0.f; bin_final[q][s] = -1;
            if (inside[q]) {
                const float Tf = p.final_Ts[pix * S + s];
                Tm_[q][s] = Tf * inv_s;
                D_[q][s] = Tf * inv_s * (voa - bgdot);
                bin_final[q][s] = min(p.final_idx[pix * S + s], range.y - 1);
                smax[s] = max(smax[s], bin_final[q][s]);
            }
        }
    }
    const WarpWindow win = warp_window<2>(inside, px, py, roll);
    int wmax = -1;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        smax[s] = __reduce_max_sync(0xffffffffu, smax[s]);
        wmax = max(wmax, smax[s]);
    }
    if (lane == 0) s_max[warp] = wmax;
    // packed per-pixel state
    f2 Tm[S], D[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { Tm[s] = f2_make(Tm_[0][s], Tm_[1][s]); D[s] = f2_make(D_[0][s], D_[1][s]); }
    const f2 PX = f2_make(px[0], px[1]), PY = f2_make(py[0], py[1]), ROLL = f2_make(roll[0], roll[1]);
    const f2 VO0 = f2_make(vo[0][0], vo[1][0]), VO1 = f2_make(vo[0][1], vo[1][1]), VO2 = f2_make(vo[0][2], vo[1][2]);
    float blur[S];
#pragma unroll
    for (int s = 0; s < S; ++s) blur[s] = blur_offset<S>(s, p.g.exposure);

    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < BLEND_STAGES; ++st) mbar_init(&s_bar[st], 1);
        fence_mbar_init();
    }
    __syncthreads();
    int hi = -1;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) hi = max(hi, s_max[w]);
    const int total = hi - range.x + 1;
    const int nb = total > 0 ? (total + BLEND_BATCH - 1) / BLEND_BATCH : 0;

    float *my_dst = nullptr;
    int my_stride = 0;
    {
        const int k = lane >> 1;
        if ((lane & 1) == 0) {
            if (k < 3) { my_dst = p.v_rgb + k; my_stride = 3; }
            else if (k < 6) { my_dst = p.v_conic + (k - 3); my_stride = 3; }
            else if (k < 8) { my_dst = p.v_xy + (k - 6); my_stride = 2; }
            else if (k < 10) { my_dst = p.v_xy_abs + (k - 8); my_stride = 2; }
            else if (k < 12) { my_dst = p.v_pix_vel + (k - 10); my_stride = 2; }
            else if (k == 12) { my_dst = p.v_opac; my_stride = 1; }
        }
    }

#if B200_BWD_SMEM_REDUCE
    float *const red_row = &s_red[warp][lane];
    // lanes 26..31 own no sum (my_dst is null): they re-read row 12 (a broadcast) instead of branching
    const ulonglong2 *const red_half =
        reinterpret_cast<const ulonglong2 *>(&s_red[warp][min(lane >> 1, RED_VALUES - 1) * RED_STRIDE + (lane & 1) * 16]);
#endif

    auto issue = [&](int b) {
        const int st = b & 1;
        const int top = hi - b * BLEND_BATCH;
        const int cnt = min(BLEND_BATCH, top - range.x + 1);
        if (tid == 0) mbar_arrive_expect_tx(&s_bar[st], (uint32_t)cnt * (uint32_t)sizeof(PackedGaussian));
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int e = tid + r * NT;
            if (e < cnt) {
                const int g = __ldg(p.ids_sorted + top - e);
                tma_bulk_g2s(&s_rec[st][e], p.packed + g, (uint32_t)sizeof(PackedGaussian), &s_bar[st]);
            }
        }
    };

    B200_COUNT_DECL;
    if (nb > 0) issue(0);
    for (int b = 0; b < nb; ++b) {
        const int st = b & 1;
        if (b + 1 < nb) issue(b + 1);
        mbar_wait(&s_bar[st], (uint32_t)((b >> 1) & 1));
        const int top = hi - b * BLEND_BATCH;
        const int cnt = min(BLEND_BATCH, top - range.x + 1);

        if (wmax >= range.x) {
            for (int c0 = 0; c0 < cnt; c0 += 32) {
                const int e = c0 + lane;
                const unsigned my_mask =
                    ((e < cnt) && (top - e <= wmax)) ? sample_mask_exact<S>(s_rec[st][e], win, p.g.exposure) : 0u;
                unsigned m = __ballot_sync(0xffffffffu, my_mask != 0u);
                if ((e < cnt) && (top - e <= wmax)) B200_COUNT(0, 1);
                while (m) {
                    const int src = __ffs(m) - 1;
                    const int k = c0 + src;
                    m &= m - 1;
                    const unsigned smask = __shfl_sync(0xffffffffu, my_mask, src);
                    if (lane == 0) { B200_COUNT(1, 1); B200_COUNT(2, __popc(smask)); }
                    const int idx = top - k;
                    const float4 A = *reinterpret_cast<const float4 *>(&s_rec[st][k].x);    // x y vx vy
                    const float4 Bq = *reinterpret_cast<const float4 *>(&s_rec[st][k].ca);  // a b c opac
                    const float4 C = *reinterpret_cast<const float4 *>(&s_rec[st][k].r);    // r g b thr
                    const float cut = C.w + 1e-4f;
                    const f2 VX = f2_splat(A.z), VY = f2_splat(A.w);
                    const f2 CA = f2_splat(Bq.x), CB = f2_splat(Bq.y), CC = f2_splat(Bq.z);
                    const f2 HA = f2_splat(0.5f * Bq.x), HC = f2_splat(0.5f * Bq.z), NOP = f2_splat(-Bq.w);
                    const f2 dx0 = f2_sub(f2_splat(A.x), PX), dy0 = f2_sub(f2_splat(A.y), PY);
                    const f2 cdot = f2_fma(f2_splat(C.x), VO0, f2_fma(f2_splat(C.y), VO1, f2_mul(f2_splat(C.z), VO2)));
                    const f2 zero = f2_splat(0.f);
                    const f2 ncdot = f2_sub(zero, cdot);  // (negations hoisted out of the sample blocks: exact)
                    f2 sxx = zero, sxy = zero, syy = zero, gxs = zero, gys = zero, pvx = zero, pvy = zero, vop = zero, facsum = zero;
                    float gxa = 0.f, gya = 0.f;
                    bool any = false;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        if (!(smask & (1u << s)) || idx > smax[s]) continue;  // warp-uniform
                        const f2 tau = f2_add(f2_splat(blur[s]), ROLL);
                        const f2 dx = f2_fma(tau, VX, dx0), dy = f2_fma(tau, VY, dy0);
                        const f2 u0 = f2_fma(HA, dx, f2_mul(CB, dy));
                        const f2 sigma = f2_fma(dx, u0, f2_mul(f2_mul(HC, dy), dy));
                        const float sg0 = f2_lo(sigma), sg1 = f2_hi(sigma);
                        bool ok0 = (idx <= bin_final[0][s]) && !(sg0 > cut || sg0 < 0.f);  // backward.cu:252-254,276
                        bool ok1 = (idx <= bin_final[1][s]) && !(sg1 > cut || sg1 < 0.f);
                        if (inside[0]) B200_COUNT(3, 1);
                        if (inside[1]) B200_COUNT(3, 1);
                        // exp(-sigma) = 2^(-sigma log2 e) (exp_neg_approx's own form, the product packed): masked pixels evaluate 2^0
                        const f2 ex = f2_mul(sigma, f2_splat(-1.4426950408889634f));
                        float v0, v1;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(v0) : "f"(ok0 ? f2_lo(ex) : 0.f));
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(v1) : "f"(ok1 ? f2_hi(ex) : 0.f));
                        float a0 = fminf(0.99f, Bq.w * v0), a1 = fminf(0.99f, Bq.w * v1);
                        ok0 = ok0 && !(a0 < 1.f / 255.f);
                        ok1 = ok1 && !(a1 < 1.f / 255.f);
                        if (VOTE) {
                            if (!__any_sync(0xffffffffu, ok0 || ok1)) continue;
                            any = true;
                        }
                        if (ok0) B200_COUNT(4, 1);
                        if (ok1) B200_COUNT(4, 1);
                        // masked pixels: vis = alpha = 0  =>  ra = 1, fac = 0, v_sigma = 0, every accumulated term exactly 0
                        const f2 vis = f2_make(ok0 ? v0 : 0.f, ok1 ? v1 : 0.f);
                        const f2 alpha = f2_make(ok0 ? a0 : 0.f, ok1 ? a1 : 0.f);
                        const f2 nov = f2_mul(NOP, vis);  // -opacity * exp(-sigma)
                        const f2 om = f2_sub(f2_splat(1.f), alpha);
                        const f2 ra = f2_make(rcp_approx(f2_lo(om)), rcp_approx(f2_hi(om)));
                        const f2 Tn = f2_mul(Tm[s], ra);                 // T / S of backward.cu:294-296
                        const f2 fac = f2_mul(alpha, Tn);
                        const f2 v_alpha = f2_fma(Tn, cdot, f2_mul(ra, D[s]));  // backward.cu:303-311
                        // no zeroing when the clamp is active (backward.cu:317)
                        const f2 v_sigma = f2_mul(nov, v_alpha);
#if B200_BWD_T_SELECT
                        Tm[s] = f2_make(ok0 ? f2_lo(Tn) : f2_lo(Tm[s]), ok1 ? f2_hi(Tn) : f2_hi(Tm[s]));
#else
                        Tm[s] = Tn;  // masked pixels: alpha = 0, ra = rcp(1) = 1 exactly (tools/micro/rcp_check.cu), Tn == Tm bit for bit
#endif
                        D[s] = f2_fma(fac, ncdot, D[s]);                  // running buffer, :313-315
                        facsum = f2_add(facsum, fac);
                        const f2 u = f2_mul(v_sigma, dx), w = f2_mul(v_sigma, dy);
                        sxx = f2_fma(u, dx, sxx); sxy = f2_fma(u, dy, sxy); syy = f2_fma(w, dy, syy);
                        const f2 gx = f2_fma(CA, u, f2_mul(CB, w));
                        const f2 gy = f2_fma(CB, u, f2_mul(CC, w));
                        gxs = f2_add(gxs, gx); gys = f2_add(gys, gy);
                        gxa += fabsf(f2_lo(gx)); gxa += fabsf(f2_hi(gx));
                        gya += fabsf(f2_lo(gy)); gya += fabsf(f2_hi(gy));
                        pvx = f2_fma(gx, tau, pvx); pvy = f2_fma(gy, tau, pvy);
                        vop = f2_fma(vis, v_alpha, vop);
                    }
                    if (VOTE && !__any_sync(0xffffffffu, any)) continue;  // backward.cu:281-283
                    // (without the votes a visit whose every pixel failed the alpha test reduces 13 exact zeros and the
                    // `tot != 0` guard below drops the atomics: same memory effect as the reference's skip)
                    if (lane == 0) B200_COUNT(5, 1);
#if B200_BWD_SMEM_REDUCE
                    // lane l parks value j at row j, column l (bank 4 j + l: conflict free); lanes 2 j and 2 j + 1 then add up
                    // the two halves of row j with 4 LDS.128 + 7 packed adds each and meet in one shuffle.  Same owner lane
                    // (2 j) per sum as the butterfly, 13 STS + 4 LDS + 1 SHFL instead of 16 SHFL + 27 selects.
                    red_row[0 * RED_STRIDE] = f2_sum(f2_mul(facsum, VO0));
                    red_row[1 * RED_STRIDE] = f2_sum(f2_mul(facsum, VO1));
                    red_row[2 * RED_STRIDE] = f2_sum(f2_mul(facsum, VO2));
                    red_row[3 * RED_STRIDE] = 0.5f * f2_sum(sxx);
                    red_row[4 * RED_STRIDE] = f2_sum(sxy);
                    red_row[5 * RED_STRIDE] = 0.5f * f2_sum(syy);
                    red_row[6 * RED_STRIDE] = f2_sum(gxs);
                    red_row[7 * RED_STRIDE] = f2_sum(gys);
                    red_row[8 * RED_STRIDE] = gxa;
                    red_row[9 * RED_STRIDE] = gya;
                    red_row[10 * RED_STRIDE] = f2_sum(pvx);
                    red_row[11 * RED_STRIDE] = f2_sum(pvy);
                    red_row[12 * RED_STRIDE] = f2_sum(vop);
                    __syncwarp();
                    float tot;
                    {
                        const ulonglong2 q0 = red_half[0], q1 = red_half[1], q2 = red_half[2], q3 = red_half[3];
                        const f2 a0 = f2_add(f2{q0.x}, f2{q0.y}), a1 = f2_add(f2{q1.x}, f2{q1.y});
                        const f2 a2 = f2_add(f2{q2.x}, f2{q2.y}), a3 = f2_add(f2{q3.x}, f2{q3.y});
                        tot = f2_sum(f2_add(f2_add(a0, a1), f2_add(a2, a3)));
                    }
                    tot += __shfl_xor_sync(0xffffffffu, tot, 1);
                    __syncwarp();  // every lane has read its half before the next visit overwrites the rows
#else
                    const float v[16] = {f2_sum(f2_mul(facsum, VO0)), f2_sum(f2_mul(facsum, VO1)), f2_sum(f2_mul(facsum, VO2)),
                                         0.5f * f2_sum(sxx), f2_sum(sxy), 0.5f * f2_sum(syy), f2_sum(gxs), f2_sum(gys), gxa, gya,
                                         f2_sum(pvx), f2_sum(pvy), f2_sum(vop), 0.f, 0.f, 0.f};
                    const float tot = butterfly16(v, lane);
#endif
                    if (my_dst && tot != 0.f) {
                        const int gid = s_rec[st][k].id;
                        atomicAdd(my_dst + (unsigned)gid * (unsigned)my_stride, tot);  // N * 3 < 2^32
                    }
                }
            }
        }
        __syncthreads();  // stage `st` is refilled two batches from now
    }
    B200_COUNT_FLUSH(8);
}

template <int S>
static int launch_bwd(const BlendBwdParams &p, cudaStream_t st) {
    if (p.g.bw == 16 && blend_pixels_per_lane(true) == 4)  // experimental (B200_BLEND_PPL_BWD=4)
        blend_backward_kernel<S, 4><<<p.g.tbx * p.g.tby, BLEND_THREADS / 4, 0, st>>>(p);
    else if (p.g.bw == 16 && blend_pixels_per_lane(true) == 2 && blend_packed()) {
        // per-sample votes ("does any lane hold a contributing pixel", then skip the gradient algebra): without rolling
        // shutter the exact cull leaves 4 % of the entered sample blocks empty and the vote + branch costs more than the
        // masked arithmetic (c2: 646 -> 617 us); with it the warp's time window widens the cull and the skip pays (c4:
        // 4988 us with, 5573 without) -- profiles/r2x_blend_ab.txt
        if (B200_SAMPLE_VOTE || p.g.rs_time != 0.f)
            blend_backward_kernel2<S, true><<<p.g.tbx * p.g.tby, BLEND_THREADS / 2, 0, st>>>(p);
        else
            blend_backward_kernel2<S, false><<<p.g.tbx * p.g.tby, BLEND_THREADS / 2, 0, st>>>(p);
    }
    else if (p.g.bw == 16 && blend_pixels_per_lane(true) == 2)  // B200_BLEND_PACKED=0: the scalar two-pixel kernel (A/B)
        blend_backward_kernel<S, 2><<<p.g.tbx * p.g.tby, BLEND_THREADS / 2, 0, st>>>(p);
    else
        blend_backward_kernel<S, 1><<<p.g.tbx * p.g.tby, BLEND_THREADS, 0, st>>>(p);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

}  // namespace b200

using namespace b200;

static int run_blend_backward(int num_points, unsigned img_height, unsigned img_width, unsigned block_width,
                              unsigned n_blur_samples, const int32_t *gaussian_ids_sorted, const int32_t *tile_bins,
                              const void *packed, float rolling_shutter_time, float exposure_time,
                              const float *background, const float *final_Ts, const int32_t *final_idx,
                              const float *v_output, const float *v_output_alpha, float *v_xy, float *v_xy_abs,
                              float *v_pix_vels, float *v_conic, float *v_colors, float *v_opacity, bool outputs_are_zero,
                              cudaStream_t st) {
    B200_REQUIRE(n_blur_samples > 0 && n_blur_samples <= B200_MAX_BLUR_SAMPLES, "unsupported blur size");
    B200_REQUIRE(num_points >= 1, "num_points must be >= 1");
    B200_REQUIRE(block_width > 1 && block_width <= 16, "block_width must be between 2 and 16");
    B200_REQUIRE(img_height > 0 && img_width > 0, "image size must be positive");
    B200_REQUIRE(tile_bins && background && packed && aligned16(packed), "null / misaligned input pointer");
    B200_REQUIRE(final_Ts && final_idx && v_output, "null saved / cotangent pointer");
    B200_REQUIRE(v_xy && v_xy_abs && v_pix_vels && v_conic && v_colors && v_opacity, "null output pointer");
    const size_t n = (size_t)num_points;
    if (!outputs_are_zero) {
        B200_CUDA(cudaMemsetAsync(v_xy, 0, n * 2 * sizeof(float), st));
        B200_CUDA(cudaMemsetAsync(v_xy_abs, 0, n * 2 * sizeof(float), st));
        B200_CUDA(cudaMemsetAsync(v_pix_vels, 0, n * 2 * sizeof(float), st));
        B200_CUDA(cudaMemsetAsync(v_conic, 0, n * 3 * sizeof(float), st));
        B200_CUDA(cudaMemsetAsync(v_colors, 0, n * 3 * sizeof(float), st));
        B200_CUDA(cudaMemsetAsync(v_opacity, 0, n * sizeof(float), st));
    }
    BlendBwdParams p;
    p.g = BlendGeom{(int)img_height, (int)img_width, (int)block_width,
                    (int)((img_width + block_width - 1) / block_width),
                    (int)((img_height + block_width - 1) / block_width), rolling_shutter_time, exposure_time};
    p.ids_sorted = gaussian_ids_sorted;
    p.tile_bins = reinterpret_cast<const int2 *>(tile_bins);
    p.packed = reinterpret_cast<const PackedGaussian *>(packed);
    p.background = background;
    p.final_Ts = final_Ts; p.final_idx = final_idx; p.v_out = v_output; p.v_out_alpha = v_output_alpha;
    p.v_xy = v_xy; p.v_xy_abs = v_xy_abs; p.v_pix_vel = v_pix_vels; p.v_conic = v_conic; p.v_rgb = v_colors;
    p.v_opac = v_opacity;
    switch (n_blur_samples) {
        case 1: return launch_bwd<1>(p, st);
        case 2: return launch_bwd<2>(p, st);
        case 3: return launch_bwd<3>(p, st);
        case 4: return launch_bwd<4>(p, st);
        case 5: return launch_bwd<5>(p, st);
        case 6: return launch_bwd<6>(p, st);
        case 7: return launch_bwd<7>(p, st);
        case 8: return launch_bwd<8>(p, st);
        case 9: return launch_bwd<9>(p, st);
        default: return launch_bwd<10>(p, st);
    }
}

extern "C" int b200_blend_backward_packed(int num_points, unsigned img_height, unsigned img_width, unsigned block_width,
                                          unsigned n_blur_samples, const int32_t *gaussian_ids_sorted,
                                          const int32_t *tile_bins, const void *packed, float rolling_shutter_time,
                                          float exposure_time, const float *background, const float *final_Ts,
                                          const int32_t *final_idx, const float *v_output, const float *v_output_alpha,
                                          float *v_xy, float *v_xy_abs, float *v_pix_vels, float *v_conic,
                                          float *v_colors, float *v_opacity, int outputs_are_zero, void *stream) {
    return run_blend_backward(num_points, img_height, img_width, block_width, n_blur_samples, gaussian_ids_sorted,
                              tile_bins, packed, rolling_shutter_time, exposure_time, background, final_Ts, final_idx,
                              v_output, v_output_alpha, v_xy, v_xy_abs, v_pix_vels, v_conic, v_colors, v_opacity,
                              outputs_are_zero != 0, as_stream(stream));
}

extern "C" int b200_rasterize_backward(int num_points, unsigned img_height, unsigned img_width, unsigned block_width,
                                       unsigned n_blur_samples, const int32_t *gaussian_ids_sorted,
                                       const int32_t *tile_bins, const float *xys, const float *pix_vels,
                                       float rolling_shutter_time, float exposure_time, const float *conics,
                                       const float *colors, const float *opacities, const float *background,
                                       const float *final_Ts, const int32_t *final_idx, const float *v_output,
                                       const float *v_output_alpha, void *packed_ws, float *v_xy, float *v_xy_abs,
                                       float *v_pix_vels, float *v_conic, float *v_colors, float *v_opacity,
                                       void *stream) {
    B200_REQUIRE(num_points >= 1, "num_points must be >= 1");
    B200_REQUIRE(gaussian_ids_sorted && tile_bins && xys && pix_vels && conics && colors && opacities && background,
                 "null input pointer");
    B200_REQUIRE(packed_ws && aligned16(packed_ws), "packed_ws must be a 16-byte aligned scratch buffer");
    cudaStream_t st = as_stream(stream);
    int rc = launch_pack(num_points, xys, pix_vels, conics, colors, opacities, packed_ws, st);
    if (rc) return rc;
    return run_blend_backward(num_points, img_height, img_width, block_width, n_blur_samples, gaussian_ids_sorted,
                              tile_bins, packed_ws, rolling_shutter_time, exposure_time, background, final_Ts, final_idx,
                              v_output, v_output_alpha, v_xy, v_xy_abs, v_pix_vels, v_conic, v_colors, v_opacity, false, st);
}
