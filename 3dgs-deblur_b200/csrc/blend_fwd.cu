// Blur / rolling-shutter alpha-blend, forward.  Semantics of
// /root/reference/gsplat/gsplat/cuda/csrc/forward.cu:306-456.
//
// Differences in *how* (not what) from the reference kernel:
//   * the tile list is walked ONCE: every staged Gaussian is evaluated for all S blur samples with the
//     S transmittances held in registers (the reference re-stages the whole list per sample);
//   * Gaussians are staged as one 64-byte packed record each, gathered into shared memory with
//     per-entry TMA bulk copies (cp.async.bulk + mbarrier), double buffered against the blend;
//   * each warp owns a compact 8x4 pixel block and first culls the staged batch against that
//     rectangle (one Gaussian per lane + ballot), so only Gaussians that can reach alpha >= 1/255
//     somewhere in the warp's pixels / sample times are evaluated.  The cull is conservative and the
//     exact per-pixel tests of the reference are kept, so results do not change.
// Bound: FP32 issue + MUFU.EX2 (hundreds of flop per staged byte), not HBM -- see DESIGN.md.
#include "blend_common.cuh"

namespace b200 {

struct BlendFwdParams {
    BlendGeom g;
    const int32_t *ids_sorted;
    const int2 *tile_bins;
    const PackedGaussian *packed;
    const float *background;  // device float[3]
    float *out_img;           // (H,W,3)
    float *final_Ts;          // (H,W,S)
    float *out_alpha;         // (H,W) or null: 1 - mean_s final_T, what rasterize.py:161-163 derives with two torch ops
    int32_t *final_idx;       // (H,W,S)
    const int32_t *ref_isect;  // DEVICE int32 or null: the reference's num_intersects of this image, when the host never
                               // learned it (capacity-mode binning); < 1 selects the reference's empty-render branch
};

// What the reference returns when nothing intersects a tile (rasterize.py:136-144): the background colour and all-zero
// final_Ts / final_idx, hence alpha = 1.  Only reachable in the capacity mode, where the count lives on the device.
template <int S>
__device__ __forceinline__ void write_empty_render(const BlendFwdParams &p, bool inside, int i, int j) {
    if (!inside) return;
    const size_t pix = (size_t)i * p.g.W + j;
#pragma unroll
    for (int s = 0; s < S; ++s) { p.final_Ts[pix * S + s] = 0.f; p.final_idx[pix * S + s] = 0; }
    p.out_img[3 * pix] = __ldg(p.background); p.out_img[3 * pix + 1] = __ldg(p.background + 1);
    p.out_img[3 * pix + 2] = __ldg(p.background + 2);
    if (p.out_alpha) p.out_alpha[pix] = 1.0f;
}

template <int S, int PPL>
__global__ void __launch_bounds__(BLEND_THREADS / PPL, PPL == 2 ? 6 : 1) blend_forward_kernel(const BlendFwdParams p) {
    constexpr int NT = BLEND_THREADS / PPL;  // threads per tile
    __shared__ __align__(128) PackedGaussian s_rec[BLEND_STAGES][BLEND_BATCH];
    __shared__ __align__(8) uint64_t s_bar[BLEND_STAGES];

    const int tid = threadIdx.x, lane = tid & 31;
    const int tile = blockIdx.x;
    const int tile_x = tile % p.g.tbx, tile_y = tile / p.g.tbx;
    bool inside[PPL];
    float px[PPL], py[PPL], roll[PPL];
    int pi[PPL], pj[PPL];
#pragma unroll
    for (int q = 0; q < PPL; ++q) {
        int lx, ly;
        bool has_pixel;
        tile_pixel_ppl<PPL>(p.g.bw, tid, q, lx, ly, has_pixel);
        pj[q] = tile_x * p.g.bw + lx; pi[q] = tile_y * p.g.bw + ly;
        inside[q] = has_pixel && pi[q] < p.g.H && pj[q] < p.g.W;
        px[q] = (float)pj[q] + 0.5f; py[q] = (float)pi[q] + 0.5f;
        roll[q] = (float)((double)p.g.rs_time * ((double)(py[q] / (float)p.g.H) - 0.5));  // forward.cu:360
    }
    if (p.ref_isect && __ldg(p.ref_isect) < 1) {  // block-uniform
#pragma unroll
        for (int q = 0; q < PPL; ++q) write_empty_render<S>(p, inside[q], pi[q], pj[q]);
        return;
    }
    float blur[S];  // blur_rel of forward.cu:363 without the roll part; tau = blur[s] + roll[q]
#pragma unroll
    for (int s = 0; s < S; ++s) blur[s] = blur_offset<S>(s, p.g.exposure);

    const WarpWindow win = warp_window<PPL>(inside, px, py, roll);

    const int2 range = p.tile_bins[tile];
    const int total = range.y - range.x;
    const int nb = (total + BLEND_BATCH - 1) / BLEND_BATCH;

    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < BLEND_STAGES; ++st) mbar_init(&s_bar[st], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int b) {
        const int st = b & 1;
        const int start = range.x + b * BLEND_BATCH;
        const int cnt = min(BLEND_BATCH, range.y - start);
        if (tid == 0) mbar_arrive_expect_tx(&s_bar[st], (uint32_t)cnt * (uint32_t)sizeof(PackedGaussian));
#pragma unroll
        for (int r = 0; r < PPL; ++r) {
            const int e = tid + r * NT;
            if (e < cnt) {
                const int g = __ldg(p.ids_sorted + start + e);
                tma_bulk_g2s(&s_rec[st][e], p.packed + g, (uint32_t)sizeof(PackedGaussian), &s_bar[st]);
            }
        }
    };

    float T[PPL][S];
    int last[PPL][S];
    float acc[PPL][3];
    unsigned alive[PPL];
#pragma unroll
    for (int q = 0; q < PPL; ++q) {
#pragma unroll
        for (int s = 0; s < S; ++s) { T[q][s] = 1.f; last[q][s] = 0; }
        acc[q][0] = acc[q][1] = acc[q][2] = 0.f;
        alive[q] = inside[q] ? ((1u << S) - 1u) : 0u;
    }
    const float inv_s = 1.0f / (float)S;
    auto any_alive = [&]() {
        unsigned a = 0u;
#pragma unroll
        for (int q = 0; q < PPL; ++q) a |= alive[q];
        return a != 0u;
    };

    B200_COUNT_DECL;
    int b = 0;
    bool pending = false;
    if (nb > 0) { issue(0); pending = true; }
    for (; b < nb; ++b) {
        const int st = b & 1;
        if (b + 1 < nb) issue(b + 1);
        mbar_wait(&s_bar[st], (uint32_t)((b >> 1) & 1));
        pending = (b + 1 < nb);
        const int start = range.x + b * BLEND_BATCH;
        const int cnt = min(BLEND_BATCH, range.y - start);

        if (__any_sync(0xffffffffu, any_alive())) {
            for (int c0 = 0; c0 < cnt; c0 += 32) {
                const int e = c0 + lane;
                const unsigned my_mask = (e < cnt) ? sample_mask_exact<S>(s_rec[st][e], win, p.g.exposure) : 0u;
                unsigned m = __ballot_sync(0xffffffffu, my_mask != 0u);
                if (e < cnt) B200_COUNT(0, 1);
                while (m) {
                    const int src = __ffs(m) - 1;
                    const int k = c0 + src;
                    m &= m - 1;
                    const unsigned smask = __shfl_sync(0xffffffffu, my_mask, src);
                    if (lane == 0) { B200_COUNT(1, 1); B200_COUNT(2, __popc(smask)); }
#ifdef B200_BLEND_COUNTERS
                    bool blended_ = false;
#endif
                    const float4 A = *reinterpret_cast<const float4 *>(&s_rec[st][k].x);    // x y vx vy
                    const float4 Bq = *reinterpret_cast<const float4 *>(&s_rec[st][k].ca);  // a b c opac
                    const float4 C = *reinterpret_cast<const float4 *>(&s_rec[st][k].r);    // r g b thr
                    const float cut = C.w + 1e-4f;
                    const SigmaEntry se = sigma_entry(A, Bq);
                    float dx0[PPL], dy0[PPL];
#pragma unroll
                    for (int q = 0; q < PPL; ++q) { dx0[q] = A.x - px[q]; dy0[q] = A.y - py[q]; }
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        if (!(smask & (1u << s))) continue;  // warp-uniform
#pragma unroll
                        for (int q = 0; q < PPL; ++q) {
                            if (!(alive[q] & (1u << s))) continue;
                            const float tau = blur[s] + roll[q];
                            float dx, dy, sigma;
                            sigma_eval(se, px[q], py[q], dx0[q], dy0[q], tau, dx, dy, sigma);
                            B200_COUNT(3, 1);
                            if (sigma > cut || sigma < 0.f) continue;  // alpha < 1/255 guaranteed above thr
                            const float alpha = fminf(0.999f, Bq.w * exp_neg_approx(sigma));
                            if (alpha < 1.f / 255.f) continue;
                            B200_COUNT(4, 1);
#ifdef B200_BLEND_COUNTERS
                            blended_ = true;
#endif
                            const float next_T = T[q][s] * (1.f - alpha);
                            if (next_T <= 1e-4f) {  // forward.cu:421-427: this sample is done, entry not blended
                                alive[q] &= ~(1u << s);
                                continue;
                            }
                            const float vis = alpha * T[q][s] * inv_s;
                            acc[q][0] += C.x * vis; acc[q][1] += C.y * vis; acc[q][2] += C.z * vis;
                            T[q][s] = next_T;
                            last[q][s] = start + k;
                        }
                    }
#ifdef B200_BLEND_COUNTERS
                    if (__any_sync(0xffffffffu, blended_) && lane == 0) B200_COUNT(5, 1);
#endif
                    if (!__any_sync(0xffffffffu, any_alive())) { m = 0; c0 = cnt; }
                }
            }
        }
        // all warps are done with stage `st` (it is refilled two batches from now) + early exit vote
        if (!__syncthreads_or(any_alive())) { ++b; break; }
    }
    if (pending && b < nb) mbar_wait(&s_bar[b & 1], (uint32_t)((b >> 1) & 1));  // drain the prefetch before exit
    B200_COUNT_FLUSH(0);

    const float bg0 = __ldg(p.background), bg1 = __ldg(p.background + 1), bg2 = __ldg(p.background + 2);
#pragma unroll
    for (int q = 0; q < PPL; ++q) {
        if (!inside[q]) continue;
        const size_t pix = (size_t)pi[q] * p.g.W + pj[q];
        float meanT = 0.f, sumT = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            meanT += T[q][s] * inv_s;
            sumT += T[q][s];
            p.final_Ts[pix * S + s] = T[q][s];
            p.final_idx[pix * S + s] = last[q][s];
        }
        p.out_img[3 * pix] = acc[q][0] + meanT * bg0;
        p.out_img[3 * pix + 1] = acc[q][1] + meanT * bg1;
        p.out_img[3 * pix + 2] = acc[q][2] + meanT * bg2;
        if (p.out_alpha) p.out_alpha[pix] = 1.0f - sumT * inv_s;
    }
}


// ---- two pixels per lane, packed (PPL = 2, 16x16 tiles): the kernel the train step runs ------------------------------
// Same walk, staging and cull as blend_forward_kernel<S, 2>.  The lane's two pixels live in the halves of packed float
// pairs (f2, blend_common.cuh) and are evaluated side by side in straight-line code: the sigma quadratic, alpha * T and the
// three colour accumulations issue once for both pixels (FFMA2 / FMUL2 / FADD2), failed tests mask instead of branching
// (alpha = 0 leaves T, the colour sums and `last` untouched), and one warp vote per sample skips the blend arithmetic
// when no lane passes.  The branchy one-pixel-at-a-time form costs ~32 issue slots per pixel-sample, this ~23.
#ifndef B200_FWD_MIN_CTAS
#define B200_FWD_MIN_CTAS 6  // 80 registers
#endif
template <int S>
__global__ void __launch_bounds__(128, B200_FWD_MIN_CTAS) blend_forward_kernel2(const BlendFwdParams p) {
    constexpr int NT = 128;
    __shared__ __align__(128) PackedGaussian s_rec[BLEND_STAGES][BLEND_BATCH];
    __shared__ __align__(8) uint64_t s_bar[BLEND_STAGES];

    const int tid = threadIdx.x, lane = tid & 31;
    const int tile = blockIdx.x;
    const int tile_x = tile % p.g.tbx, tile_y = tile / p.g.tbx;
    bool inside[2];
    float px[2], py[2], roll[2];
    int pi[2], pj[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        int lx, ly;
        bool has_pixel;
        tile_pixel_ppl<2>(p.g.bw, tid, q, lx, ly, has_pixel);
        pj[q] = tile_x * p.g.bw + lx; pi[q] = tile_y * p.g.bw + ly;
        inside[q] = has_pixel && pi[q] < p.g.H && pj[q] < p.g.W;
        px[q] = (float)pj[q] + 0.5f; py[q] = (float)pi[q] + 0.5f;
        roll[q] = (float)((double)p.g.rs_time * ((double)(py[q] / (float)p.g.H) - 0.5));  // forward.cu:360
    }
    if (p.ref_isect && __ldg(p.ref_isect) < 1) {  // block-uniform
#pragma unroll
        for (int q = 0; q < 2; ++q) write_empty_render<S>(p, inside[q], pi[q], pj[q]);
        return;
    }
    float blur[S];
#pragma unroll
    for (int s = 0; s < S; ++s) blur[s] = blur_offset<S>(s, p.g.exposure);
    const WarpWindow win = warp_window<2>(inside, px, py, roll);
    const f2 PX = f2_make(px[0], px[1]), PY = f2_make(py[0], py[1]), ROLL = f2_make(roll[0], roll[1]);

    const int2 range = p.tile_bins[tile];
    const int total = range.y - range.x;
    const int nb = (total + BLEND_BATCH - 1) / BLEND_BATCH;

    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < BLEND_STAGES; ++st) mbar_init(&s_bar[st], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int b) {
        const int st = b & 1;
        const int start = range.x + b * BLEND_BATCH;
        const int cnt = min(BLEND_BATCH, range.y - start);
        if (tid == 0) mbar_arrive_expect_tx(&s_bar[st], (uint32_t)cnt * (uint32_t)sizeof(PackedGaussian));
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int e = tid + r * NT;
            if (e < cnt) {
                const int g = __ldg(p.ids_sorted + start + e);
                tma_bulk_g2s(&s_rec[st][e], p.packed + g, (uint32_t)sizeof(PackedGaussian), &s_bar[st]);
            }
        }
    };

    const float inv_s = 1.0f / (float)S;
    f2 T[S];
    int last[2][S];
    f2 acc0 = f2_splat(0.f), acc1 = f2_splat(0.f), acc2 = f2_splat(0.f);
    unsigned alive[2];
#pragma unroll
    for (int s = 0; s < S; ++s) { T[s] = f2_splat(1.f); last[0][s] = 0; last[1][s] = 0; }
#pragma unroll
    for (int q = 0; q < 2; ++q) alive[q] = inside[q] ? ((1u << S) - 1u) : 0u;
    unsigned walive = __reduce_or_sync(0xffffffffu, alive[0] | alive[1]);  // samples some lane of the warp still blends

    B200_COUNT_DECL;
    int b = 0;
    bool pending = false;
    if (nb > 0) { issue(0); pending = true; }
    for (; b < nb; ++b) {
        const int st = b & 1;
        if (b + 1 < nb) issue(b + 1);
        mbar_wait(&s_bar[st], (uint32_t)((b >> 1) & 1));
        pending = (b + 1 < nb);
        const int start = range.x + b * BLEND_BATCH;
        const int cnt = min(BLEND_BATCH, range.y - start);

        if (walive) {
            for (int c0 = 0; c0 < cnt; c0 += 32) {
                const int e = c0 + lane;
                const unsigned my_mask = (e < cnt) ? sample_mask_exact<S>(s_rec[st][e], win, p.g.exposure) : 0u;
                unsigned m = __ballot_sync(0xffffffffu, my_mask != 0u);
                if (e < cnt) B200_COUNT(0, 1);
#if B200_FWD_ALIVE_PER_CHUNK
                const unsigned alive_chunk = alive[0] + alive[1];  // (bits only ever clear: the sum drops iff one did)
#endif
                while (m) {
                    const int src = __ffs(m) - 1;
                    const int k = c0 + src;
                    m &= m - 1;
                    const unsigned smask = __shfl_sync(0xffffffffu, my_mask, src) & walive;
                    if (smask == 0u) continue;
                    if (lane == 0) { B200_COUNT(1, 1); B200_COUNT(2, __popc(smask)); }
                    const float4 A = *reinterpret_cast<const float4 *>(&s_rec[st][k].x);    // x y vx vy
                    const float4 Bq = *reinterpret_cast<const float4 *>(&s_rec[st][k].ca);  // a b c opac
                    const float4 C = *reinterpret_cast<const float4 *>(&s_rec[st][k].r);    // r g b thr
                    const float cut = C.w + 1e-4f;
                    const f2 VX = f2_splat(A.z), VY = f2_splat(A.w);
                    const f2 CBq = f2_splat(Bq.y), HA = f2_splat(0.5f * Bq.x), HC = f2_splat(0.5f * Bq.z), OP = f2_splat(Bq.w);
                    const f2 CR = f2_splat(C.x * inv_s), CG = f2_splat(C.y * inv_s), CBl = f2_splat(C.z * inv_s);
                    const f2 dx0 = f2_sub(f2_splat(A.x), PX), dy0 = f2_sub(f2_splat(A.y), PY);
                    const int idx = start + k;
#if !B200_FWD_ALIVE_PER_CHUNK
                    const unsigned alive_before = alive[0] + alive[1];
#endif
#ifdef B200_BLEND_COUNTERS
                    bool blended_ = false;
#endif
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        if (!(smask & (1u << s))) continue;  // warp-uniform
                        const f2 tau = f2_add(f2_splat(blur[s]), ROLL);
                        const f2 dx = f2_fma(tau, VX, dx0), dy = f2_fma(tau, VY, dy0);
                        const f2 u0 = f2_fma(HA, dx, f2_mul(CBq, dy));
                        const f2 sigma = f2_fma(dx, u0, f2_mul(f2_mul(HC, dy), dy));
                        const float sg0 = f2_lo(sigma), sg1 = f2_hi(sigma);
                        bool ok0 = (alive[0] & (1u << s)) != 0u, ok1 = (alive[1] & (1u << s)) != 0u;
                        if (ok0) B200_COUNT(3, 1);
                        if (ok1) B200_COUNT(3, 1);
                        ok0 = ok0 && !(sg0 > cut || sg0 < 0.f);  // alpha < 1/255 guaranteed above thr (NaN passes, as in the reference)
                        ok1 = ok1 && !(sg1 > cut || sg1 < 0.f);
                        const f2 ex = f2_mul(sigma, f2_splat(-1.4426950408889634f));
                        float v0, v1;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(v0) : "f"(f2_lo(ex)));
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(v1) : "f"(f2_hi(ex)));
                        const f2 ov = f2_mul(OP, f2_make(v0, v1));
                        const float a0 = fminf(0.999f, f2_lo(ov)), a1 = fminf(0.999f, f2_hi(ov));
                        ok0 = ok0 && !(a0 < 1.f / 255.f);
                        ok1 = ok1 && !(a1 < 1.f / 255.f);
#if B200_SAMPLE_VOTE
                        if (!__any_sync(0xffffffffu, ok0 || ok1)) continue;
#endif
                        if (ok0) B200_COUNT(4, 1);
                        if (ok1) B200_COUNT(4, 1);
#ifdef B200_BLEND_COUNTERS
                        blended_ = true;
#endif
                        const f2 nT = f2_mul(T[s], f2_sub(f2_splat(1.f), f2_make(a0, a1)));
                        // forward.cu:421-427: T would drop to <= 1e-4: this sample is done for the pixel, entry not blended
                        const bool stop0 = ok0 && (f2_lo(nT) <= 1e-4f), stop1 = ok1 && (f2_hi(nT) <= 1e-4f);
                        if (stop0) alive[0] &= ~(1u << s);
                        if (stop1) alive[1] &= ~(1u << s);
                        ok0 = ok0 && !stop0;
                        ok1 = ok1 && !stop1;
                        const f2 vis = f2_mul(f2_make(ok0 ? a0 : 0.f, ok1 ? a1 : 0.f), T[s]);  // alpha * T (1/S is in the colour)
                        acc0 = f2_fma(CR, vis, acc0); acc1 = f2_fma(CG, vis, acc1); acc2 = f2_fma(CBl, vis, acc2);
                        T[s] = f2_make(ok0 ? f2_lo(nT) : f2_lo(T[s]), ok1 ? f2_hi(nT) : f2_hi(T[s]));
                        last[0][s] = ok0 ? idx : last[0][s];
                        last[1][s] = ok1 ? idx : last[1][s];
                    }
#ifdef B200_BLEND_COUNTERS
                    if (__any_sync(0xffffffffu, blended_) && lane == 0) B200_COUNT(5, 1);
#endif
#if !B200_FWD_ALIVE_PER_CHUNK
                    // (bits only ever clear, so the sum of the two masks drops iff one did)
                    if (__any_sync(0xffffffffu, (alive[0] + alive[1]) != alive_before)) {  // rare: refresh the warp's live set
                        walive = __reduce_or_sync(0xffffffffu, alive[0] | alive[1]);
                        if (walive == 0u) { m = 0; c0 = cnt; }
                    }
#endif
                }
#if B200_FWD_ALIVE_PER_CHUNK
                // the warp's live-sample set is refreshed once per 32 entries instead of after every visit (a vote + compare
                // per visit for an event that happens S times per pixel): until then a sample that just finished for every
                // lane is still entered, fully masked -- no effect on any output
                if (__any_sync(0xffffffffu, (alive[0] + alive[1]) != alive_chunk)) {
                    walive = __reduce_or_sync(0xffffffffu, alive[0] | alive[1]);
                    if (walive == 0u) c0 = cnt;
                }
#endif
            }
        }
        // all warps are done with stage `st` (it is refilled two batches from now) + early exit vote
        if (!__syncthreads_or(walive != 0u)) { ++b; break; }
    }
    if (pending && b < nb) mbar_wait(&s_bar[b & 1], (uint32_t)((b >> 1) & 1));  // drain the prefetch before exit
    B200_COUNT_FLUSH(0);

    const float bg0 = __ldg(p.background), bg1 = __ldg(p.background + 1), bg2 = __ldg(p.background + 2);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (!inside[q]) continue;
        const size_t pix = (size_t)pi[q] * p.g.W + pj[q];
        float meanT = 0.f, sumT = 0.f;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const float Ts = q == 0 ? f2_lo(T[s]) : f2_hi(T[s]);
            meanT += Ts * inv_s;
            sumT += Ts;
            p.final_Ts[pix * S + s] = Ts;
            p.final_idx[pix * S + s] = last[q][s];
        }
        const float r = q == 0 ? f2_lo(acc0) : f2_hi(acc0), g = q == 0 ? f2_lo(acc1) : f2_hi(acc1), bl = q == 0 ? f2_lo(acc2) : f2_hi(acc2);
        p.out_img[3 * pix] = r + meanT * bg0;
        p.out_img[3 * pix + 1] = g + meanT * bg1;
        p.out_img[3 * pix + 2] = bl + meanT * bg2;
        if (p.out_alpha) p.out_alpha[pix] = 1.0f - sumT * inv_s;
    }
}

template <int S>
static int launch_fwd(const BlendFwdParams &p, cudaStream_t st) {
    // two pixels per lane when the tile is the full 16x16 (the only size Splatfacto uses, splatfacto.py:815)
    if (p.g.bw == 16 && blend_pixels_per_lane(false) == 4)  // experimental (B200_BLEND_PPL_FWD=4)
        blend_forward_kernel<S, 4><<<p.g.tbx * p.g.tby, BLEND_THREADS / 4, 0, st>>>(p);
    else if (p.g.bw == 16 && blend_pixels_per_lane(false) == 2 && blend_packed())
        blend_forward_kernel2<S><<<p.g.tbx * p.g.tby, BLEND_THREADS / 2, 0, st>>>(p);
    else if (p.g.bw == 16 && blend_pixels_per_lane(false) == 2)  // B200_BLEND_PACKED=0: the scalar two-pixel kernel (A/B)
        blend_forward_kernel<S, 2><<<p.g.tbx * p.g.tby, BLEND_THREADS / 2, 0, st>>>(p);
    else
        blend_forward_kernel<S, 1><<<p.g.tbx * p.g.tby, BLEND_THREADS, 0, st>>>(p);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200_packed_record_bytes(void) { return sizeof(PackedGaussian); }

static int run_blend_forward(unsigned img_height, unsigned img_width, unsigned block_width, unsigned n_blur_samples,
                             const int32_t *gaussian_ids_sorted, const int32_t *tile_bins, const void *packed,
                             float rolling_shutter_time, float exposure_time, const float *background, float *out_img,
                             float *final_Ts, int32_t *final_idx, float *out_alpha, cudaStream_t st,
                             const int32_t *ref_isect = nullptr) {
    B200_REQUIRE(n_blur_samples > 0 && n_blur_samples <= B200_MAX_BLUR_SAMPLES, "unsupported blur size");  // bindings.cu:450-452
    B200_REQUIRE(block_width > 1 && block_width <= 16, "block_width must be between 2 and 16");
    B200_REQUIRE(img_height > 0 && img_width > 0, "image size must be positive");
    B200_REQUIRE(tile_bins && background && packed, "null input pointer");
    B200_REQUIRE(aligned16(packed), "packed records must be 16-byte aligned");
    B200_REQUIRE(out_img && final_Ts && final_idx, "null output pointer");
    BlendFwdParams p;
    p.g = BlendGeom{(int)img_height, (int)img_width, (int)block_width,
                    (int)((img_width + block_width - 1) / block_width),
                    (int)((img_height + block_width - 1) / block_width), rolling_shutter_time, exposure_time};
    p.ids_sorted = gaussian_ids_sorted;
    p.tile_bins = reinterpret_cast<const int2 *>(tile_bins);
    p.packed = reinterpret_cast<const PackedGaussian *>(packed);
    p.background = background;
    p.out_img = out_img; p.final_Ts = final_Ts; p.final_idx = final_idx; p.out_alpha = out_alpha;
    p.ref_isect = ref_isect;
    switch (n_blur_samples) {
        case 1: return launch_fwd<1>(p, st);
        case 2: return launch_fwd<2>(p, st);
        case 3: return launch_fwd<3>(p, st);
        case 4: return launch_fwd<4>(p, st);
        case 5: return launch_fwd<5>(p, st);
        case 6: return launch_fwd<6>(p, st);
        case 7: return launch_fwd<7>(p, st);
        case 8: return launch_fwd<8>(p, st);
        case 9: return launch_fwd<9>(p, st);
        default: return launch_fwd<10>(p, st);
    }
}

extern "C" int b200_pack_records(int num_points, const float *xys, const float *pix_vels, const float *conics,
                                 const float *colors, const float *opacities, void *packed, void *stream) {
    B200_REQUIRE(num_points >= 1, "num_points must be >= 1");
    B200_REQUIRE(xys && pix_vels && conics && colors && opacities, "null input pointer");
    B200_REQUIRE(packed && aligned16(packed), "packed must be a 16-byte aligned buffer");
    return launch_pack(num_points, xys, pix_vels, conics, colors, opacities, packed, as_stream(stream));
}

extern "C" int b200_blend_forward_packed(unsigned img_height, unsigned img_width, unsigned block_width,
                                         unsigned n_blur_samples, const int32_t *gaussian_ids_sorted,
                                         const int32_t *tile_bins, const void *packed, float rolling_shutter_time,
                                         float exposure_time, const float *background, float *out_img, float *final_Ts,
                                         int32_t *final_idx, float *out_alpha, void *stream) {
    return run_blend_forward(img_height, img_width, block_width, n_blur_samples, gaussian_ids_sorted, tile_bins, packed,
                             rolling_shutter_time, exposure_time, background, out_img, final_Ts, final_idx, out_alpha,
                             as_stream(stream));
}

extern "C" int b200_rasterize_forward(int num_points, unsigned img_height, unsigned img_width, unsigned block_width,
                                      unsigned n_blur_samples, const int32_t *gaussian_ids_sorted,
                                      const int32_t *tile_bins, const float *xys, const float *pix_vels,
                                      float rolling_shutter_time, float exposure_time, const float *conics,
                                      const float *colors, const float *opacities, const float *background,
                                      void *packed_ws, float *out_img, float *final_Ts, int32_t *final_idx,
                                      void *stream) {
    B200_REQUIRE(n_blur_samples > 0 && n_blur_samples <= B200_MAX_BLUR_SAMPLES, "unsupported blur size");  // bindings.cu:450-452
    B200_REQUIRE(num_points >= 1, "num_points must be >= 1");
    B200_REQUIRE(gaussian_ids_sorted && tile_bins && xys && pix_vels && conics && colors && opacities && background,
                 "null input pointer");
    B200_REQUIRE(packed_ws && aligned16(packed_ws), "packed_ws must be a 16-byte aligned scratch buffer");
    cudaStream_t st = as_stream(stream);
    int rc = launch_pack(num_points, xys, pix_vels, conics, colors, opacities, packed_ws, st);
    if (rc) return rc;
    return run_blend_forward(img_height, img_width, block_width, n_blur_samples, gaussian_ids_sorted, tile_bins, packed_ws,
                             rolling_shutter_time, exposure_time, background, out_img, final_Ts, final_idx, nullptr, st);
}


// Capacity-mode companion of b200_blend_forward_packed: `status` is the DEVICE int32[4] block written by
// b200_bin_cull_emit_capacity; status[3] (the reference's num_intersects) < 1 reproduces the reference's empty-render
// branch (rasterize.py:136-144) without the host ever reading the count.
extern "C" int b200_blend_forward_packed_status(unsigned img_height, unsigned img_width, unsigned block_width,
                                                unsigned n_blur_samples, const int32_t *gaussian_ids_sorted,
                                                const int32_t *tile_bins, const void *packed, float rolling_shutter_time,
                                                float exposure_time, const float *background, const int32_t *status,
                                                float *out_img, float *final_Ts, int32_t *final_idx, float *out_alpha,
                                                void *stream) {
    B200_REQUIRE(status, "null status pointer");
    return run_blend_forward(img_height, img_width, block_width, n_blur_samples, gaussian_ids_sorted, tile_bins, packed,
                             rolling_shutter_time, exposure_time, background, out_img, final_Ts, final_idx, out_alpha,
                             as_stream(stream), status + 3);
}

namespace b200 {
// colours are the one part of the packed record the tile binning does not read: a caller that bins before it has shaded
// (the geometry of step k+1 is projected and binned while the SH exchange of step k is still in flight) packs with any
// colours and patches them in afterwards.  12 of every 64 bytes; the cull data (thr, hx, hy) stays.
static __global__ void __launch_bounds__(256) set_record_colors_kernel(int n, const float *__restrict__ colors,
                                                                       PackedGaussian *__restrict__ rec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    rec[i].r = colors[3 * (size_t)i]; rec[i].g = colors[3 * (size_t)i + 1]; rec[i].b = colors[3 * (size_t)i + 2];
}
}  // namespace b200

extern "C" int b200_set_record_colors(int num_points, const float *colors, void *packed, void *stream) {
    B200_REQUIRE(num_points >= 1, "num_points must be >= 1");
    B200_REQUIRE(colors && packed && aligned16(packed), "null / misaligned pointer");
    b200::set_record_colors_kernel<<<ceil_div(num_points, 256), 256, 0, as_stream(stream)>>>(
        num_points, colors, reinterpret_cast<PackedGaussian *>(packed));
    B200_LAUNCH_CHECK();
    return B200_OK;
}
