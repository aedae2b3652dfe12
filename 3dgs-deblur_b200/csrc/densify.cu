// Device-side densification for the flat Gaussian buffers ("next" row f-2).
//
// Semantics of the reference's adaptive density control, nerfstudio/models/splatfacto.py:
//   after_train            :408-434  running statistics (sum of |d loss / d xy| norms, visibility counts, max screen size)
//   refinement_after       :443-531  split / duplicate masks, concatenation order, culling, optimizer-state surgery
//   cull_gaussians         :533-566  low opacity / too big in world or on screen
//   split_gaussians        :568-611  children at mean + R(q) (exp(scale) * randn), scales / 1.6
//   dup_gaussians          :613-622
// The reference builds boolean masks, torch.cat's every parameter and every Adam moment, re-wraps the Parameters and
// empties the caching allocator (:376).  Here one kernel decides every flag, two prefix sums give every surviving row
// its final position (the reference's order: kept originals, then the children of sample 0, sample 1, ..., then the
// duplicates), and one gather per field writes the new flat parameter / moment buffers -- no masks, no cat, no host
// work besides reading the new count once to size the buffers.
#include <cub/cub.cuh>

#include "common.cuh"

namespace b200 {

struct DensifyCfg {
    float half_max_dim;        // 0.5 * max(H, W)                                   (:459)
    float grad_thresh;         // densify_grad_thresh                               (:460)
    float size_thresh;         // densify_size_thresh                               (:461, :469)
    float split_screen_size;   // < 0: screen-size splitting off (step >= stop_screen_size_at, :462)
    float cull_alpha_thresh;   // (:540)
    float cull_scale_thresh;   // < 0: no "too big" culling yet (step <= refine_every * reset_alpha_every, :545)
    float cull_screen_size;    // < 0: off (step >= stop_screen_size_at, :548)
    int samps;                 // n_split_samples
    int do_densify;            // 0: cull only (:503-504)
};

static size_t align256_(size_t x) { return (x + 255) & ~(size_t)255; }

struct DensifyWs {
    size_t flags, pos, split_mask, split_rank, cub, cub_bytes, total;
};
static DensifyWs densify_ws(int n, int samps) {
    DensifyWs L;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off += align256_(b); return o; };
    const size_t m = (size_t)(samps + 2) * (size_t)n;
    L.flags = take(4 * m); L.pos = take(4 * m); L.split_mask = take(4 * (size_t)n); L.split_rank = take(4 * (size_t)n);
    size_t b1 = 0, b2 = 0;
    cub::DeviceScan::ExclusiveSum((void *)nullptr, b1, (const int32_t *)nullptr, (int32_t *)nullptr, (int)std::min<size_t>(m, 0x7fffffff));
    cub::DeviceScan::ExclusiveSum((void *)nullptr, b2, (const int32_t *)nullptr, (int32_t *)nullptr, n);
    L.cub_bytes = (b1 > b2 ? b1 : b2) + 256;
    L.cub = take(L.cub_bytes);
    L.total = off;
    return L;
}

// running statistics of one training image (splatfacto.py:408-434)
__global__ void __launch_bounds__(256) densify_accumulate_kernel(int n, const float2 *__restrict__ absgrad,
                                                                 const int32_t *__restrict__ radii, float max_dim,
                                                                 int first, float *__restrict__ grad_norm,
                                                                 float *__restrict__ vis, float *__restrict__ max2d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 g = absgrad[i];
    const float gn = sqrtf(g.x * g.x + g.y * g.y);
    const int r = radii[i];
    const bool visible = r > 0;
    if (first) {  // :417-419 the first image initialises every Gaussian (count 1, its norm -- zero if it was not seen)
        grad_norm[i] = gn;
        vis[i] = 1.f;
        max2d[i] = 0.f;
    } else if (visible) {
        vis[i] += 1.f;
        grad_norm[i] += gn;
    }
    if (visible) max2d[i] = fmaxf(max2d[i], (float)r / max_dim);  // :429-433 (a true division: r = 40 of 800 must compare equal to 0.05)
}

// every decision of refinement_after / cull_gaussians for Gaussian i and for its potential children
__global__ void __launch_bounds__(256) densify_flags_kernel(int n, const float *__restrict__ log_scales,
                                                            const float *__restrict__ opacity_logit,
                                                            const float *__restrict__ grad_norm,
                                                            const float *__restrict__ vis,
                                                            const float *__restrict__ max2d, DensifyCfg c,
                                                            int32_t *__restrict__ flags, int32_t *__restrict__ split_mask,
                                                            int32_t *__restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int split = 0, dup = 0, cull = 0;
    if (i < n) {
        const float s0 = expf(log_scales[3 * (size_t)i]), s1 = expf(log_scales[3 * (size_t)i + 1]), s2 = expf(log_scales[3 * (size_t)i + 2]);
        const float smax = fmaxf(s0, fmaxf(s1, s2));
        const float m2 = max2d ? max2d[i] : 0.f;
        const float alpha = 1.f / (1.f + expf(-opacity_logit[i]));
        const bool alpha_low = alpha < c.cull_alpha_thresh;                                        // :540
        const bool too_big_world = c.cull_scale_thresh >= 0.f && smax > c.cull_scale_thresh;       // :547
        const bool too_big_screen = c.cull_scale_thresh >= 0.f && c.cull_screen_size >= 0.f && m2 > c.cull_screen_size;  // :548-551
        bool keep_split = false, keep_dup = false;
        if (c.do_densify) {
            const float avg = (grad_norm[i] / vis[i]) * c.half_max_dim;                             // :459
            const bool high = avg > c.grad_thresh;                                                  // :460
            bool sp = smax > c.size_thresh;                                                         // :461
            if (c.split_screen_size >= 0.f) sp = sp || (m2 > c.split_screen_size);                  // :462-463
            split = (sp && high) ? 1 : 0;                                                           // :464
            // split_gaussians shrinks the PARENT in place too (:597) BEFORE the duplicate mask is formed (:469): a split
            // parent whose shrunken size falls under the threshold is duplicated as well, with the shrunken scales
            const float smax_now = split ? expf(logf(smax / 1.6f)) : smax;
            dup = (smax_now <= c.size_thresh && high) ? 1 : 0;                                      // :469-470
            // children carry the parent's opacity; a split child has scales / 1.6 (:596), a duplicate the parent's
            // current ones; both start with max_2Dsize = 0 (:478-485)
            const bool shrunk_too_big = c.cull_scale_thresh >= 0.f && expf(logf(smax / 1.6f)) > c.cull_scale_thresh;
            keep_split = split && !alpha_low && !shrunk_too_big;
            keep_dup = dup && !alpha_low && !(split ? shrunk_too_big : too_big_world);
        }
        // the original: culled if transparent, too big, or split (its children replace it, :493-501)
        const bool keep_orig = !(alpha_low || too_big_world || too_big_screen || split);
        flags[i] = keep_orig ? 1 : 0;
        for (int j = 0; j < c.samps; ++j) flags[(size_t)(1 + j) * n + i] = keep_split ? 1 : 0;
        flags[(size_t)(1 + c.samps) * n + i] = keep_dup ? 1 : 0;
        split_mask[i] = split;
        cull = keep_orig ? 0 : 1;
    }
    split = __reduce_add_sync(0xffffffffu, split);
    dup = __reduce_add_sync(0xffffffffu, dup);
    cull = __reduce_add_sync(0xffffffffu, cull);
    if ((threadIdx.x & 31) == 0) {
        if (split) atomicAdd(counts + 1, split);
        if (dup) atomicAdd(counts + 2, dup);
        if (cull) atomicAdd(counts + 3, cull);
    }
}

__global__ void densify_total_kernel(size_t m, const int32_t *__restrict__ flags, const int32_t *__restrict__ pos,
                                     int32_t *__restrict__ counts) {
    if (threadIdx.x == 0 && blockIdx.x == 0) counts[0] = pos[m - 1] + flags[m - 1];
}

// FIELD: 0 copy; 1 means (children of a split are re-sampled); 2 log-scales (children of a split shrink); 3 optimizer
// moment (children start at zero, :384-399)
template <int FIELD>
__global__ void __launch_bounds__(256) densify_gather_kernel(int n, int samps, int width, const float *__restrict__ src,
                                                             float *__restrict__ dst, const int32_t *__restrict__ flags,
                                                             const int32_t *__restrict__ pos,
                                                             const int32_t *__restrict__ split_rank,
                                                             const int32_t *__restrict__ split_mask,
                                                             const int32_t *__restrict__ counts,
                                                             const float *__restrict__ log_scales,
                                                             const float *__restrict__ quats,
                                                             const float *__restrict__ randn) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t m = (size_t)(samps + 2) * n;
    if (t >= m) return;
    if (!flags[t]) return;
    const int k = (int)(t / n), i = (int)(t - (size_t)k * n);
    const size_t r = (size_t)pos[t];
    const bool split_child = k >= 1 && k <= samps;
    if (FIELD == 3 && k != 0) {
        for (int w = 0; w < width; ++w) dst[r * width + w] = 0.f;
        return;
    }
    if (FIELD == 1 && split_child) {
        // new mean = mean + R(q / |q|) (exp(scale) * z), z = row (sample * n_splits + rank among the splits) of the
        // caller's torch.randn((samps * n_splits, 3)) -- the reference's draw (:574-582), so equal seeds give equal children
        const int n_splits = counts[1];
        const float *z = randn + 3 * ((size_t)(k - 1) * n_splits + split_rank[i]);
        float qw = quats[4 * (size_t)i], qx = quats[4 * (size_t)i + 1], qy = quats[4 * (size_t)i + 2], qz = quats[4 * (size_t)i + 3];
        const float inv = 1.f / sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
        qw *= inv; qx *= inv; qy *= inv; qz *= inv;
        float R[9];
        quat_to_rotmat(qw, qx, qy, qz, R);
        const float v0 = expf(log_scales[3 * (size_t)i]) * z[0], v1 = expf(log_scales[3 * (size_t)i + 1]) * z[1],
                    v2 = expf(log_scales[3 * (size_t)i + 2]) * z[2];
        dst[r * 3 + 0] = (R[0] * v0 + R[1] * v1 + R[2] * v2) + src[3 * (size_t)i];
        dst[r * 3 + 1] = (R[3] * v0 + R[4] * v1 + R[5] * v2) + src[3 * (size_t)i + 1];
        dst[r * 3 + 2] = (R[6] * v0 + R[7] * v1 + R[8] * v2) + src[3 * (size_t)i + 2];
        return;
    }
    if (FIELD == 2 && (split_child || (k == samps + 1 && split_mask[i]))) {  // (a duplicate of a split parent: shrunken too)
        for (int w = 0; w < 3; ++w) dst[r * 3 + w] = logf(expf(src[3 * (size_t)i + w]) / 1.6f);  // :596-597
        return;
    }
    for (int w = 0; w < width; ++w) dst[r * width + w] = src[(size_t)i * width + w];
}

}  // namespace b200

using namespace b200;

extern "C" int b200_densify_accumulate(int num_points, const float *absgrad, const int32_t *radii, float max_dim, int first,
                                       float *grad_norm, float *vis_counts, float *max_2d, void *stream) {
    B200_REQUIRE(num_points >= 1, "num_points must be >= 1");
    B200_REQUIRE(absgrad && radii && grad_norm && vis_counts && max_2d, "null pointer");
    densify_accumulate_kernel<<<ceil_div(num_points, 256), 256, 0, as_stream(stream)>>>(
        num_points, reinterpret_cast<const float2 *>(absgrad), radii, max_dim, first, grad_norm, vis_counts, max_2d);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" size_t b200_densify_ws_bytes(int num_points, int n_split_samples) {
    return densify_ws(num_points > 0 ? num_points : 1, n_split_samples > 0 ? n_split_samples : 1).total;
}

extern "C" int b200_densify_plan(int num_points, const float *log_scales, const float *opacity_logit, const float *grad_norm,
                                 const float *vis_counts, const float *max_2d, float half_max_dim, float densify_grad_thresh,
                                 float densify_size_thresh, float split_screen_size, int n_split_samples,
                                 float cull_alpha_thresh, float cull_scale_thresh, float cull_screen_size, int do_densify,
                                 void *ws, size_t ws_bytes, int32_t *counts, void *stream) {
    B200_REQUIRE(num_points >= 1 && n_split_samples >= 1, "bad sizes");
    B200_REQUIRE(log_scales && opacity_logit && ws && counts, "null pointer");
    B200_REQUIRE(!do_densify || (grad_norm && vis_counts), "densification needs the running statistics");
    const int n = num_points;
    const DensifyWs L = densify_ws(n, n_split_samples);
    B200_REQUIRE(ws_bytes >= L.total, "workspace too small: %zu < %zu", ws_bytes, L.total);
    B200_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255u) == 0, "workspace must be 256-byte aligned");
    const size_t m = (size_t)(n_split_samples + 2) * n;
    B200_REQUIRE(m < 0x7fffffffull, "too many rows");
    char *base = static_cast<char *>(ws);
    int32_t *flags = (int32_t *)(base + L.flags), *pos = (int32_t *)(base + L.pos);
    int32_t *split_mask = (int32_t *)(base + L.split_mask), *split_rank = (int32_t *)(base + L.split_rank);
    void *cub_ws = base + L.cub;
    size_t cub_bytes = L.cub_bytes;
    cudaStream_t st = as_stream(stream);
    DensifyCfg c{half_max_dim, densify_grad_thresh, densify_size_thresh, split_screen_size, cull_alpha_thresh, cull_scale_thresh,
                 cull_screen_size, n_split_samples, do_densify};
    B200_CUDA(cudaMemsetAsync(counts, 0, 4 * sizeof(int32_t), st));
    densify_flags_kernel<<<ceil_div(n, 256), 256, 0, st>>>(n, log_scales, opacity_logit, grad_norm, vis_counts, max_2d, c, flags,
                                                           split_mask, counts);
    B200_LAUNCH_CHECK();
    B200_CUDA(cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, flags, pos, (int)m, st));
    B200_CUDA(cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, split_mask, split_rank, n, st));
    count_launch(4);
    densify_total_kernel<<<1, 32, 0, st>>>(m, flags, pos, counts);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_densify_gather(int num_points, int n_split_samples, int field, int width, const float *src, float *dst,
                                   const void *ws, const int32_t *counts, const float *log_scales, const float *quats,
                                   const float *randn, void *stream) {
    B200_REQUIRE(num_points >= 1 && n_split_samples >= 1 && width >= 1, "bad sizes");
    B200_REQUIRE(field >= 0 && field <= 3, "field must be 0 (copy), 1 (means), 2 (log-scales) or 3 (optimizer moment)");
    B200_REQUIRE(src && dst && ws && counts, "null pointer");
    B200_REQUIRE(field != 1 || (log_scales && quats && randn && width == 3), "means need log_scales, quats and the normal samples");
    B200_REQUIRE(field != 2 || width == 3, "log-scales are (N, 3)");
    const int n = num_points;
    const DensifyWs L = densify_ws(n, n_split_samples);
    const char *base = static_cast<const char *>(ws);
    const int32_t *flags = (const int32_t *)(base + L.flags), *pos = (const int32_t *)(base + L.pos);
    const int32_t *split_rank = (const int32_t *)(base + L.split_rank), *split_mask = (const int32_t *)(base + L.split_mask);
    const size_t m = (size_t)(n_split_samples + 2) * n;
    const int blocks = (int)((m + 255) / 256);
    cudaStream_t st = as_stream(stream);
    switch (field) {
        case 0: densify_gather_kernel<0><<<blocks, 256, 0, st>>>(n, n_split_samples, width, src, dst, flags, pos, split_rank, split_mask, counts, log_scales, quats, randn); break;
        case 1: densify_gather_kernel<1><<<blocks, 256, 0, st>>>(n, n_split_samples, width, src, dst, flags, pos, split_rank, split_mask, counts, log_scales, quats, randn); break;
        case 2: densify_gather_kernel<2><<<blocks, 256, 0, st>>>(n, n_split_samples, width, src, dst, flags, pos, split_rank, split_mask, counts, log_scales, quats, randn); break;
        default: densify_gather_kernel<3><<<blocks, 256, 0, st>>>(n, n_split_samples, width, src, dst, flags, pos, split_rank, split_mask, counts, log_scales, quats, randn); break;
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}
