// Fused per-Gaussian pre/post-processing for callers that hand over RAW Splatfacto parameters ("next" row f-1 of
// the caller-side glue of nerfstudio/models/splatfacto.py:816-856 folded into the kernels).
//
// forward  (one pass over the Gaussians): exp(log_scales), quats / |quats|, projection (forward.cu:13-112 semantics),
//           SH colour + clamp(rgb + 0.5, 0) for the Gaussians that survive the bbox test only, sigmoid(opacity) *
//           compensation, and the 64-byte blend record -- instead of ~12 PyTorch kernels, the `torch.cat` of the SH
//           tensors (384 B per Gaussian of copy traffic) and three library kernels.  Culled Gaussians (79 % in BASELINE
//           config 2) never touch their 192 bytes of SH coefficients.
// backward (one pass): recomputes the cheap forward quantities, chains blend gradients through the SH basis, the clamp,
//           sigmoid * compensation, the projection VJP (clamp-aware, with camera-velocity and exact view-matrix
//           gradients), exp and the quaternion normalisation, and writes every gradient row (zeros for culled
//           Gaussians, so the caller needs no memset) with coalesced stores.
#include "blend_common.cuh"
#include "projection_math.cuh"
#include "sh_math.cuh"

namespace b200 {

constexpr int FUSED_THREADS = 128;

struct FusedParams {
    ProjCommon proj;        // means = raw means; scales/quats unused (raw pointers below)
    const float *log_scales, *quats_raw, *opacity_logit, *sh_dc, *sh_rest, *cam_pos;
    int K, deg_use;
    bool quats_al;          // quats_raw is 16-byte aligned (vector loads); otherwise scalar loads
};

__device__ __forceinline__ float4 load_quat(const FusedParams &f, int i) {
    if (f.quats_al) return reinterpret_cast<const float4 *>(f.quats_raw)[i];
    const float *q = f.quats_raw + 4 * (size_t)i;
    return make_float4(q[0], q[1], q[2], q[3]);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// SH colour of Gaussian i before the + 0.5 / clamp (splatfacto.py:840-846): "fast" basis, dc and rest tensors read apart.
template <int DEG_USE>
__device__ __forceinline__ void sh_colour(const FusedParams &f, int i, const float *m, float &c0, float &c1, float &c2) {
    constexpr int KU = (DEG_USE + 1) * (DEG_USE + 1);
    float B[KU];
    sh_basis<B200_SH_FAST>(DEG_USE, m[0] - __ldg(f.cam_pos), m[1] - __ldg(f.cam_pos + 1), m[2] - __ldg(f.cam_pos + 2), B);
    const float *dc = f.sh_dc + 3 * (size_t)i;
    c0 = B[0] * dc[0]; c1 = B[0] * dc[1]; c2 = B[0] * dc[2];
    const float *rest = f.sh_rest + (size_t)i * (f.K - 1) * 3;
#pragma unroll
    for (int k = 1; k < KU; ++k) {
        c0 = fmaf(B[k], rest[3 * (k - 1)], c0);
        c1 = fmaf(B[k], rest[3 * (k - 1) + 1], c1);
        c2 = fmaf(B[k], rest[3 * (k - 1) + 2], c2);
    }
}

// DEG_USE < 0: geometry only -- the record's colour is left 0 for fused_colors_kernel to fill in later (a trainer that
// projects and bins the next image while the SH block of the previous step is still being exchanged / updated).
template <int DEG_USE>
__global__ void __launch_bounds__(FUSED_THREADS) fused_forward_kernel(FusedParams f, PackedGaussian *__restrict__ rec,
                                                                      float *__restrict__ depths,
                                                                      int32_t *__restrict__ radii,
                                                                      int32_t *__restrict__ tiles_hit) {
    constexpr int KU = DEG_USE < 0 ? 1 : (DEG_USE + 1) * (DEG_USE + 1);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f.proj.n) return;
    const float *m = f.proj.means + 3 * (size_t)i, *ls = f.log_scales + 3 * (size_t)i;
    const float4 q = load_quat(f, i);
    const float qn = 1.f / sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    ProjGaussIn in{m[0], m[1], m[2], expf(ls[0]), expf(ls[1]), expf(ls[2]), q.x * qn, q.y * qn, q.z * qn, q.w * qn};
    ProjGaussOut r;
    project_forward_one(f.proj, in, r);
    depths[i] = r.depth; radii[i] = r.radius_i; tiles_hit[i] = r.tiles;
    PackedGaussian g;
    if (r.tiles > 0) {
        float c0 = -0.5f, c1 = -0.5f, c2 = -0.5f;
        if constexpr (DEG_USE >= 0) sh_colour<DEG_USE>(f, i, m, c0, c1, c2);
        const float opac = sigmoidf_(f.opacity_logit[i]) * r.comp;  // splatfacto.py:853-854 ("antialiased")
        g = make_record(i, r.xy[0], r.xy[1], r.vel[0], r.vel[1], r.conic[0], r.conic[1], r.conic[2], opac,
                        fmaxf(c0 + 0.5f, 0.f), fmaxf(c1 + 0.5f, 0.f), fmaxf(c2 + 0.5f, 0.f));  // :846
    } else {
        g = make_record(i, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f);  // hx = -1: never blended
    }
    store_record(rec + i, g);
}

// The colour half of fused_forward_kernel for records built geometry-only: clamp(SH colour + 0.5, 0) of every Gaussian
// that survived the projection (radius > 0, the set the backward kernel differentiates) into its record.
template <int DEG_USE>
__global__ void __launch_bounds__(FUSED_THREADS) fused_colors_kernel(FusedParams f, const int32_t *__restrict__ radii,
                                                                     PackedGaussian *__restrict__ rec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f.proj.n || radii[i] <= 0) return;
    float c0, c1, c2;
    sh_colour<DEG_USE>(f, i, f.proj.means + 3 * (size_t)i, c0, c1, c2);
    rec[i].r = fmaxf(c0 + 0.5f, 0.f); rec[i].g = fmaxf(c1 + 0.5f, 0.f); rec[i].b = fmaxf(c2 + 0.5f, 0.f);
}

struct FusedBwdIO {
    const PackedGaussian *rec;
    const int32_t *radii;
    const float *v_xy, *v_pix_vel, *v_conic, *v_colors, *v_opacity;
    float *g_means, *g_log_scales, *g_quats, *g_opacity, *g_sh_dc, *g_sh_rest;
    float *g_lin, *g_ang, *g_viewmat;  // optional, pre-zeroed accumulators
    bool gq_al;                        // g_quats is 16-byte aligned
};

// Coalesced store of a block's contiguous chunk of `row`-float rows staged in shared memory with pitch `pitch`.
__device__ __forceinline__ void store_rows(float *dst, const float *s, int row, int pitch, int count) {
    const int nfloat = row * count;
    for (int fidx = threadIdx.x; fidx < nfloat; fidx += FUSED_THREADS) {
        const int r = fidx / row;
        dst[fidx] = s[r * pitch + (fidx - r * row)];
    }
}

template <int DEG_USE>
__global__ void __launch_bounds__(FUSED_THREADS) fused_backward_kernel(FusedParams f, FusedBwdIO io) {
    constexpr int KU = (DEG_USE + 1) * (DEG_USE + 1);
    extern __shared__ __align__(16) float s_rows[];  // [FUSED_THREADS][pitch] gradient rows of sh_rest
    __shared__ float s_red[FUSED_THREADS / 32][18];
    const int rest_row = 3 * (f.K - 1);
    const int pitch = (rest_row & 1) ? rest_row : rest_row + 1;
    const int t = threadIdx.x;
    const int base = blockIdx.x * FUSED_THREADS;
    const int count = min(FUSED_THREADS, f.proj.n - base);
    const int i = base + t;
    const bool want_cam = io.g_lin || io.g_ang || io.g_viewmat;

    ProjGaussGrad gr;
#pragma unroll
    for (int k = 0; k < 18; ++k) gr.red[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) gr.v_c3[k] = 0.f;
    gr.v_mean[0] = gr.v_mean[1] = gr.v_mean[2] = 0.f; gr.v_scale[0] = gr.v_scale[1] = gr.v_scale[2] = 0.f;
    gr.v_quat[0] = gr.v_quat[1] = gr.v_quat[2] = gr.v_quat[3] = 0.f; gr.vc2[0] = gr.vc2[1] = gr.vc2[2] = 0.f;
    float g_ls[3] = {0.f, 0.f, 0.f}, g_q[4] = {0.f, 0.f, 0.f, 0.f}, g_op = 0.f, g_dc[3] = {0.f, 0.f, 0.f};
    float *my_rows = s_rows + t * pitch;
    bool active = false;
    if (t < count) {
        active = io.radii[i] > 0;
        if (!active)
            for (int k = 0; k < rest_row; ++k) my_rows[k] = 0.f;
    }
    if (active) {
        const float *m = f.proj.means + 3 * (size_t)i, *ls = f.log_scales + 3 * (size_t)i;
        const float4 q = load_quat(f, i);
        const float qn = 1.f / sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
        ProjGaussIn in{m[0], m[1], m[2], expf(ls[0]), expf(ls[1]), expf(ls[2]), q.x * qn, q.y * qn, q.z * qn, q.w * qn};
        ProjGaussOut r;
        project_forward_one(f.proj, in, r);  // recompute instead of saving 40 B per Gaussian
        const float4 C = *reinterpret_cast<const float4 *>(&io.rec[i].r);  // clamped colour of the forward
        // colour: clamp gate, then basis outer product (sh.cuh:467-498)
        const float v0 = C.x > 0.f ? io.v_colors[3 * (size_t)i] : 0.f, v1 = C.y > 0.f ? io.v_colors[3 * (size_t)i + 1] : 0.f,
                    v2 = C.z > 0.f ? io.v_colors[3 * (size_t)i + 2] : 0.f;
        float B[KU];
        sh_basis<B200_SH_FAST>(DEG_USE, m[0] - __ldg(f.cam_pos), m[1] - __ldg(f.cam_pos + 1), m[2] - __ldg(f.cam_pos + 2), B);
        g_dc[0] = B[0] * v0; g_dc[1] = B[0] * v1; g_dc[2] = B[0] * v2;
#pragma unroll
        for (int k = 1; k < KU; ++k) {
            my_rows[3 * (k - 1)] = B[k] * v0; my_rows[3 * (k - 1) + 1] = B[k] * v1; my_rows[3 * (k - 1) + 2] = B[k] * v2;
        }
        for (int k = 3 * (KU - 1); k < rest_row; ++k) my_rows[k] = 0.f;
        // opacity = sigmoid(logit) * comp
        const float sg = sigmoidf_(f.opacity_logit[i]);
        const float vo = io.v_opacity[i];
        g_op = vo * r.comp * sg * (1.f - sg);
        // projection VJP
        ProjGaussSaved sv;
#pragma unroll
        for (int k = 0; k < 6; ++k) sv.cov3d[k] = r.cov3d[k];
        sv.conic[0] = r.conic[0]; sv.conic[1] = r.conic[1]; sv.conic[2] = r.conic[2];
        sv.comp = r.comp;
        ProjGaussCot ct;
        const float2 gxy = reinterpret_cast<const float2 *>(io.v_xy)[i], gpv = reinterpret_cast<const float2 *>(io.v_pix_vel)[i];
        ct.v_xy[0] = gxy.x; ct.v_xy[1] = gxy.y; ct.v_pix_vel[0] = gpv.x; ct.v_pix_vel[1] = gpv.y;
        ct.v_depth = 0.f;
        ct.v_conic[0] = io.v_conic[3 * (size_t)i]; ct.v_conic[1] = io.v_conic[3 * (size_t)i + 1]; ct.v_conic[2] = io.v_conic[3 * (size_t)i + 2];
        ct.v_comp = vo * sg;
        project_backward_one<true>(f.proj, want_cam, io.g_viewmat != nullptr, in, sv, ct, gr);
        // exp and normalisation chains
        g_ls[0] = gr.v_scale[0] * in.s0; g_ls[1] = gr.v_scale[1] * in.s1; g_ls[2] = gr.v_scale[2] * in.s2;
        const float dotq = in.qw * gr.v_quat[0] + in.qx * gr.v_quat[1] + in.qy * gr.v_quat[2] + in.qz * gr.v_quat[3];
        g_q[0] = (gr.v_quat[0] - in.qw * dotq) * qn; g_q[1] = (gr.v_quat[1] - in.qx * dotq) * qn;
        g_q[2] = (gr.v_quat[2] - in.qy * dotq) * qn; g_q[3] = (gr.v_quat[3] - in.qz * dotq) * qn;
    }
    if (t < count) {
        io.g_means[3 * (size_t)i] = gr.v_mean[0]; io.g_means[3 * (size_t)i + 1] = gr.v_mean[1]; io.g_means[3 * (size_t)i + 2] = gr.v_mean[2];
        io.g_log_scales[3 * (size_t)i] = g_ls[0]; io.g_log_scales[3 * (size_t)i + 1] = g_ls[1]; io.g_log_scales[3 * (size_t)i + 2] = g_ls[2];
        if (io.gq_al) {
            reinterpret_cast<float4 *>(io.g_quats)[i] = make_float4(g_q[0], g_q[1], g_q[2], g_q[3]);
        } else {
            float *gq = io.g_quats + 4 * (size_t)i;
            gq[0] = g_q[0]; gq[1] = g_q[1]; gq[2] = g_q[2]; gq[3] = g_q[3];
        }
        io.g_opacity[i] = g_op;
        io.g_sh_dc[3 * (size_t)i] = g_dc[0]; io.g_sh_dc[3 * (size_t)i + 1] = g_dc[1]; io.g_sh_dc[3 * (size_t)i + 2] = g_dc[2];
    }
    __syncthreads();
    store_rows(io.g_sh_rest + (size_t)base * rest_row, s_rows, rest_row, pitch, count);

    if (want_cam) {
        const int lane = t & 31, warp = t >> 5;
#pragma unroll
        for (int k = 0; k < 18; ++k) {
            const float s = warp_sum(gr.red[k]);
            if (lane == 0) s_red[warp][k] = s;
        }
        __syncthreads();
        if (t < 18) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < FUSED_THREADS / 32; ++w) s += s_red[w][t];
            float *dst = t < 3 ? (io.g_lin ? io.g_lin + t : nullptr)
                       : t < 6 ? (io.g_ang ? io.g_ang + (t - 3) : nullptr)
                               : (io.g_viewmat ? io.g_viewmat + (t - 6) : nullptr);
            if (dst && s != 0.f) atomicAdd(dst, s);
        }
    }
}

static int fill_fused(FusedParams &f, int n, const float *means, const float *log_scales, const float *quats,
                      const float *opacity_logit, const float *sh_dc, const float *sh_rest, int sh_bases,
                      int degrees_to_use, const float *viewmat, const float *cam_pos, const float *lin_vel,
                      const float *ang_vel, float rs, float exposure, float fx, float fy, float cx, float cy, unsigned H,
                      unsigned W, unsigned bw, float clip, bool with_colors = true) {
    B200_REQUIRE(n >= 1, "num_points must be >= 1");
    B200_REQUIRE(means && log_scales && quats && opacity_logit && viewmat, "null input pointer");
    if (with_colors) {
        B200_REQUIRE(sh_dc && cam_pos, "null input pointer");
        B200_REQUIRE(sh_bases == 1 || sh_bases == 4 || sh_bases == 9 || sh_bases == 16 || sh_bases == 25, "bad SH basis count %d", sh_bases);
        B200_REQUIRE(degrees_to_use >= 0 && (degrees_to_use + 1) * (degrees_to_use + 1) <= sh_bases, "degrees_to_use too large");
        B200_REQUIRE(sh_bases == 1 || sh_rest, "null sh_rest");
    }
    B200_REQUIRE(bw > 1 && bw <= 16, "block_width must be between 2 and 16");
    B200_REQUIRE(H > 0 && W > 0, "image size must be positive");
    f.proj = ProjCommon{n, means, nullptr, nullptr, lin_vel, ang_vel, viewmat, 1.0f, rs, exposure, fx, fy, cx, cy, (int)H, (int)W, (int)bw, clip};
    f.log_scales = log_scales; f.quats_raw = quats; f.opacity_logit = opacity_logit; f.sh_dc = sh_dc; f.sh_rest = sh_rest;
    f.cam_pos = cam_pos; f.K = sh_bases; f.deg_use = degrees_to_use; f.quats_al = aligned16(quats);
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_fused_preprocess_forward(int num_points, const float *means, const float *log_scales,
                                             const float *quats, const float *opacity_logit, const float *sh_dc,
                                             const float *sh_rest, int sh_bases, int degrees_to_use,
                                             const float *viewmat, const float *cam_pos, const float *lin_vel,
                                             const float *ang_vel, float rolling_shutter_time, float exposure_time,
                                             float fx, float fy, float cx, float cy, unsigned img_height,
                                             unsigned img_width, unsigned block_width, float clip_thresh, void *packed,
                                             float *depths, int32_t *radii, int32_t *num_tiles_hit, void *stream) {
    FusedParams f;
    int rc = fill_fused(f, num_points, means, log_scales, quats, opacity_logit, sh_dc, sh_rest, sh_bases, degrees_to_use,
                        viewmat, cam_pos, lin_vel, ang_vel, rolling_shutter_time, exposure_time, fx, fy, cx, cy, img_height,
                        img_width, block_width, clip_thresh);
    if (rc) return rc;
    B200_REQUIRE(packed && aligned16(packed) && depths && radii && num_tiles_hit, "null / misaligned output pointer");
    const int blocks = ceil_div(num_points, FUSED_THREADS);
    cudaStream_t st = as_stream(stream);
    PackedGaussian *rec = reinterpret_cast<PackedGaussian *>(packed);
    switch (degrees_to_use) {
        case 0: fused_forward_kernel<0><<<blocks, FUSED_THREADS, 0, st>>>(f, rec, depths, radii, num_tiles_hit); break;
        case 1: fused_forward_kernel<1><<<blocks, FUSED_THREADS, 0, st>>>(f, rec, depths, radii, num_tiles_hit); break;
        case 2: fused_forward_kernel<2><<<blocks, FUSED_THREADS, 0, st>>>(f, rec, depths, radii, num_tiles_hit); break;
        case 3: fused_forward_kernel<3><<<blocks, FUSED_THREADS, 0, st>>>(f, rec, depths, radii, num_tiles_hit); break;
        default: fused_forward_kernel<4><<<blocks, FUSED_THREADS, 0, st>>>(f, rec, depths, radii, num_tiles_hit); break;
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_fused_geometry_forward(int num_points, const float *means, const float *log_scales, const float *quats,
                                           const float *opacity_logit, const float *viewmat, const float *lin_vel,
                                           const float *ang_vel, float rolling_shutter_time, float exposure_time, float fx,
                                           float fy, float cx, float cy, unsigned img_height, unsigned img_width,
                                           unsigned block_width, float clip_thresh, void *packed, float *depths,
                                           int32_t *radii, int32_t *num_tiles_hit, void *stream) {
    FusedParams f;
    int rc = fill_fused(f, num_points, means, log_scales, quats, opacity_logit, nullptr, nullptr, 1, 0, viewmat, nullptr,
                        lin_vel, ang_vel, rolling_shutter_time, exposure_time, fx, fy, cx, cy, img_height, img_width,
                        block_width, clip_thresh, false);
    if (rc) return rc;
    B200_REQUIRE(packed && aligned16(packed) && depths && radii && num_tiles_hit, "null / misaligned output pointer");
    fused_forward_kernel<-1><<<ceil_div(num_points, FUSED_THREADS), FUSED_THREADS, 0, as_stream(stream)>>>(
        f, reinterpret_cast<PackedGaussian *>(packed), depths, radii, num_tiles_hit);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_fused_colors_forward(int num_points, const float *means, const float *sh_dc, const float *sh_rest,
                                         int sh_bases, int degrees_to_use, const float *cam_pos, const int32_t *radii,
                                         void *packed, void *stream) {
    B200_REQUIRE(num_points >= 1, "num_points must be >= 1");
    B200_REQUIRE(means && sh_dc && cam_pos && radii && packed && aligned16(packed), "null / misaligned pointer");
    B200_REQUIRE(sh_bases == 1 || sh_bases == 4 || sh_bases == 9 || sh_bases == 16 || sh_bases == 25, "bad SH basis count %d", sh_bases);
    B200_REQUIRE(degrees_to_use >= 0 && (degrees_to_use + 1) * (degrees_to_use + 1) <= sh_bases, "degrees_to_use too large");
    B200_REQUIRE(sh_bases == 1 || sh_rest, "null sh_rest");
    FusedParams f{};
    f.proj.n = num_points; f.proj.means = means;
    f.sh_dc = sh_dc; f.sh_rest = sh_rest; f.cam_pos = cam_pos; f.K = sh_bases; f.deg_use = degrees_to_use;
    const int blocks = ceil_div(num_points, FUSED_THREADS);
    cudaStream_t st = as_stream(stream);
    PackedGaussian *rec = reinterpret_cast<PackedGaussian *>(packed);
    switch (degrees_to_use) {
        case 0: fused_colors_kernel<0><<<blocks, FUSED_THREADS, 0, st>>>(f, radii, rec); break;
        case 1: fused_colors_kernel<1><<<blocks, FUSED_THREADS, 0, st>>>(f, radii, rec); break;
        case 2: fused_colors_kernel<2><<<blocks, FUSED_THREADS, 0, st>>>(f, radii, rec); break;
        case 3: fused_colors_kernel<3><<<blocks, FUSED_THREADS, 0, st>>>(f, radii, rec); break;
        default: fused_colors_kernel<4><<<blocks, FUSED_THREADS, 0, st>>>(f, radii, rec); break;
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_fused_preprocess_backward(int num_points, const float *means, const float *log_scales,
                                              const float *quats, const float *opacity_logit, const float *sh_dc,
                                              const float *sh_rest, int sh_bases, int degrees_to_use,
                                              const float *viewmat, const float *cam_pos, const float *lin_vel,
                                              const float *ang_vel, float rolling_shutter_time, float exposure_time,
                                              float fx, float fy, float cx, float cy, unsigned img_height,
                                              unsigned img_width, unsigned block_width, float clip_thresh,
                                              const void *packed, const int32_t *radii, const float *v_xy,
                                              const float *v_pix_vel, const float *v_conic, const float *v_colors,
                                              const float *v_opacity, float *g_means, float *g_log_scales,
                                              float *g_quats, float *g_opacity_logit, float *g_sh_dc, float *g_sh_rest,
                                              float *g_lin_vel, float *g_ang_vel, float *g_viewmat, void *stream) {
    FusedParams f;
    int rc = fill_fused(f, num_points, means, log_scales, quats, opacity_logit, sh_dc, sh_rest, sh_bases, degrees_to_use,
                        viewmat, cam_pos, lin_vel, ang_vel, rolling_shutter_time, exposure_time, fx, fy, cx, cy, img_height,
                        img_width, block_width, clip_thresh);
    if (rc) return rc;
    B200_REQUIRE(packed && radii && v_xy && v_pix_vel && v_conic && v_colors && v_opacity, "null input pointer");
    B200_REQUIRE(g_means && g_log_scales && g_quats && g_opacity_logit && g_sh_dc && (sh_bases == 1 || g_sh_rest),
                 "null gradient pointer");
    cudaStream_t st = as_stream(stream);
    if (g_lin_vel) B200_CUDA(cudaMemsetAsync(g_lin_vel, 0, 3 * sizeof(float), st));
    if (g_ang_vel) B200_CUDA(cudaMemsetAsync(g_ang_vel, 0, 3 * sizeof(float), st));
    if (g_viewmat) B200_CUDA(cudaMemsetAsync(g_viewmat, 0, 12 * sizeof(float), st));
    FusedBwdIO io{reinterpret_cast<const PackedGaussian *>(packed), radii, v_xy, v_pix_vel, v_conic, v_colors, v_opacity,
                  g_means, g_log_scales, g_quats, g_opacity_logit, g_sh_dc, g_sh_rest, g_lin_vel, g_ang_vel, g_viewmat,
                  aligned16(g_quats)};
    const int rest_row = 3 * (sh_bases - 1);
    const int pitch = (rest_row & 1) ? rest_row : rest_row + 1;
    const size_t smem = sizeof(float) * FUSED_THREADS * (size_t)(pitch > 0 ? pitch : 1);
    const int blocks = ceil_div(num_points, FUSED_THREADS);
    switch (degrees_to_use) {
        case 0: fused_backward_kernel<0><<<blocks, FUSED_THREADS, smem, st>>>(f, io); break;
        case 1: fused_backward_kernel<1><<<blocks, FUSED_THREADS, smem, st>>>(f, io); break;
        case 2: fused_backward_kernel<2><<<blocks, FUSED_THREADS, smem, st>>>(f, io); break;
        case 3: fused_backward_kernel<3><<<blocks, FUSED_THREADS, smem, st>>>(f, io); break;
        default: fused_backward_kernel<4><<<blocks, FUSED_THREADS, smem, st>>>(f, io); break;
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}
