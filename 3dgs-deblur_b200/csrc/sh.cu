// Spherical-harmonics colour evaluation, forward and backward (semantics of
// /root/reference/gsplat/gsplat/cuda/csrc/sh.cuh:54-498).
//
// The op is a pure HBM stream: 12*K + 24 bytes per Gaussian each way (K = 16 -> 216 B).  The reference
// reads / writes its (N,K,3) rows with one thread striding through 192 B at a time; here a block moves its
// contiguous chunk with 16-byte coalesced accesses and transposes through padded shared memory (odd row
// pitch -> conflict-free per-thread reads), so DRAM sees full-line bursts only.
#include "sh_math.cuh"

namespace b200 {

constexpr int SH_THREADS = 128;

// KU = number of bases actually used ((degrees_to_use+1)^2), compile-time so B[] stays in registers.
template <int METHOD, int DEG_USE, bool VEC>
__global__ void __launch_bounds__(SH_THREADS) sh_forward_kernel(int n, int K, const float *__restrict__ dirs,
                                                                const float *__restrict__ coeffs,
                                                                float *__restrict__ colors) {
    extern __shared__ __align__(16) float s_rows[];
    constexpr int KU = (DEG_USE + 1) * (DEG_USE + 1);
    const int row = 3 * K, pitch = sh_pitch(row);
    const int base = blockIdx.x * SH_THREADS;
    const int count = min(SH_THREADS, n - base);
    const float *src = coeffs + (size_t)base * row;
    const int nfloat = count * row;
    if (VEC) {
        const int nvec = nfloat >> 2;
        const float4 *src4 = reinterpret_cast<const float4 *>(src);
        for (int i = threadIdx.x; i < nvec; i += SH_THREADS) {
            const float4 v = __ldg(src4 + i);
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int f = 4 * i + j;
                const int r = f / row;
                s_rows[r * pitch + (f - r * row)] = e[j];
            }
        }
        for (int f = (nvec << 2) + threadIdx.x; f < nfloat; f += SH_THREADS) {
            const int r = f / row;
            s_rows[r * pitch + (f - r * row)] = __ldg(src + f);
        }
    } else {
        for (int f = threadIdx.x; f < nfloat; f += SH_THREADS) {
            const int r = f / row;
            s_rows[r * pitch + (f - r * row)] = __ldg(src + f);
        }
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= count) return;
    const int idx = base + t;
    float B[KU > 1 ? KU : 1];
    sh_basis<METHOD>(DEG_USE, dirs[3 * (size_t)idx], dirs[3 * (size_t)idx + 1], dirs[3 * (size_t)idx + 2], B);
    const float *my = s_rows + t * pitch;
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int k = 0; k < KU; ++k) {
        c0 = fmaf(B[k], my[3 * k], c0);
        c1 = fmaf(B[k], my[3 * k + 1], c1);
        c2 = fmaf(B[k], my[3 * k + 2], c2);
    }
    colors[3 * (size_t)idx] = c0; colors[3 * (size_t)idx + 1] = c1; colors[3 * (size_t)idx + 2] = c2;
}

template <int METHOD, int DEG_USE, bool VEC>
__global__ void __launch_bounds__(SH_THREADS) sh_backward_kernel(int n, int K, const float *__restrict__ dirs,
                                                                 const float *__restrict__ v_colors,
                                                                 float *__restrict__ v_coeffs) {
    extern __shared__ __align__(16) float s_rows[];
    constexpr int KU = (DEG_USE + 1) * (DEG_USE + 1);
    const int row = 3 * K, pitch = sh_pitch(row);
    const int base = blockIdx.x * SH_THREADS;
    const int count = min(SH_THREADS, n - base);
    const int t = threadIdx.x;
    if (t < count) {
        const int idx = base + t;
        float B[KU > 1 ? KU : 1];
        sh_basis<METHOD>(DEG_USE, dirs[3 * (size_t)idx], dirs[3 * (size_t)idx + 1], dirs[3 * (size_t)idx + 2], B);
        const float v0 = v_colors[3 * (size_t)idx], v1 = v_colors[3 * (size_t)idx + 1], v2 = v_colors[3 * (size_t)idx + 2];
        float *my = s_rows + t * pitch;
#pragma unroll
        for (int k = 0; k < KU; ++k) {
            my[3 * k] = B[k] * v0; my[3 * k + 1] = B[k] * v1; my[3 * k + 2] = B[k] * v2;
        }
        for (int f = 3 * KU; f < row; ++f) my[f] = 0.f;  // bases above degrees_to_use (bindings.cu:123-124)
    }
    __syncthreads();
    float *dst = v_coeffs + (size_t)base * row;
    const int nfloat = count * row;
    if (VEC) {
        const int nvec = nfloat >> 2;
        for (int i = threadIdx.x; i < nvec; i += SH_THREADS) {
            float e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int f = 4 * i + j;
                const int r = f / row;
                e[j] = s_rows[r * pitch + (f - r * row)];
            }
            reinterpret_cast<float4 *>(dst)[i] = make_float4(e[0], e[1], e[2], e[3]);
        }
        for (int f = (nvec << 2) + threadIdx.x; f < nfloat; f += SH_THREADS) {
            const int r = f / row;
            dst[f] = s_rows[r * pitch + (f - r * row)];
        }
    } else {
        for (int f = threadIdx.x; f < nfloat; f += SH_THREADS) {
            const int r = f / row;
            dst[f] = s_rows[r * pitch + (f - r * row)];
        }
    }
}

template <int METHOD, int DEG_USE>
static int launch_sh(bool fwd, int n, int K, const float *dirs, const float *in, float *out, cudaStream_t st) {
    const int blocks = ceil_div(n, SH_THREADS);
    const size_t smem = sizeof(float) * SH_THREADS * sh_pitch(3 * K);
    const bool vec = fwd ? aligned16(in) : aligned16(out);
    if (fwd) {
        if (vec) sh_forward_kernel<METHOD, DEG_USE, true><<<blocks, SH_THREADS, smem, st>>>(n, K, dirs, in, out);
        else sh_forward_kernel<METHOD, DEG_USE, false><<<blocks, SH_THREADS, smem, st>>>(n, K, dirs, in, out);
    } else {
        if (vec) sh_backward_kernel<METHOD, DEG_USE, true><<<blocks, SH_THREADS, smem, st>>>(n, K, dirs, in, out);
        else sh_backward_kernel<METHOD, DEG_USE, false><<<blocks, SH_THREADS, smem, st>>>(n, K, dirs, in, out);
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}

template <int METHOD>
static int dispatch_deg(bool fwd, int deg_use, int n, int K, const float *dirs, const float *in, float *out, cudaStream_t st) {
    switch (deg_use) {
        case 0: return launch_sh<METHOD, 0>(fwd, n, K, dirs, in, out, st);
        case 1: return launch_sh<METHOD, 1>(fwd, n, K, dirs, in, out, st);
        case 2: return launch_sh<METHOD, 2>(fwd, n, K, dirs, in, out, st);
        case 3: return launch_sh<METHOD, 3>(fwd, n, K, dirs, in, out, st);
        default: return launch_sh<METHOD, 4>(fwd, n, K, dirs, in, out, st);
    }
}

static int sh_entry(bool fwd, int method, int n, int degree, int deg_use, const float *dirs, const float *in,
                    float *out, void *stream) {
    B200_REQUIRE(method == B200_SH_POLY || method == B200_SH_FAST, "Invalid method: %d", method);  // bindings.cu:99-101
    B200_REQUIRE(n >= 1, "num_points must be >= 1");
    B200_REQUIRE(degree >= 0 && degree <= 4, "degree must be in [0,4] (got %d)", degree);
    B200_REQUIRE(deg_use >= 0 && deg_use <= degree, "degrees_to_use (%d) must be in [0, degree=%d]", deg_use, degree);
    B200_REQUIRE(dirs && in && out, "null pointer");
    const int K = sh_num_bases(degree);
    if (method == B200_SH_FAST) return dispatch_deg<B200_SH_FAST>(fwd, deg_use, n, K, dirs, in, out, as_stream(stream));
    return dispatch_deg<B200_SH_POLY>(fwd, deg_use, n, K, dirs, in, out, as_stream(stream));
}

}  // namespace b200

extern "C" int b200_compute_sh_forward(int method, int num_points, int degree, int degrees_to_use,
                                       const float *viewdirs, const float *coeffs, float *colors, void *stream) {
    return b200::sh_entry(true, method, num_points, degree, degrees_to_use, viewdirs, coeffs, colors, stream);
}

extern "C" int b200_compute_sh_backward(int method, int num_points, int degree, int degrees_to_use,
                                        const float *viewdirs, const float *v_colors, float *v_coeffs, void *stream) {
    return b200::sh_entry(false, method, num_points, degree, degrees_to_use, viewdirs, v_colors, v_coeffs, stream);
}
