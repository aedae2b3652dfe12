// Spherical-harmonics colour evaluation, forward and backward (semantics of
// /root/reference/gsplat/gsplat/cuda/csrc/sh.cuh:54-498).
//
// The op is a pure HBM stream: 12*K + 24 bytes per Gaussian each way (K = 16 -> 216 B).  The reference
// reads / writes its (N,K,3) rows with one thread striding through 192 B at a time; here a block moves its
// contiguous chunk with 16-byte coalesced accesses and transposes through padded shared memory (odd row
// pitch -> conflict-free per-thread reads), so DRAM sees full-line bursts only.
#include "common.cuh"

namespace b200 {

constexpr int SH_THREADS = 128;

__host__ __device__ inline int sh_num_bases(int degree) {
    return degree == 0 ? 1 : degree == 1 ? 4 : degree == 2 ? 9 : degree == 3 ? 16 : 25;  // sh.cuh:42-52
}
__host__ __device__ inline int sh_pitch(int row_floats) { return (row_floats & 1) ? row_floats : row_floats + 1; }

template <int METHOD>
__device__ __forceinline__ void sh_basis(int deg, float dx, float dy, float dz, float *B) {
    if (METHOD == B200_SH_FAST) {  // Sloan's recurrence, sh.cuh:54-156
        B[0] = 0.2820947917738781f;
        if (deg < 1) return;
        const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
        const float x = dx / nrm, y = dy / nrm, z = dz / nrm;
        const float a0 = 0.48860251190292f;
        B[1] = -a0 * y; B[2] = a0 * z; B[3] = -a0 * x;
        if (deg < 2) return;
        const float z2 = z * z;
        const float b0 = -1.092548430592079f * z, a1 = 0.5462742152960395f;
        const float c1 = x * x - y * y, s1 = 2.f * x * y;
        B[6] = 0.9461746957575601f * z2 - 0.3153915652525201f;
        B[7] = b0 * x; B[5] = b0 * y; B[8] = a1 * c1; B[4] = a1 * s1;
        if (deg < 3) return;
        const float c0 = -2.285228997322329f * z2 + 0.4570457994644658f;
        const float b1 = 1.445305721320277f * z, a2 = -0.5900435899266435f;
        const float c2 = x * c1 - y * s1, s2 = x * s1 + y * c1;
        B[12] = z * (1.865881662950577f * z2 - 1.119528997770346f);
        B[13] = c0 * x; B[11] = c0 * y; B[14] = b1 * c1; B[10] = b1 * s1; B[15] = a2 * c2; B[9] = a2 * s2;
        if (deg < 4) return;
        const float d0 = z * (-4.683325804901025f * z2 + 2.007139630671868f);
        const float cc = 3.31161143515146f * z2 - 0.47308734787878f;
        const float b2 = -1.770130769779931f * z, a3 = 0.6258357354491763f;
        const float c3 = x * c2 - y * s2, s3 = x * s2 + y * c2;
        B[20] = 1.984313483298443f * z * B[12] - 1.006230589874905f * B[6];
        B[21] = d0 * x; B[19] = d0 * y; B[22] = cc * c1; B[18] = cc * s1;
        B[23] = b2 * c2; B[17] = b2 * s2; B[24] = a3 * c3; B[16] = a3 * s3;
    } else {  // explicit polynomials, sh.cuh:268-340
        B[0] = 0.28209479177387814f;
        if (deg < 1) return;
        const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
        const float x = dx / nrm, y = dy / nrm, z = dz / nrm;
        const float xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
        const float C1 = 0.4886025119029199f;
        B[1] = -C1 * y; B[2] = C1 * z; B[3] = -C1 * x;
        if (deg < 2) return;
        B[4] = 1.0925484305920792f * xy; B[5] = -1.0925484305920792f * yz;
        B[6] = 0.31539156525252005f * (2.f * zz - xx - yy);
        B[7] = -1.0925484305920792f * xz; B[8] = 0.5462742152960396f * (xx - yy);
        if (deg < 3) return;
        B[9] = -0.5900435899266435f * y * (3.f * xx - yy); B[10] = 2.890611442640554f * xy * z;
        B[11] = -0.4570457994644658f * y * (4.f * zz - xx - yy);
        B[12] = 0.3731763325901154f * z * (2.f * zz - 3.f * xx - 3.f * yy);
        B[13] = -0.4570457994644658f * x * (4.f * zz - xx - yy); B[14] = 1.445305721320277f * z * (xx - yy);
        B[15] = -0.5900435899266435f * x * (xx - 3.f * yy);
        if (deg < 4) return;
        B[16] = 2.5033429417967046f * xy * (xx - yy); B[17] = -1.7701307697799304f * yz * (3.f * xx - yy);
        B[18] = 0.9461746957575601f * xy * (7.f * zz - 1.f); B[19] = -0.6690465435572892f * yz * (7.f * zz - 3.f);
        B[20] = 0.10578554691520431f * (zz * (35.f * zz - 30.f) + 3.f);
        B[21] = -0.6690465435572892f * xz * (7.f * zz - 3.f); B[22] = 0.47308734787878004f * (xx - yy) * (7.f * zz - 1.f);
        B[23] = -1.7701307697799304f * xz * (xx - 3.f * yy);
        B[24] = 0.6258357354491761f * (xx * (xx - 3.f * yy) - yy * (3.f * xx - yy));
    }
}

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
