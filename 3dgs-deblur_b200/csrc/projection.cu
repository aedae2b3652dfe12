// Fused projection forward / backward for sm_100a.
//
// forward : view transform + near clip, cov3d = (RS)(RS)^T, EWA cov2d (+0.3 I, fov clamp,
//           compensation), conic + 3-sigma radius, pixel mean, pixel velocity, blur-inflated
//           tile bbox  -- semantics of /root/reference/gsplat/gsplat/cuda/csrc/forward.cu:13-112.
// backward: VJP of all of the above in one kernel, plus (new) dL/d(lin_vel), dL/d(ang_vel) and
//           dL/d(viewmat) block-reduced in the same pass -- reference backward.cu:371-572 has
//           none of the three; its Python layer adds an approximate viewmat gradient with nine
//           torch.dot launches (project_gaussians.py:272-307) and falls back to ~60 torch ops when
//           velocities need grad (project_gaussians.py:81-112).
//
// Both kernels are HBM-streaming (108 B / 160 B algorithmic per Gaussian): one thread per Gaussian,
// the block's AoS input chunk is fetched with coalesced 16-byte loads into shared memory and read
// back with a conflict-free stride of 3 words.
#include "common.cuh"

namespace b200 {

struct ProjCommon {
    int n;
    const float *means, *scales, *quats;
    const float *lin_vel, *ang_vel;  // device float[3] or nullptr
    const float *viewmat;            // device, >= 12 floats
    float glob_scale, rs_time, exposure;
    float fx, fy, cx, cy;
    int H, W, bw;
    float clip;
};

constexpr int PROJ_THREADS = 256;

// Cooperative load of `count` AoS rows of `WIDTH` floats starting at row `base` into smem.
template <int WIDTH, bool VEC>
__device__ __forceinline__ void stage_rows(float *s, const float *g, int base, int count) {
    const int nfloat = WIDTH * count;
    const float *src = g + (size_t)WIDTH * base;
    if (VEC) {
        const int nvec = nfloat >> 2;
        const float4 *src4 = reinterpret_cast<const float4 *>(src);
        for (int i = threadIdx.x; i < nvec; i += PROJ_THREADS) reinterpret_cast<float4 *>(s)[i] = __ldg(src4 + i);
        for (int i = (nvec << 2) + threadIdx.x; i < nfloat; i += PROJ_THREADS) s[i] = __ldg(src + i);
    } else {
        for (int i = threadIdx.x; i < nfloat; i += PROJ_THREADS) s[i] = __ldg(src + i);
    }
}

struct ProjFwdOut {
    float *cov3d, *xys, *depths, *pix_vels, *conics, *comp;
    int32_t *radii, *tiles_hit;
    int32_t *quat_flag;  // optional: set to 1 if any |q| - 1 >= 1e-6 (the reference's Python-side assert, fused)
};

template <bool VEC>
__global__ void __launch_bounds__(PROJ_THREADS) project_forward_kernel(ProjCommon p, ProjFwdOut o) {
    __shared__ __align__(16) float s_means[3 * PROJ_THREADS];
    __shared__ __align__(16) float s_scales[3 * PROJ_THREADS];
    __shared__ __align__(16) float s_quats[4 * PROJ_THREADS];
    const int base = blockIdx.x * PROJ_THREADS;
    const int count = min(PROJ_THREADS, p.n - base);
    stage_rows<3, VEC>(s_means, p.means, base, count);
    stage_rows<3, VEC>(s_scales, p.scales, base, count);
    stage_rows<4, VEC>(s_quats, p.quats, base, count);
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= count) return;
    const int idx = base + t;

    if (o.quat_flag) {  // project_gaussians.py:69: assert (quats.norm(dim=-1) - 1 < 1e-6).all()
        const float qw = s_quats[4 * t], qx = s_quats[4 * t + 1], qy = s_quats[4 * t + 2], qz = s_quats[4 * t + 3];
        if (!(sqrtf(qw * qw + qx * qx + qy * qy + qz * qz) - 1.f < 1e-6f)) atomicOr(o.quat_flag, 1);
    }

    float vm[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) vm[i] = __ldg(p.viewmat + i);

    // outputs default to the zeros the reference gets from torch::zeros (bindings.cu:208-223)
    float cov3d[6] = {0, 0, 0, 0, 0, 0}, conic[3] = {0, 0, 0}, xy[2] = {0, 0}, vel[2] = {0, 0};
    float depth = 0.f, comp = 0.f;
    int radius_i = 0, tiles = 0;

    const float px = s_means[3 * t], py = s_means[3 * t + 1], pz = s_means[3 * t + 2];
    const float vx = vm[0] * px + vm[1] * py + vm[2] * pz + vm[3];
    const float vy = vm[4] * px + vm[5] * py + vm[6] * pz + vm[7];
    const float vz = vm[8] * px + vm[9] * py + vm[10] * pz + vm[11];
    if (vz > p.clip) {  // forward.cu:49
        float R[9];
        quat_to_rotmat(s_quats[4 * t], s_quats[4 * t + 1], s_quats[4 * t + 2], s_quats[4 * t + 3], R);
        const float s0 = p.glob_scale * s_scales[3 * t], s1 = p.glob_scale * s_scales[3 * t + 1],
                    s2 = p.glob_scale * s_scales[3 * t + 2];
        float M[9];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            M[3 * r] = R[3 * r] * s0; M[3 * r + 1] = R[3 * r + 1] * s1; M[3 * r + 2] = R[3 * r + 2] * s2;
        }
        float V[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                V[3 * r + c] = M[3 * r] * M[3 * c] + M[3 * r + 1] * M[3 * c + 1] + M[3 * r + 2] * M[3 * c + 2];
        cov3d[0] = V[0]; cov3d[1] = V[1]; cov3d[2] = V[2]; cov3d[3] = V[4]; cov3d[4] = V[5]; cov3d[5] = V[8];

        // EWA with the 1.3*tan(fov) clamp (forward.cu:459-534)
        const float limx = 1.3f * (0.5f * (float)p.W / p.fx), limy = 1.3f * (0.5f * (float)p.H / p.fy);
        const float tx = vz * fminf(limx, fmaxf(-limx, vx / vz));
        const float ty = vz * fminf(limy, fmaxf(-limy, vy / vz));
        const float rz = 1.f / vz, rz2 = rz * rz;
        const float J00 = p.fx * rz, J02 = -p.fx * tx * rz2, J11 = p.fy * rz, J12 = -p.fy * ty * rz2;
        float T[6];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            T[c] = J00 * vm[c] + J02 * vm[8 + c];
            T[3 + c] = J11 * vm[4 + c] + J12 * vm[8 + c];
        }
        float TV[6];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                TV[3 * r + c] = T[3 * r] * V[c] + T[3 * r + 1] * V[3 + c] + T[3 * r + 2] * V[6 + c];
        const float c00 = TV[0] * T[0] + TV[1] * T[1] + TV[2] * T[2];
        const float c01 = TV[0] * T[3] + TV[1] * T[4] + TV[2] * T[5];
        const float c11 = TV[3] * T[3] + TV[4] * T[4] + TV[5] * T[5];
        const float det_orig = c00 * c11 - c01 * c01;
        const float a = c00 + 0.3f, b = c01, c = c11 + 0.3f;
        const float det_blur = a * c - b * b;
        const float compensation = sqrtf(fmaxf(0.f, det_orig / det_blur));

        float radius;
        if (cov2d_to_conic_radius(a, b, c, conic[0], conic[1], conic[2], radius)) {  // forward.cu:75-79
            const float rw = 1.f / (vz + 1e-6f);  // helpers.cuh:128-135
            const float mx = vx * rw * p.fx + p.cx, my = vy * rw * p.fy + p.cy;
            if (p.rs_time > 0.f || p.exposure > 0.f) {  // helpers.cuh:224-253, forward.cu:88-91
                float lv[3] = {0, 0, 0}, av[3] = {0, 0, 0};
                if (p.lin_vel) { lv[0] = __ldg(p.lin_vel); lv[1] = __ldg(p.lin_vel + 1); lv[2] = __ldg(p.lin_vel + 2); }
                if (p.ang_vel) { av[0] = __ldg(p.ang_vel); av[1] = __ldg(p.ang_vel + 1); av[2] = __ldg(p.ang_vel + 2); }
                const float t0 = lv[0] + (av[1] * vz - av[2] * vy);
                const float t1 = lv[1] + (av[2] * vx - av[0] * vz);
                const float t2 = lv[2] + (av[0] * vy - av[1] * vx);
                const float z1 = 1.f / vz, z2 = z1 * z1;
                vel[0] = -(p.fx * z1 * t0 + (-p.fx * vx * z2) * t2);
                vel[1] = -(p.fy * z1 * t1 + (-p.fy * vy * z2) * t2);
                radius = (float)((double)radius + (double)sqrtf(vel[0] * vel[0] + vel[1] * vel[1]) * 0.5 *
                                                      (double)(p.exposure + p.rs_time));
            }
            int x0, y0, x1, y1;
            tile_bbox(mx, my, radius, (p.W + p.bw - 1) / p.bw, (p.H + p.bw - 1) / p.bw, (float)p.bw, x0, y0, x1, y1);
            const int area = (x1 - x0) * (y1 - y0);
            if (area > 0) {
                tiles = area; depth = vz; radius_i = (int)radius; xy[0] = mx; xy[1] = my; comp = compensation;
            }
        }
    }
    float2 *c3 = reinterpret_cast<float2 *>(o.cov3d + 6 * (size_t)idx);
    c3[0] = make_float2(cov3d[0], cov3d[1]); c3[1] = make_float2(cov3d[2], cov3d[3]); c3[2] = make_float2(cov3d[4], cov3d[5]);
    reinterpret_cast<float2 *>(o.xys)[idx] = make_float2(xy[0], xy[1]);
    reinterpret_cast<float2 *>(o.pix_vels)[idx] = make_float2(vel[0], vel[1]);
    o.conics[3 * (size_t)idx] = conic[0]; o.conics[3 * (size_t)idx + 1] = conic[1]; o.conics[3 * (size_t)idx + 2] = conic[2];
    o.depths[idx] = depth; o.comp[idx] = comp; o.radii[idx] = radius_i; o.tiles_hit[idx] = tiles;
}

// ------------------------------------------------------------------------------ backward

struct ProjBwdIO {
    const float *cov3d, *conics, *comp;
    const int32_t *radii;
    const float *v_xy, *v_depth, *v_pix_vel, *v_conic, *v_comp;
    float *v_cov2d, *v_cov3d;  // optional
    float *v_mean, *v_scale, *v_quat;
    float *v_lin, *v_ang, *v_viewmat;  // optional accumulators (pre-zeroed)
};

template <bool VEC, bool EXACT>
__global__ void __launch_bounds__(PROJ_THREADS) project_backward_kernel(ProjCommon p, ProjBwdIO io) {
    __shared__ __align__(16) float s_means[3 * PROJ_THREADS];
    __shared__ __align__(16) float s_scales[3 * PROJ_THREADS];
    __shared__ __align__(16) float s_quats[4 * PROJ_THREADS];
    __shared__ float s_red[PROJ_THREADS / 32][18];
    const int base = blockIdx.x * PROJ_THREADS;
    const int count = min(PROJ_THREADS, p.n - base);
    stage_rows<3, VEC>(s_means, p.means, base, count);
    stage_rows<3, VEC>(s_scales, p.scales, base, count);
    stage_rows<4, VEC>(s_quats, p.quats, base, count);
    __syncthreads();
    const int t = threadIdx.x;
    const int idx = base + t;
    const bool want_cam = io.v_lin || io.v_ang || io.v_viewmat;

    float red[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) red[i] = 0.f;
    float v_mean[3] = {0, 0, 0}, v_scale[3] = {0, 0, 0}, v_quat[4] = {0, 0, 0, 0};
    float vc2[3] = {0, 0, 0}, v_c3[6] = {0, 0, 0, 0, 0, 0};

    const bool active = t < count && io.radii[idx] > 0;  // backward.cu:400
    if (active) {
        float vm[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) vm[i] = __ldg(p.viewmat + i);
        const float px = s_means[3 * t], py = s_means[3 * t + 1], pz = s_means[3 * t + 2];
        const float vx = vm[0] * px + vm[1] * py + vm[2] * pz + vm[3];
        const float vy = vm[4] * px + vm[5] * py + vm[6] * pz + vm[7];
        const float vz = vm[8] * px + vm[9] * py + vm[10] * pz + vm[11];

        // ---- dL/d(p_view) from the pixel velocity, the pixel mean and the depth
        float vpv[3] = {0.f, 0.f, 0.f};
        if (p.rs_time > 0.f || p.exposure > 0.f) {  // helpers.cuh:255-326
            const float2 g = reinterpret_cast<const float2 *>(io.v_pix_vel)[idx];
            float lv[3] = {0, 0, 0}, av[3] = {0, 0, 0};
            if (p.lin_vel) { lv[0] = __ldg(p.lin_vel); lv[1] = __ldg(p.lin_vel + 1); lv[2] = __ldg(p.lin_vel + 2); }
            if (p.ang_vel) { av[0] = __ldg(p.ang_vel); av[1] = __ldg(p.ang_vel + 1); av[2] = __ldg(p.ang_vel + 2); }
            const float t0 = lv[0] + (av[1] * vz - av[2] * vy);
            const float t1 = lv[1] + (av[2] * vx - av[0] * vz);
            const float t2 = lv[2] + (av[0] * vy - av[1] * vx);
            const float z1 = 1.f / vz, z2 = z1 * z1, z3 = z2 * z1;
            vpv[0] = g.x * p.fx * z2 * t2;
            vpv[1] = g.y * p.fy * z2 * t2;
            vpv[2] = -(g.x * (-p.fx * z2 * t0 + 2.f * p.fx * vx * z3 * t2) + g.y * (-p.fy * z2 * t1 + 2.f * p.fy * vy * z3 * t2));
            // dL/d(total velocity) = -J^T g ; total = lin + ang x p_view
            const float w0 = -(p.fx * z1 * g.x), w1 = -(p.fy * z1 * g.y), w2 = p.fx * vx * z2 * g.x + p.fy * vy * z2 * g.y;
            vpv[0] -= av[1] * w2 - av[2] * w1;
            vpv[1] -= av[2] * w0 - av[0] * w2;
            vpv[2] -= av[0] * w1 - av[1] * w0;
            if (want_cam) {
                red[0] = w0; red[1] = w1; red[2] = w2;  // dL/d lin_vel
                red[3] = vy * w2 - vz * w1;             // dL/d ang_vel = p_view x w
                red[4] = vz * w0 - vx * w2;
                red[5] = vx * w1 - vy * w0;
            }
        }
        {
            const float2 g = reinterpret_cast<const float2 *>(io.v_xy)[idx];
            const float rw = 1.f / (vz + 1e-6f);  // helpers.cuh:138-147
            const float gx = p.fx * g.x, gy = p.fy * g.y;
            vpv[0] += gx * rw;
            vpv[1] += gy * rw;
            vpv[2] += -(gx * vx + gy * vy) * rw * rw + io.v_depth[idx];
        }

        // ---- conic + compensation -> cov2d (helpers.cuh:68-94)
        const float ca = io.conics[3 * (size_t)idx], cb = io.conics[3 * (size_t)idx + 1], cc = io.conics[3 * (size_t)idx + 2];
        {
            const float g0 = io.v_conic[3 * (size_t)idx], g1 = 0.5f * io.v_conic[3 * (size_t)idx + 1], g2 = io.v_conic[3 * (size_t)idx + 2];
            const float xg00 = ca * g0 + cb * g1, xg01 = ca * g1 + cb * g2, xg10 = cb * g0 + cc * g1, xg11 = cb * g1 + cc * g2;
            const float S00 = -(xg00 * ca + xg01 * cb), S01 = -(xg00 * cb + xg01 * cc);
            const float S10 = -(xg10 * ca + xg11 * cb), S11 = -(xg10 * cb + xg11 * cc);
            vc2[0] = S00; vc2[1] = S01 + S10; vc2[2] = S11;
            const float comp = io.comp[idx];
            const float inv_det = ca * cc - cb * cb;
            const float om = 1.f - comp * comp;
            const float vsq = io.v_comp[idx] * 0.5f / (comp + 1e-6f);
            vc2[0] += vsq * (om * ca - 0.3f * inv_det);
            vc2[1] += 2.f * vsq * (om * cb);
            vc2[2] += vsq * (om * cc - 0.3f * inv_det);
        }

        // ---- EWA vjp (backward.cu:454-532).  EXACT: differentiate through the fov clamp like the torch path.
        float tx = vx, ty = vy;
        bool clx = false, cly = false;
        float sgx = 0.f, sgy = 0.f;
        if (EXACT) {
            const float limx = 1.3f * (0.5f * (float)p.W / p.fx), limy = 1.3f * (0.5f * (float)p.H / p.fy);
            const float qx = vx / vz, qy = vy / vz;
            if (qx > limx) { clx = true; sgx = limx; } else if (qx < -limx) { clx = true; sgx = -limx; }
            if (qy > limy) { cly = true; sgy = limy; } else if (qy < -limy) { cly = true; sgy = -limy; }
            if (clx) tx = vz * sgx;
            if (cly) ty = vz * sgy;
        }
        const float rz = 1.f / vz, rz2 = rz * rz, rz3 = rz2 * rz;
        const float J00 = p.fx * rz, J02 = -p.fx * tx * rz2, J11 = p.fy * rz, J12 = -p.fy * ty * rz2;
        float T[6];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            T[c] = J00 * vm[c] + J02 * vm[8 + c];
            T[3 + c] = J11 * vm[4 + c] + J12 * vm[8 + c];
        }
        const float *c3 = io.cov3d + 6 * (size_t)idx;
        const float V[9] = {c3[0], c3[1], c3[2], c3[1], c3[3], c3[4], c3[2], c3[4], c3[5]};
        const float vC[4] = {vc2[0], 0.5f * vc2[1], 0.5f * vc2[1], vc2[2]};
        float vCT[6];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) vCT[3 * r + c] = vC[2 * r] * T[c] + vC[2 * r + 1] * T[3 + c];
        float vV[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) vV[3 * r + c] = T[r] * vCT[c] + T[3 + r] * vCT[3 + c];
        v_c3[0] = vV[0]; v_c3[1] = vV[1] + vV[3]; v_c3[2] = vV[2] + vV[6];
        v_c3[3] = vV[4]; v_c3[4] = vV[5] + vV[7]; v_c3[5] = vV[8];
        float vT[6];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                vT[3 * r + c] = 2.f * (vCT[3 * r] * V[c] + vCT[3 * r + 1] * V[3 + c] + vCT[3 * r + 2] * V[6 + c]);
        float vJ[6];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                vJ[3 * r + c] = vT[3 * r] * vm[4 * c] + vT[3 * r + 1] * vm[4 * c + 1] + vT[3 * r + 2] * vm[4 * c + 2];
        {
            const float v_tx = -p.fx * rz2 * vJ[2], v_ty = -p.fy * rz2 * vJ[5];
            float v_z = -p.fx * rz2 * vJ[0] + 2.f * p.fx * tx * rz3 * vJ[2] - p.fy * rz2 * vJ[4] + 2.f * p.fy * ty * rz3 * vJ[5];
            float v_x = v_tx, v_y = v_ty;
            if (EXACT) {
                if (clx) { v_x = 0.f; v_z += sgx * v_tx; }
                if (cly) { v_y = 0.f; v_z += sgy * v_ty; }
            }
            vpv[0] += v_x; vpv[1] += v_y; vpv[2] += v_z;
        }
        // dL/d mean = W^T dL/d p_view  (helpers.cuh:97-104, backward.cu:529-531)
#pragma unroll
        for (int c = 0; c < 3; ++c) v_mean[c] = vm[c] * vpv[0] + vm[4 + c] * vpv[1] + vm[8 + c] * vpv[2];

        if (io.v_viewmat) {
            // d/dW through p_view = W p + t : outer(dL/dp_view, p); d/dt = dL/dp_view.
            float gcam[3] = {vpv[0], vpv[1], vpv[2]};
            if (!EXACT) {
                // reference CUDA-path approximation: v_cam = R v_mean (project_gaussians.py:295)
#pragma unroll
                for (int r = 0; r < 3; ++r) gcam[r] = vm[4 * r] * v_mean[0] + vm[4 * r + 1] * v_mean[1] + vm[4 * r + 2] * v_mean[2];
            }
            const float pw[3] = {px, py, pz};
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) red[6 + 4 * r + c] = gcam[r] * pw[c];
                red[6 + 4 * r + 3] = gcam[r];
            }
            if (EXACT) {
                // + J^T vT : the rotation's effect on the projected covariance (T = J W)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    red[6 + c] += J00 * vT[c];
                    red[6 + 4 + c] += J11 * vT[3 + c];
                    red[6 + 8 + c] += J02 * vT[c] + J12 * vT[3 + c];
                }
            }
        }

        // ---- cov3d -> scale, quat (backward.cu:536-572)
        float R[9];
        const float qw = s_quats[4 * t], qx_ = s_quats[4 * t + 1], qy_ = s_quats[4 * t + 2], qz_ = s_quats[4 * t + 3];
        quat_to_rotmat(qw, qx_, qy_, qz_, R);
        const float s[3] = {p.glob_scale * s_scales[3 * t], p.glob_scale * s_scales[3 * t + 1], p.glob_scale * s_scales[3 * t + 2]};
        float M[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) M[3 * r + c] = R[3 * r + c] * s[c];
        const float Vs[9] = {v_c3[0], 0.5f * v_c3[1], 0.5f * v_c3[2], 0.5f * v_c3[1], v_c3[3], 0.5f * v_c3[4], 0.5f * v_c3[2], 0.5f * v_c3[4], v_c3[5]};
        float vM[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                vM[3 * r + c] = 2.f * (Vs[3 * r] * M[c] + Vs[3 * r + 1] * M[3 + c] + Vs[3 * r + 2] * M[6 + c]);
#pragma unroll
        for (int c = 0; c < 3; ++c) v_scale[c] = (R[c] * vM[c] + R[3 + c] * vM[3 + c] + R[6 + c] * vM[6 + c]) * p.glob_scale;
        float g[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) g[3 * r + c] = vM[3 * r + c] * s[c];
        v_quat[0] = 2.f * (qx_ * (g[7] - g[5]) + qy_ * (g[2] - g[6]) + qz_ * (g[3] - g[1]));
        v_quat[1] = 2.f * (-2.f * qx_ * (g[4] + g[8]) + qy_ * (g[3] + g[1]) + qz_ * (g[6] + g[2]) + qw * (g[7] - g[5]));
        v_quat[2] = 2.f * (qx_ * (g[3] + g[1]) - 2.f * qy_ * (g[0] + g[8]) + qz_ * (g[7] + g[5]) + qw * (g[2] - g[6]));
        v_quat[3] = 2.f * (qx_ * (g[6] + g[2]) + qy_ * (g[7] + g[5]) - 2.f * qz_ * (g[0] + g[4]) + qw * (g[3] - g[1]));
    }

    if (t < count) {
        io.v_mean[3 * (size_t)idx] = v_mean[0]; io.v_mean[3 * (size_t)idx + 1] = v_mean[1]; io.v_mean[3 * (size_t)idx + 2] = v_mean[2];
        io.v_scale[3 * (size_t)idx] = v_scale[0]; io.v_scale[3 * (size_t)idx + 1] = v_scale[1]; io.v_scale[3 * (size_t)idx + 2] = v_scale[2];
        reinterpret_cast<float4 *>(io.v_quat)[idx] = make_float4(v_quat[0], v_quat[1], v_quat[2], v_quat[3]);
        if (io.v_cov2d) { io.v_cov2d[3 * (size_t)idx] = vc2[0]; io.v_cov2d[3 * (size_t)idx + 1] = vc2[1]; io.v_cov2d[3 * (size_t)idx + 2] = vc2[2]; }
        if (io.v_cov3d) {
#pragma unroll
            for (int i = 0; i < 6; ++i) io.v_cov3d[6 * (size_t)idx + i] = v_c3[i];
        }
    }

    if (want_cam) {  // block reduction of the 6 velocity + 12 view-matrix partials, one atomic per block each
        const int lane = t & 31, warp = t >> 5;
#pragma unroll
        for (int i = 0; i < 18; ++i) {
            const float s = warp_sum(red[i]);
            if (lane == 0) s_red[warp][i] = s;
        }
        __syncthreads();
        if (t < 18) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < PROJ_THREADS / 32; ++w) s += s_red[w][t];
            float *dst = t < 3 ? (io.v_lin ? io.v_lin + t : nullptr)
                       : t < 6 ? (io.v_ang ? io.v_ang + (t - 3) : nullptr)
                               : (io.v_viewmat ? io.v_viewmat + (t - 6) : nullptr);
            if (dst && s != 0.f) atomicAdd(dst, s);
        }
    }
}

__global__ void cov2d_bounds_kernel(int n, const float *__restrict__ cov2d, float *__restrict__ conics,
                                    float *__restrict__ radii) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float ca = 0.f, cb = 0.f, cc = 0.f, r = 0.f;
    cov2d_to_conic_radius(cov2d[3 * i], cov2d[3 * i + 1], cov2d[3 * i + 2], ca, cb, cc, r);
    conics[3 * i] = ca; conics[3 * i + 1] = cb; conics[3 * i + 2] = cc;
    radii[i] = r;
}

}  // namespace b200

using namespace b200;

static int fill_common(ProjCommon &p, int n, const float *means, const float *scales, float glob_scale,
                       const float *quats, const float *lin_vel, const float *ang_vel, float rs, float exposure,
                       const float *viewmat, float fx, float fy, float cx, float cy, unsigned H, unsigned W,
                       unsigned bw, float clip) {
    B200_REQUIRE(n >= 1, "num_points must be >= 1 (got %d)", n);  // project_gaussians.py:160-161
    B200_REQUIRE(means && scales && quats && viewmat, "null input pointer");
    B200_REQUIRE(bw > 1 && bw <= 16, "block_width must be between 2 and 16 (got %u)", bw);
    B200_REQUIRE(H > 0 && W > 0, "image size must be positive");
    p = ProjCommon{n, means, scales, quats, lin_vel, ang_vel, viewmat, glob_scale, rs, exposure, fx, fy, cx, cy, (int)H, (int)W, (int)bw, clip};
    return B200_OK;
}

extern "C" int b200_project_gaussians_forward(int num_points, const float *means3d, const float *scales,
                                              float glob_scale, const float *quats, const float *lin_vel,
                                              const float *ang_vel, float rolling_shutter_time, float exposure_time,
                                              const float *viewmat, float fx, float fy, float cx, float cy,
                                              unsigned img_height, unsigned img_width, unsigned block_width,
                                              float clip_thresh, float *cov3d, float *xys, float *depths,
                                              float *pix_vels, int32_t *radii, float *conics, float *compensation,
                                              int32_t *num_tiles_hit, int32_t *quat_norm_flag, void *stream) {
    ProjCommon p;
    int rc = fill_common(p, num_points, means3d, scales, glob_scale, quats, lin_vel, ang_vel, rolling_shutter_time,
                         exposure_time, viewmat, fx, fy, cx, cy, img_height, img_width, block_width, clip_thresh);
    if (rc) return rc;
    B200_REQUIRE(cov3d && xys && depths && pix_vels && radii && conics && compensation && num_tiles_hit, "null output pointer");
    ProjFwdOut o{cov3d, xys, depths, pix_vels, conics, compensation, radii, num_tiles_hit, quat_norm_flag};
    const int blocks = ceil_div(num_points, PROJ_THREADS);
    const bool vec = aligned16(means3d) && aligned16(scales) && aligned16(quats);
    if (vec) project_forward_kernel<true><<<blocks, PROJ_THREADS, 0, as_stream(stream)>>>(p, o);
    else project_forward_kernel<false><<<blocks, PROJ_THREADS, 0, as_stream(stream)>>>(p, o);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_project_gaussians_backward(int num_points, const float *means3d, const float *scales,
                                               float glob_scale, const float *quats, const float *lin_vel,
                                               const float *ang_vel, float rolling_shutter_time, float exposure_time,
                                               const float *viewmat, float fx, float fy, float cx, float cy,
                                               unsigned img_height, unsigned img_width, const float *cov3d,
                                               const int32_t *radii, const float *conics, const float *compensation,
                                               const float *v_xy, const float *v_depth, const float *v_pix_vel,
                                               const float *v_conic, const float *v_compensation, unsigned flags,
                                               float *v_cov2d, float *v_cov3d, float *v_mean3d, float *v_scale,
                                               float *v_quat, float *v_lin_vel, float *v_ang_vel, float *v_viewmat,
                                               void *stream) {
    ProjCommon p;
    int rc = fill_common(p, num_points, means3d, scales, glob_scale, quats, lin_vel, ang_vel, rolling_shutter_time,
                         exposure_time, viewmat, fx, fy, cx, cy, img_height, img_width, 16, 0.f);
    if (rc) return rc;
    B200_REQUIRE(cov3d && radii && conics && compensation, "null saved-forward pointer");
    B200_REQUIRE(v_xy && v_depth && v_pix_vel && v_conic && v_compensation, "null cotangent pointer");
    B200_REQUIRE(v_mean3d && v_scale && v_quat, "null gradient output pointer");
    B200_REQUIRE(aligned16(v_quat), "v_quat must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    if (v_lin_vel) B200_CUDA(cudaMemsetAsync(v_lin_vel, 0, 3 * sizeof(float), st));
    if (v_ang_vel) B200_CUDA(cudaMemsetAsync(v_ang_vel, 0, 3 * sizeof(float), st));
    if (v_viewmat) B200_CUDA(cudaMemsetAsync(v_viewmat, 0, 12 * sizeof(float), st));
    ProjBwdIO io{cov3d, conics, compensation, radii, v_xy, v_depth, v_pix_vel, v_conic, v_compensation,
                 v_cov2d, v_cov3d, v_mean3d, v_scale, v_quat, v_lin_vel, v_ang_vel, v_viewmat};
    const int blocks = ceil_div(num_points, PROJ_THREADS);
    const bool vec = aligned16(means3d) && aligned16(scales) && aligned16(quats);
    const bool exact = (flags & B200_PROJ_EXACT) != 0;
    if (vec && exact) project_backward_kernel<true, true><<<blocks, PROJ_THREADS, 0, st>>>(p, io);
    else if (vec) project_backward_kernel<true, false><<<blocks, PROJ_THREADS, 0, st>>>(p, io);
    else if (exact) project_backward_kernel<false, true><<<blocks, PROJ_THREADS, 0, st>>>(p, io);
    else project_backward_kernel<false, false><<<blocks, PROJ_THREADS, 0, st>>>(p, io);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_compute_cov2d_bounds(int num_pts, const float *cov2d, float *conics, float *radii, void *stream) {
    B200_REQUIRE(num_pts > 0, "num_pts must be positive");
    B200_REQUIRE(cov2d && conics && radii, "null pointer");
    cov2d_bounds_kernel<<<ceil_div(num_pts, 256), 256, 0, as_stream(stream)>>>(num_pts, cov2d, conics, radii);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
