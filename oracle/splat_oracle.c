/*
 * splat_oracle.c -- CPU restatement of the reference rasterizer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product path (3dgs-deblur_b200/) never links or calls it.
 *
 * Every function restates, in plain fp32 C, the semantics of one kernel of the
 * reference's vendored gsplat fork (paths relative to
 * /root/reference/gsplat/gsplat/cuda/csrc/):
 *
 *   orc_project_forward        forward.cu:13-112, :459-557, helpers.cuh:7-65,:107-135,:224-253
 *   orc_project_backward       backward.cu:371-572, helpers.cuh:68-94,:138-147,:170-211,:255-326
 *   orc_sh_forward/backward    sh.cuh:54-265 ("fast"), sh.cuh:268-432 ("poly"), :434-498
 *   orc_cov2d_bounds           bindings.cu:19-37, helpers.cuh:42-65
 *   orc_map_intersects         forward.cu:116-153
 *   orc_sort_intersects        gsplat/utils.py:179-180 (torch.sort + gather; here: stable)
 *   orc_tile_bin_edges         forward.cu:158-180
 *   orc_rasterize_forward      forward.cu:306-456   (blur + rolling shutter blend)
 *   orc_rasterize_backward     backward.cu:143-369
 *   orc_nd_rasterize_forward   forward.cu:185-304   (fp16 accumulators)
 *   orc_nd_rasterize_backward  backward.cu:22-141
 *
 * Parity pin: projection / SH / map / bin edges are checked against the
 * reference's own torch implementation (gsplat/_torch_impl.py) through the
 * golden vectors in tests/golden/ (made by tests/golden/make_golden.py, which
 * imports the reference in the build container), and every function incl. the
 * blend is checked against the UNMODIFIED reference CUDA extension
 * (oracle/_ref, built by oracle/build_ref.py) on the GPU box
 * (tests/test_ref_cuda_pin.py, fixtures in tests/golden/refcuda_*.npz).
 *
 * Arithmetic is fp32 with contraction off (gcc -ffp-contract=off); the CUDA
 * reference contracts to FMA and (AOT build) uses fast-math, so agreement is to
 * float tolerance, not bitwise, except for integer outputs.  Gradient sums are
 * accumulated in double (the reference uses unordered fp32 atomics).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_BLUR 10 /* helpers.cuh:222 */

/* thread control for the timed CPU baseline (bench.py): launchers such as torchrun export OMP_NUM_THREADS=1 */
int orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* ------------------------------------------------------------------ helpers */

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* helpers.cuh:7-40 : tile bbox, C truncation then clamp to [0, bounds] */
static void tile_bbox(float cx, float cy, float radius, int tbx, int tby, int bw,
                      int *minx, int *miny, int *maxx, int *maxy) {
    float tcx = cx / (float)bw, tcy = cy / (float)bw, tr = radius / (float)bw;
    *minx = clampi((int)(tcx - tr), 0, tbx);
    *maxx = clampi((int)(tcx + tr + 1), 0, tbx);
    *miny = clampi((int)(tcy - tr), 0, tby);
    *maxy = clampi((int)(tcy + tr + 1), 0, tby);
}

/* row-major rotation from a (w,x,y,z) quaternion; helpers.cuh:149-168 */
static void quat_to_R(const float *q, float R[9]) {
    float w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - w * z); R[2] = 2.f * (x * z + w * y);
    R[3] = 2.f * (x * y + w * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - w * x);
    R[6] = 2.f * (x * z - w * y); R[7] = 2.f * (y * z + w * x); R[8] = 1.f - 2.f * (x * x + y * y);
}

/* helpers.cuh:42-65 */
static int cov2d_bounds(const float cov[3], float conic[3], float *radius) {
    float det = cov[0] * cov[2] - cov[1] * cov[1];
    if (det == 0.f) return 0;
    float inv = 1.f / det;
    conic[0] = cov[2] * inv; conic[1] = -cov[1] * inv; conic[2] = cov[0] * inv;
    float b = 0.5f * (cov[0] + cov[2]);
    float d = sqrtf(fmaxf(0.1f, b * b - det));
    float v1 = b + d, v2 = b - d;
    *radius = ceilf(3.f * sqrtf(fmaxf(v1, v2)));
    return 1;
}

void orc_cov2d_bounds(int n, const float *cov2d, float *conics, float *radii) {
    for (int i = 0; i < n; ++i) {
        float c[3] = {0, 0, 0}, r = 0.f;
        /* bindings.cu:30 ignores the bool: outputs keep whatever the helper wrote */
        cov2d_bounds(cov2d + 3 * i, c, &r);
        conics[3 * i] = c[0]; conics[3 * i + 1] = c[1]; conics[3 * i + 2] = c[2];
        radii[i] = r;
    }
}

/* ------------------------------------------------------- projection forward */

void orc_project_forward(int n, const float *means, const float *scales, float glob_scale,
                         const float *quats, const float *lin_vel, const float *ang_vel,
                         float rs_time, float exposure, const float *vm, float fx, float fy,
                         float cx, float cy, int H, int W, int bw, float clip,
                         float *cov3d, float *xys, float *depths, float *pix_vels, int *radii,
                         float *conics, float *comp, int *tiles_hit) {
    int tbx = (W + bw - 1) / bw, tby = (H + bw - 1) / bw;
    float tan_fovx = (float)(0.5 * (double)W / (double)fx), tan_fovy = (float)(0.5 * (double)H / (double)fy); /* :63-64 */
    for (int i = 0; i < n; ++i) {
        radii[i] = 0; tiles_hit[i] = 0; /* forward.cu:42-43 (others stay at the zeros alloc) */
        const float *p = means + 3 * i;
        float pv[3];
        for (int r = 0; r < 3; ++r) /* helpers.cuh:107-114 */
            pv[r] = vm[4 * r] * p[0] + vm[4 * r + 1] * p[1] + vm[4 * r + 2] * p[2] + vm[4 * r + 3];
        if (pv[2] <= clip) continue; /* forward.cu:49 */

        /* cov3d = (R S)(R S)^T, upper triangle; forward.cu:537-557 */
        float R[9]; quat_to_R(quats + 4 * i, R);
        float M[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) M[3 * r + c] = R[3 * r + c] * (glob_scale * scales[3 * i + c]);
        float V[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                V[3 * r + c] = M[3 * r] * M[3 * c] + M[3 * r + 1] * M[3 * c + 1] + M[3 * r + 2] * M[3 * c + 2];
        float *c3 = cov3d + 6 * i;
        c3[0] = V[0]; c3[1] = V[1]; c3[2] = V[2]; c3[3] = V[4]; c3[4] = V[5]; c3[5] = V[8];

        /* EWA; forward.cu:459-534 */
        float t[3] = {pv[0], pv[1], pv[2]};
        float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
        t[0] = t[2] * fminf(limx, fmaxf(-limx, t[0] / t[2]));
        t[1] = t[2] * fminf(limy, fmaxf(-limy, t[1] / t[2]));
        float rz = 1.f / t[2], rz2 = rz * rz;
        float J[6] = {fx * rz, 0.f, -fx * t[0] * rz2, 0.f, fy * rz, -fy * t[1] * rz2};
        float T[6];
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < 3; ++c)
                T[3 * r + c] = J[3 * r] * vm[c] + J[3 * r + 1] * vm[4 + c] + J[3 * r + 2] * vm[8 + c];
        float TV[6];
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < 3; ++c)
                TV[3 * r + c] = T[3 * r] * V[c] + T[3 * r + 1] * V[3 + c] + T[3 * r + 2] * V[6 + c];
        float c00 = TV[0] * T[0] + TV[1] * T[1] + TV[2] * T[2];
        float c01 = TV[0] * T[3] + TV[1] * T[4] + TV[2] * T[5];
        float c11 = TV[3] * T[3] + TV[4] * T[4] + TV[5] * T[5];
        float det_orig = c00 * c11 - c01 * c01;
        float cov2d[3] = {c00 + 0.3f, c01, c11 + 0.3f};
        float det_blur = cov2d[0] * cov2d[2] - cov2d[1] * cov2d[1];
        float compensation = sqrtf(fmaxf(0.f, det_orig / det_blur));

        float conic[3], radius;
        if (!cov2d_bounds(cov2d, conic, &radius)) continue; /* forward.cu:75-77 */
        conics[3 * i] = conic[0]; conics[3 * i + 1] = conic[1]; conics[3 * i + 2] = conic[2]; /* :79 */

        /* helpers.cuh:128-135 */
        float rw = 1.f / (pv[2] + 1e-6f);
        float ctr[2] = {pv[0] * rw * fx + cx, pv[1] * rw * fy + cy};

        float vel[2] = {0.f, 0.f};
        if (rs_time > 0 || exposure > 0) { /* forward.cu:88-91, helpers.cuh:224-253 */
            float rot[3] = {ang_vel[1] * pv[2] - ang_vel[2] * pv[1], ang_vel[2] * pv[0] - ang_vel[0] * pv[2],
                            ang_vel[0] * pv[1] - ang_vel[1] * pv[0]};
            float tv[3] = {lin_vel[0] + rot[0], lin_vel[1] + rot[1], lin_vel[2] + rot[2]};
            float z1 = 1.f / pv[2], z2 = z1 * z1;
            vel[0] = -(fx * z1 * tv[0] + (-fx * pv[0] * z2) * tv[2]);
            vel[1] = -(fy * z1 * tv[1] + (-fy * pv[1] * z2) * tv[2]);
            radius = (float)((double)radius + (double)sqrtf(vel[0] * vel[0] + vel[1] * vel[1]) * 0.5 * (double)(exposure + rs_time));
        }
        pix_vels[2 * i] = vel[0]; pix_vels[2 * i + 1] = vel[1]; /* :92 unconditional */

        int x0, y0, x1, y1;
        tile_bbox(ctr[0], ctr[1], radius, tbx, tby, bw, &x0, &y0, &x1, &y1);
        int area = (x1 - x0) * (y1 - y0);
        if (area <= 0) continue;
        tiles_hit[i] = area; depths[i] = pv[2]; radii[i] = (int)radius;
        xys[2 * i] = ctr[0]; xys[2 * i + 1] = ctr[1]; comp[i] = compensation;
    }
}

/* ------------------------------------------------------ projection backward */

void orc_project_backward(int n, const float *means, const float *scales, float glob_scale,
                          const float *quats, const float *lin_vel, const float *ang_vel,
                          float rs_time, float exposure, const float *vm, float fx, float fy,
                          const float *cov3d, const int *radii, const float *conics,
                          const float *comp, const float *v_xy, const float *v_depth,
                          const float *v_pix_vel, const float *v_conic, const float *v_comp,
                          float *v_cov2d, float *v_cov3d, float *v_mean, float *v_scale, float *v_quat) {
    for (int i = 0; i < n; ++i) {
        if (radii[i] <= 0) continue; /* backward.cu:400 (outputs stay zero) */
        const float *p = means + 3 * i;
        float pv[3];
        for (int r = 0; r < 3; ++r)
            pv[r] = vm[4 * r] * p[0] + vm[4 * r + 1] * p[1] + vm[4 * r + 2] * p[2] + vm[4 * r + 3];

        float vpv[3] = {0, 0, 0};
        if (rs_time > 0 || exposure > 0) { /* helpers.cuh:255-326 */
            float g[2] = {v_pix_vel[2 * i], v_pix_vel[2 * i + 1]};
            float rot[3] = {ang_vel[1] * pv[2] - ang_vel[2] * pv[1], ang_vel[2] * pv[0] - ang_vel[0] * pv[2],
                            ang_vel[0] * pv[1] - ang_vel[1] * pv[0]};
            float tv[3] = {lin_vel[0] + rot[0], lin_vel[1] + rot[1], lin_vel[2] + rot[2]};
            float z1 = 1.f / pv[2], z2 = z1 * z1, z3 = z2 * z1;
            /* d(pix_vel)/d(p_view) through J */
            vpv[0] -= g[0] * (-fx * z2 * tv[2]);
            vpv[1] -= g[1] * (-fy * z2 * tv[2]);
            vpv[2] -= g[0] * (-fx * z2 * tv[0] + 2.f * fx * pv[0] * z3 * tv[2]) +
                      g[1] * (-fy * z2 * tv[1] + 2.f * fy * pv[1] * z3 * tv[2]);
            /* through total_vel = lin + ang x p */
            float vt[3] = {-(fx * z1 * g[0]), -(fy * z1 * g[1]), -((-fx * pv[0] * z2) * g[0] + (-fy * pv[1] * z2) * g[1])};
            float cr[3] = {ang_vel[1] * vt[2] - ang_vel[2] * vt[1], ang_vel[2] * vt[0] - ang_vel[0] * vt[2],
                           ang_vel[0] * vt[1] - ang_vel[1] * vt[0]};
            vpv[0] -= cr[0]; vpv[1] -= cr[1]; vpv[2] -= cr[2];
        }
        /* helpers.cuh:138-147 */
        float rw = 1.f / (pv[2] + 1e-6f);
        float gx = fx * v_xy[2 * i], gy = fy * v_xy[2 * i + 1];
        float vview[3] = {gx * rw + vpv[0], gy * rw + vpv[1],
                          -(gx * pv[0] + gy * pv[1]) * rw * rw + vpv[2] + v_depth[i]};
        float vm3[3]; /* R^T v ; helpers.cuh:97-104 */
        for (int c = 0; c < 3; ++c) vm3[c] = vm[c] * vview[0] + vm[4 + c] * vview[1] + vm[8 + c] * vview[2];

        /* conic -> cov2d ; helpers.cuh:68-79 */
        const float *cn = conics + 3 * i;
        float X[4] = {cn[0], cn[1], cn[1], cn[2]};
        float G[4] = {v_conic[3 * i], v_conic[3 * i + 1] / 2.f, v_conic[3 * i + 1] / 2.f, v_conic[3 * i + 2]};
        float XG[4] = {X[0] * G[0] + X[1] * G[2], X[0] * G[1] + X[1] * G[3], X[2] * G[0] + X[3] * G[2], X[2] * G[1] + X[3] * G[3]};
        float S[4] = {-(XG[0] * X[0] + XG[1] * X[2]), -(XG[0] * X[1] + XG[1] * X[3]),
                      -(XG[2] * X[0] + XG[3] * X[2]), -(XG[2] * X[1] + XG[3] * X[3])};
        float vc2[3] = {S[0], S[1] + S[2], S[3]};
        /* compensation -> cov2d ; helpers.cuh:81-94 */
        {
            float inv_det = cn[0] * cn[2] - cn[1] * cn[1];
            float om = 1.f - comp[i] * comp[i];
            float vsq = (float)((double)v_comp[i] * 0.5 / ((double)comp[i] + 1e-6));
            vc2[0] += vsq * (om * cn[0] - 0.3f * inv_det);
            vc2[1] += 2.f * vsq * (om * cn[1]);
            vc2[2] += vsq * (om * cn[2] - 0.3f * inv_det);
        }
        v_cov2d[3 * i] = vc2[0]; v_cov2d[3 * i + 1] = vc2[1]; v_cov2d[3 * i + 2] = vc2[2];

        /* EWA vjp ; backward.cu:454-532  (uses the UNclamped t) */
        float rz = 1.f / pv[2], rz2 = rz * rz, rz3 = rz2 * rz;
        float J[6] = {fx * rz, 0.f, -fx * pv[0] * rz2, 0.f, fy * rz, -fy * pv[1] * rz2};
        float T[6];
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < 3; ++c)
                T[3 * r + c] = J[3 * r] * vm[c] + J[3 * r + 1] * vm[4 + c] + J[3 * r + 2] * vm[8 + c];
        const float *c3 = cov3d + 6 * i;
        float V[9] = {c3[0], c3[1], c3[2], c3[1], c3[3], c3[4], c3[2], c3[4], c3[5]};
        float vC[4] = {vc2[0], 0.5f * vc2[1], 0.5f * vc2[1], vc2[2]};
        float vCT[6]; /* v_cov (2x2) * T (2x3) */
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < 3; ++c) vCT[3 * r + c] = vC[2 * r] * T[c] + vC[2 * r + 1] * T[3 + c];
        float vV[9]; /* T^T * vCT */
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) vV[3 * r + c] = T[r] * vCT[c] + T[3 + r] * vCT[3 + c];
        float *o3 = v_cov3d + 6 * i;
        o3[0] = vV[0]; o3[1] = vV[1] + vV[3]; o3[2] = vV[2] + vV[6];
        o3[3] = vV[4]; o3[4] = vV[5] + vV[7]; o3[5] = vV[8];
        float vT[6]; /* 2 * vC T V  (vC symmetric) */
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < 3; ++c)
                vT[3 * r + c] = 2.f * (vCT[3 * r] * V[c] + vCT[3 * r + 1] * V[3 + c] + vCT[3 * r + 2] * V[6 + c]);
        float vJ[6]; /* vT * W^T */
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < 3; ++c)
                vJ[3 * r + c] = vT[3 * r] * vm[4 * c] + vT[3 * r + 1] * vm[4 * c + 1] + vT[3 * r + 2] * vm[4 * c + 2];
        float vt[3] = {-fx * rz2 * vJ[2], -fy * rz2 * vJ[5],
                       -fx * rz2 * vJ[0] + 2.f * fx * pv[0] * rz3 * vJ[2] - fy * rz2 * vJ[4] + 2.f * fy * pv[1] * rz3 * vJ[5]};
        for (int c = 0; c < 3; ++c) vm3[c] += vm[c] * vt[0] + vm[4 + c] * vt[1] + vm[8 + c] * vt[2];
        v_mean[3 * i] = vm3[0]; v_mean[3 * i + 1] = vm3[1]; v_mean[3 * i + 2] = vm3[2];

        /* cov3d -> scale, quat ; backward.cu:536-572 */
        float R[9]; quat_to_R(quats + 4 * i, R);
        float s[3] = {glob_scale * scales[3 * i], glob_scale * scales[3 * i + 1], glob_scale * scales[3 * i + 2]};
        float M[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) M[3 * r + c] = R[3 * r + c] * s[c];
        float Vs[9] = {o3[0], 0.5f * o3[1], 0.5f * o3[2], 0.5f * o3[1], o3[3], 0.5f * o3[4], 0.5f * o3[2], 0.5f * o3[4], o3[5]};
        float vM[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                vM[3 * r + c] = 2.f * (Vs[3 * r] * M[c] + Vs[3 * r + 1] * M[3 + c] + Vs[3 * r + 2] * M[6 + c]);
        for (int c = 0; c < 3; ++c)
            v_scale[3 * i + c] = (R[c] * vM[c] + R[3 + c] * vM[3 + c] + R[6 + c] * vM[6 + c]) * glob_scale;
        float g[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) g[3 * r + c] = vM[3 * r + c] * s[c];
        float w = quats[4 * i], x = quats[4 * i + 1], y = quats[4 * i + 2], z = quats[4 * i + 3];
        v_quat[4 * i] = 2.f * (x * (g[7] - g[5]) + y * (g[2] - g[6]) + z * (g[3] - g[1]));
        v_quat[4 * i + 1] = 2.f * (-2.f * x * (g[4] + g[8]) + y * (g[3] + g[1]) + z * (g[6] + g[2]) + w * (g[7] - g[5]));
        v_quat[4 * i + 2] = 2.f * (x * (g[3] + g[1]) - 2.f * y * (g[0] + g[8]) + z * (g[7] + g[5]) + w * (g[2] - g[6]));
        v_quat[4 * i + 3] = 2.f * (x * (g[6] + g[2]) + y * (g[7] + g[5]) - 2.f * z * (g[0] + g[4]) + w * (g[3] - g[1]));
    }
}

/* ----------------------------------------------------- spherical harmonics */

static int sh_bases(int degree) { /* sh.cuh:42-52 */
    return degree == 0 ? 1 : degree == 1 ? 4 : degree == 2 ? 9 : degree == 3 ? 16 : 25;
}

/* basis values B_k(dir) for k < (deg+1)^2; method 0 = "poly" (sh.cuh:268-340), 1 = "fast" (sh.cuh:54-156) */
static void sh_basis(int method, int deg, const float *d, float *B) {
    if (method == 1) {
        B[0] = 0.2820947917738781f;
        if (deg < 1) return;
        float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        float x = d[0] / nrm, y = d[1] / nrm, z = d[2] / nrm;
        float a0 = 0.48860251190292f;
        B[1] = -a0 * y; B[2] = a0 * z; B[3] = -a0 * x;
        if (deg < 2) return;
        float z2 = z * z;
        float b0 = -1.092548430592079f * z, a1 = 0.5462742152960395f;
        float c1 = x * x - y * y, s1 = 2.f * x * y;
        B[6] = 0.9461746957575601f * z2 - 0.3153915652525201f;
        B[7] = b0 * x; B[5] = b0 * y; B[8] = a1 * c1; B[4] = a1 * s1;
        if (deg < 3) return;
        float c0 = -2.285228997322329f * z2 + 0.4570457994644658f;
        float b1 = 1.445305721320277f * z, a2 = -0.5900435899266435f;
        float c2 = x * c1 - y * s1, s2 = x * s1 + y * c1;
        B[12] = z * (1.865881662950577f * z2 - 1.119528997770346f);
        B[13] = c0 * x; B[11] = c0 * y; B[14] = b1 * c1; B[10] = b1 * s1; B[15] = a2 * c2; B[9] = a2 * s2;
        if (deg < 4) return;
        float d0 = z * (-4.683325804901025f * z2 + 2.007139630671868f);
        float cc = 3.31161143515146f * z2 - 0.47308734787878f;
        float b2 = -1.770130769779931f * z, a3 = 0.6258357354491763f;
        float c3 = x * c2 - y * s2, s3 = x * s2 + y * c2;
        B[20] = 1.984313483298443f * z * B[12] - 1.006230589874905f * B[6];
        B[21] = d0 * x; B[19] = d0 * y; B[22] = cc * c1; B[18] = cc * s1;
        B[23] = b2 * c2; B[17] = b2 * s2; B[24] = a3 * c3; B[16] = a3 * s3;
    } else {
        const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
        const float C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f, 0.5462742152960396f};
        const float C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                             -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};
        const float C4[9] = {2.5033429417967046f, -1.7701307697799304f, 0.9461746957575601f, -0.6690465435572892f,
                             0.10578554691520431f, -0.6690465435572892f, 0.47308734787878004f, -1.7701307697799304f,
                             0.6258357354491761f};
        B[0] = C0;
        if (deg < 1) return;
        float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        float x = d[0] / nrm, y = d[1] / nrm, z = d[2] / nrm;
        float xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
        B[1] = -C1 * y; B[2] = C1 * z; B[3] = -C1 * x;
        if (deg < 2) return;
        B[4] = C2[0] * xy; B[5] = C2[1] * yz; B[6] = C2[2] * (2.f * zz - xx - yy); B[7] = C2[3] * xz; B[8] = C2[4] * (xx - yy);
        if (deg < 3) return;
        B[9] = C3[0] * y * (3.f * xx - yy); B[10] = C3[1] * xy * z; B[11] = C3[2] * y * (4.f * zz - xx - yy);
        B[12] = C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy); B[13] = C3[4] * x * (4.f * zz - xx - yy);
        B[14] = C3[5] * z * (xx - yy); B[15] = C3[6] * x * (xx - 3.f * yy);
        if (deg < 4) return;
        B[16] = C4[0] * xy * (xx - yy); B[17] = C4[1] * yz * (3.f * xx - yy); B[18] = C4[2] * xy * (7.f * zz - 1.f);
        B[19] = C4[3] * yz * (7.f * zz - 3.f); B[20] = C4[4] * (zz * (35.f * zz - 30.f) + 3.f);
        B[21] = C4[5] * xz * (7.f * zz - 3.f); B[22] = C4[6] * (xx - yy) * (7.f * zz - 1.f);
        B[23] = C4[7] * xz * (xx - 3.f * yy); B[24] = C4[8] * (xx * (xx - 3.f * yy) - yy * (3.f * xx - yy));
    }
}

/* sh.cuh:434-465 ; coeffs (N, K(degree), 3) -> colors (N,3) using bases < (degrees_to_use+1)^2 */
void orc_sh_forward(int method, int n, int degree, int deg_use, const float *dirs, const float *coeffs, float *colors) {
    int K = sh_bases(degree), Ku = sh_bases(deg_use);
    for (int i = 0; i < n; ++i) {
        float B[25];
        sh_basis(method, deg_use, dirs + 3 * i, B);
        for (int c = 0; c < 3; ++c) {
            float acc = 0.f;
            for (int k = 0; k < Ku; ++k) acc += B[k] * coeffs[((size_t)i * K + k) * 3 + c];
            colors[3 * i + c] = acc;
        }
    }
}

/* sh.cuh:467-498 ; v_coeffs zero above the used bases (bindings.cu:123-124) */
void orc_sh_backward(int method, int n, int degree, int deg_use, const float *dirs, const float *v_colors, float *v_coeffs) {
    int K = sh_bases(degree), Ku = sh_bases(deg_use);
    memset(v_coeffs, 0, sizeof(float) * (size_t)n * K * 3);
    for (int i = 0; i < n; ++i) {
        float B[25];
        sh_basis(method, deg_use, dirs + 3 * i, B);
        for (int k = 0; k < Ku; ++k)
            for (int c = 0; c < 3; ++c) v_coeffs[((size_t)i * K + k) * 3 + c] = B[k] * v_colors[3 * i + c];
    }
}

/* ----------------------------------------------------------------- binning */

/* forward.cu:116-153 ; isect_ids/gaussian_ids must be zero-initialised by the
 * caller (bindings.cu:381-384): unwritten (phantom) slots keep key 0 / id 0. */
void orc_map_intersects(int n, const float *xys, const float *depths, const int *radii,
                        const int *cum_tiles_hit, int tbx, int tby, int bw,
                        int64_t *isect_ids, int32_t *gaussian_ids) {
    for (int i = 0; i < n; ++i) {
        if (radii[i] <= 0) continue;
        int x0, y0, x1, y1;
        tile_bbox(xys[2 * i], xys[2 * i + 1], (float)radii[i], tbx, tby, bw, &x0, &y0, &x1, &y1);
        int cur = i == 0 ? 0 : cum_tiles_hit[i - 1];
        int32_t bits; memcpy(&bits, depths + i, 4);
        int64_t depth_id = (int64_t)bits;
        for (int ty = y0; ty < y1; ++ty)
            for (int tx = x0; tx < x1; ++tx) {
                int64_t tile = (int64_t)ty * tbx + tx;
                isect_ids[cur] = (tile << 32) | depth_id;
                gaussian_ids[cur] = i;
                ++cur;
            }
    }
}

/* utils.py:179-180 : ascending sort by key, values gathered; ties keep input order (stable LSD radix sort,
 * 8-bit digits over the non-negative 64-bit keys) */
void orc_sort_intersects(int m, const int64_t *keys, const int32_t *vals, int64_t *keys_out, int32_t *vals_out) {
    if (m <= 0) return;
    uint64_t *ka = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)m), *kb = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)m);
    int32_t *va = (int32_t *)malloc(sizeof(int32_t) * (size_t)m), *vb = (int32_t *)malloc(sizeof(int32_t) * (size_t)m);
    uint64_t all = 0;
    for (int i = 0; i < m; ++i) { ka[i] = (uint64_t)keys[i]; va[i] = vals[i]; all |= ka[i]; }
    for (int shift = 0; shift < 64; shift += 8) {
        if (((all >> shift) & 0xff) == 0 && (all >> shift) == 0) break;
        size_t cnt[257];
        memset(cnt, 0, sizeof(cnt));
        for (int i = 0; i < m; ++i) cnt[((ka[i] >> shift) & 0xff) + 1]++;
        for (int b = 0; b < 256; ++b) cnt[b + 1] += cnt[b];
        for (int i = 0; i < m; ++i) {
            size_t dst = cnt[(ka[i] >> shift) & 0xff]++;
            kb[dst] = ka[i]; vb[dst] = va[i];
        }
        uint64_t *tk = ka; ka = kb; kb = tk;
        int32_t *tv = va; va = vb; vb = tv;
    }
    for (int i = 0; i < m; ++i) { keys_out[i] = (int64_t)ka[i]; vals_out[i] = va[i]; }
    free(ka); free(kb); free(va); free(vb);
}

/* forward.cu:158-180 ; tile_bins (tiles,2) zero-initialised by the caller */
void orc_tile_bin_edges(int m, const int64_t *sorted, int32_t *tile_bins) {
    for (int i = 0; i < m; ++i) {
        int32_t cur = (int32_t)(sorted[i] >> 32);
        if (i == 0) tile_bins[2 * cur] = 0;
        if (i == m - 1) tile_bins[2 * cur + 1] = m;
        if (i == 0) continue;
        int32_t prev = (int32_t)(sorted[i - 1] >> 32);
        if (prev != cur) { tile_bins[2 * prev + 1] = i; tile_bins[2 * cur] = i; }
    }
}

/* ----------------------------------------------------------- blend forward */

/* forward.cu:306-456.  out_img (H,W,3), final_Ts/final_idx (H,W,S). */
static void rasterize_forward_rows(int row0, int row1, int H, int W, int bw, int S, const int32_t *ids_sorted,
                                   const int32_t *tile_bins, const float *xys, const float *pix_vels, float rs_time,
                                   float exposure, const float *conics, const float *colors, const float *opac,
                                   const float *bg, float *out_img, float *final_Ts, int32_t *final_idx) {
    int tbx = (W + bw - 1) / bw;
    float avg = 1.0f / (float)S;
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = row0; i < row1; ++i) {
        for (int j = 0; j < W; ++j) {
            int tile = (i / bw) * tbx + (j / bw);
            int r0 = tile_bins[2 * tile], r1 = tile_bins[2 * tile + 1];
            float px = (float)j + 0.5f, py = (float)i + 0.5f;
            float roll = (float)((double)rs_time * ((double)(py / (float)H) - 0.5)); /* :360 (double literal) */
            float acc[3] = {0, 0, 0}, meanT = 0.f;
            size_t pix = (size_t)i * W + j;
            for (int s = 0; s < S && s < ORC_MAX_BLUR; ++s) {
                float rel = ((S > 1) ? ((float)s / (float)(S - 1) - 0.5f) * exposure : 0.0f) + roll; /* :363 */
                float T = 1.f; int last = 0;
                for (int k = r0; k < r1; ++k) {
                    int g = ids_sorted[k];
                    float dx = xys[2 * g] + rel * pix_vels[2 * g] - px;
                    float dy = xys[2 * g + 1] + rel * pix_vels[2 * g + 1] - py;
                    const float *cn = conics + 3 * g;
                    float sigma = 0.5f * (cn[0] * dx * dx + cn[2] * dy * dy) + cn[1] * dx * dy;
                    if (sigma > 80.f) continue; /* exp(-80) * opac < 1/255 for any finite opac <= 1e30: same skip as below,
                                                   without libm's slow underflow path (CPU-baseline speed only) */
                    float alpha = fminf(0.999f, opac[g] * expf(-sigma));
                    if (sigma < 0.f || alpha < 1.f / 255.f) continue;
                    float nT = T * (1.f - alpha);
                    if (nT <= 1e-4f) break;
                    float vis = alpha * T * avg;
                    acc[0] += colors[3 * g] * vis; acc[1] += colors[3 * g + 1] * vis; acc[2] += colors[3 * g + 2] * vis;
                    T = nT; last = k;
                }
                meanT += T * avg;
                final_Ts[pix * S + s] = T; final_idx[pix * S + s] = last;
            }
            out_img[3 * pix] = acc[0] + meanT * bg[0];
            out_img[3 * pix + 1] = acc[1] + meanT * bg[1];
            out_img[3 * pix + 2] = acc[2] + meanT * bg[2];
        }
    }
}

void orc_rasterize_forward(int H, int W, int bw, int S, const int32_t *ids_sorted, const int32_t *tile_bins,
                           const float *xys, const float *pix_vels, float rs_time, float exposure,
                           const float *conics, const float *colors, const float *opac, const float *bg,
                           float *out_img, float *final_Ts, int32_t *final_idx) {
    rasterize_forward_rows(0, H, H, W, bw, S, ids_sorted, tile_bins, xys, pix_vels, rs_time, exposure, conics, colors,
                           opac, bg, out_img, final_Ts, final_idx);
}

/* Same, restricted to image rows [row0, row1): used by bench.py to time a bounded sample of a large image. */
void orc_rasterize_forward_rows(int row0, int row1, int H, int W, int bw, int S, const int32_t *ids_sorted,
                                const int32_t *tile_bins, const float *xys, const float *pix_vels, float rs_time,
                                float exposure, const float *conics, const float *colors, const float *opac,
                                const float *bg, float *out_img, float *final_Ts, int32_t *final_idx) {
    rasterize_forward_rows(row0 < 0 ? 0 : row0, row1 > H ? H : row1, H, W, bw, S, ids_sorted, tile_bins, xys, pix_vels,
                           rs_time, exposure, conics, colors, opac, bg, out_img, final_Ts, final_idx);
}

/* ---------------------------------------------------------- blend backward */

static inline void atomic_addd(double *p, double v) {
#pragma omp atomic
    *p += v;
}

/* backward.cu:143-369.  Outputs (float, written from double accumulators):
 * v_xy (N,2) v_xy_abs (N,2) v_pix_vel (N,2) v_conic (N,3) v_rgb (N,3) v_opac (N) */
void orc_rasterize_backward_rows(int row0, int row1, int n, int H, int W, int bw, int S, const int32_t *ids_sorted,
                                 const int32_t *tile_bins, const float *xys, const float *pix_vels, float rs_time,
                                 float exposure, const float *conics, const float *rgbs, const float *opac,
                                 const float *bg, const float *final_Ts, const int32_t *final_idx, const float *v_out,
                                 const float *v_out_alpha, float *v_xy, float *v_xy_abs, float *v_pix_vel,
                                 float *v_conic, float *v_rgb, float *v_opac) {
    int tbx = (W + bw - 1) / bw;
    float avg = 1.0f / (float)S;
    double *A = (double *)calloc((size_t)n * 13, sizeof(double));
    if (row0 < 0) row0 = 0;
    if (row1 > H) row1 = H;
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = row0; i < row1; ++i) {
        for (int j = 0; j < W; ++j) {
            int tile = (i / bw) * tbx + (j / bw);
            int r0 = tile_bins[2 * tile], r1 = tile_bins[2 * tile + 1];
            float px = (float)j + 0.5f, py = (float)i + 0.5f;
            float roll = (float)((double)rs_time * ((double)(py / (float)H) - 0.5));
            size_t pix = (size_t)i * W + j;
            const float *vo = v_out + 3 * pix;
            float voa = v_out_alpha[pix];
            for (int s = 0; s < S && s < ORC_MAX_BLUR; ++s) {
                float rel = ((S > 1) ? ((float)s / (float)(S - 1) - 0.5f) * exposure : 0.0f) + roll;
                float T_final = final_Ts[pix * S + s], T = T_final, Tfm = T_final * avg;
                float buf[3] = {0, 0, 0};
                int bin_final = final_idx[pix * S + s];
                if (bin_final > r1 - 1) bin_final = r1 - 1; /* batches only cover [r0, r1) (:224-235) */
                for (int k = bin_final; k >= r0; --k) { /* :250-254: entries with index <= bin_final */
                    int g = ids_sorted[k];
                    const float *cn = conics + 3 * g;
                    float dx = xys[2 * g] + rel * pix_vels[2 * g] - px;
                    float dy = xys[2 * g + 1] + rel * pix_vels[2 * g + 1] - py;
                    float sigma = 0.5f * (cn[0] * dx * dx + cn[2] * dy * dy) + cn[1] * dx * dy;
                    if (sigma > 80.f) continue; /* same outcome as the alpha < 1/255 skip below */
                    float vis = expf(-sigma);
                    float alpha = fminf(0.99f, opac[g] * vis); /* :275 -- 0.99, not 0.999 */
                    if (sigma < 0.f || alpha < 1.f / 255.f) continue;
                    float ra = 1.f / (1.f - alpha);
                    T *= ra;
                    float Tm = T * avg, fac = alpha * Tm;
                    const float *rgb = rgbs + 3 * g;
                    float va = 0.f;
                    va += (rgb[0] * Tm - buf[0] * ra) * vo[0];
                    va += (rgb[1] * Tm - buf[1] * ra) * vo[1];
                    va += (rgb[2] * Tm - buf[2] * ra) * vo[2];
                    va += Tfm * ra * voa;
                    va += -Tfm * ra * bg[0] * vo[0];
                    va += -Tfm * ra * bg[1] * vo[1];
                    va += -Tfm * ra * bg[2] * vo[2];
                    buf[0] += rgb[0] * fac; buf[1] += rgb[1] * fac; buf[2] += rgb[2] * fac;
                    float vs = -opac[g] * vis * va;
                    float gx = vs * (cn[0] * dx + cn[1] * dy), gy = vs * (cn[1] * dx + cn[2] * dy);
                    double *a = A + (size_t)g * 13;
                    atomic_addd(a + 0, gx); atomic_addd(a + 1, gy);
                    atomic_addd(a + 2, fabsf(gx)); atomic_addd(a + 3, fabsf(gy));
                    atomic_addd(a + 4, gx * rel); atomic_addd(a + 5, gy * rel);
                    atomic_addd(a + 6, 0.5f * vs * dx * dx); atomic_addd(a + 7, vs * dx * dy); atomic_addd(a + 8, 0.5f * vs * dy * dy);
                    atomic_addd(a + 9, fac * vo[0]); atomic_addd(a + 10, fac * vo[1]); atomic_addd(a + 11, fac * vo[2]);
                    atomic_addd(a + 12, vis * va);
                }
            }
        }
    }
    for (int g = 0; g < n; ++g) {
        const double *a = A + (size_t)g * 13;
        v_xy[2 * g] = (float)a[0]; v_xy[2 * g + 1] = (float)a[1];
        v_xy_abs[2 * g] = (float)a[2]; v_xy_abs[2 * g + 1] = (float)a[3];
        v_pix_vel[2 * g] = (float)a[4]; v_pix_vel[2 * g + 1] = (float)a[5];
        v_conic[3 * g] = (float)a[6]; v_conic[3 * g + 1] = (float)a[7]; v_conic[3 * g + 2] = (float)a[8];
        v_rgb[3 * g] = (float)a[9]; v_rgb[3 * g + 1] = (float)a[10]; v_rgb[3 * g + 2] = (float)a[11];
        v_opac[g] = (float)a[12];
    }
    free(A);
}

void orc_rasterize_backward(int n, int H, int W, int bw, int S, const int32_t *ids_sorted, const int32_t *tile_bins,
                            const float *xys, const float *pix_vels, float rs_time, float exposure,
                            const float *conics, const float *rgbs, const float *opac, const float *bg,
                            const float *final_Ts, const int32_t *final_idx, const float *v_out,
                            const float *v_out_alpha, float *v_xy, float *v_xy_abs, float *v_pix_vel,
                            float *v_conic, float *v_rgb, float *v_opac) {
    orc_rasterize_backward_rows(0, H, n, H, W, bw, S, ids_sorted, tile_bins, xys, pix_vels, rs_time, exposure, conics,
                                rgbs, opac, bg, final_Ts, final_idx, v_out, v_out_alpha, v_xy, v_xy_abs, v_pix_vel,
                                v_conic, v_rgb, v_opac);
}

/* ------------------------------------------------------ N-channel blend (a9) */

static inline float h2f(_Float16 h) { return (float)h; }
static inline _Float16 f2h(float f) { return (_Float16)f; }

/* forward.cu:185-304 ; fp16 per-pixel accumulators, no blur */
void orc_nd_rasterize_forward(int H, int W, int bw, int C, const int32_t *ids_sorted, const int32_t *tile_bins,
                              const float *xys, const float *conics, const float *colors, const float *opac,
                              const float *bg, float *out_img, float *final_Ts, int32_t *final_idx) {
    int tbx = (W + bw - 1) / bw;
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < H; ++i) {
        _Float16 *acc = (_Float16 *)malloc(sizeof(_Float16) * (size_t)C);
        for (int j = 0; j < W; ++j) {
            int tile = (i / bw) * tbx + (j / bw);
            int r0 = tile_bins[2 * tile], r1 = tile_bins[2 * tile + 1];
            float px = (float)j + 0.5f, py = (float)i + 0.5f;
            for (int c = 0; c < C; ++c) acc[c] = f2h(0.f);
            float T = 1.f; int last = 0;
            for (int k = r0; k < r1; ++k) {
                int g = ids_sorted[k];
                const float *cn = conics + 3 * g;
                float dx = xys[2 * g] - px, dy = xys[2 * g + 1] - py;
                float sigma = 0.5f * (cn[0] * dx * dx + cn[2] * dy * dy) + cn[1] * dx * dy;
                float alpha = fminf(0.999f, opac[g] * expf(-sigma));
                if (sigma < 0.f || alpha < 1.f / 255.f) continue;
                float nT = T * (1.f - alpha);
                if (nT <= 1e-4f) break;
                float vis = alpha * T;
                for (int c = 0; c < C; ++c) acc[c] = f2h(h2f(acc[c]) + h2f(f2h(colors[(size_t)C * g + c] * vis)));
                T = nT; last = k;
            }
            size_t pix = (size_t)i * W + j;
            final_Ts[pix] = T; final_idx[pix] = last;
            for (int c = 0; c < C; ++c) out_img[pix * C + c] = h2f(acc[c]) + T * bg[c];
        }
        free(acc);
    }
}

/* backward.cu:22-141 ; walks final_idx-1 .. range.x (note: strictly below bin_final) */
void orc_nd_rasterize_backward(int n, int H, int W, int bw, int C, const int32_t *ids_sorted, const int32_t *tile_bins,
                               const float *xys, const float *conics, const float *rgbs, const float *opac,
                               const float *bg, const float *final_Ts, const int32_t *final_idx,
                               const float *v_out, const float *v_out_alpha, float *v_xy, float *v_xy_abs,
                               float *v_conic, float *v_rgb, float *v_opac) {
    int tbx = (W + bw - 1) / bw;
    double *A = (double *)calloc((size_t)n * 8, sizeof(double));
    double *Argb = (double *)calloc((size_t)n * C, sizeof(double));
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < H; ++i) {
        _Float16 *Sb = (_Float16 *)malloc(sizeof(_Float16) * (size_t)C);
        for (int j = 0; j < W; ++j) {
            int tile = (i / bw) * tbx + (j / bw);
            int r0 = tile_bins[2 * tile];
            float px = (float)j + 0.5f, py = (float)i + 0.5f;
            size_t pix = (size_t)i * W + j;
            const float *vo = v_out + pix * C;
            float voa = v_out_alpha[pix];
            float T_final = final_Ts[pix], T = T_final;
            for (int c = 0; c < C; ++c) Sb[c] = f2h(0.f);
            int bin_final = final_idx[pix];
            for (int k = bin_final - 1; k >= r0; --k) {
                int g = ids_sorted[k];
                const float *cn = conics + 3 * g;
                float dx = xys[2 * g] - px, dy = xys[2 * g + 1] - py;
                float sigma = 0.5f * (cn[0] * dx * dx + cn[2] * dy * dy) + cn[1] * dx * dy;
                float vis = expf(-sigma);
                float alpha = fminf(0.99f, opac[g] * vis);
                if (sigma < 0.f || alpha < 1.f / 255.f) continue;
                float ra = 1.f / (1.f - alpha);
                T *= ra;
                float fac = alpha * T, va = 0.f;
                for (int c = 0; c < C; ++c) {
                    atomic_addd(Argb + (size_t)g * C + c, fac * vo[c]);
                    va += (rgbs[(size_t)C * g + c] * T - h2f(Sb[c]) * ra) * vo[c];
                    va += -T_final * ra * bg[c] * vo[c];
                    Sb[c] = f2h(h2f(Sb[c]) + h2f(f2h(rgbs[(size_t)C * g + c] * fac)));
                }
                va += T_final * ra * voa;
                float vs = -opac[g] * vis * va;
                float gx = vs * (cn[0] * dx + cn[1] * dy), gy = vs * (cn[1] * dx + cn[2] * dy);
                double *a = A + (size_t)g * 8;
                atomic_addd(a + 0, gx); atomic_addd(a + 1, gy);
                atomic_addd(a + 2, fabsf(gx)); atomic_addd(a + 3, fabsf(gy));
                atomic_addd(a + 4, 0.5f * vs * dx * dx); atomic_addd(a + 5, vs * dx * dy); atomic_addd(a + 6, 0.5f * vs * dy * dy);
                atomic_addd(a + 7, vis * va);
            }
        }
        free(Sb);
    }
    for (int g = 0; g < n; ++g) {
        const double *a = A + (size_t)g * 8;
        v_xy[2 * g] = (float)a[0]; v_xy[2 * g + 1] = (float)a[1];
        v_xy_abs[2 * g] = (float)a[2]; v_xy_abs[2 * g + 1] = (float)a[3];
        v_conic[3 * g] = (float)a[4]; v_conic[3 * g + 1] = (float)a[5]; v_conic[3 * g + 2] = (float)a[6];
        v_opac[g] = (float)a[7];
        for (int c = 0; c < C; ++c) v_rgb[(size_t)g * C + c] = (float)Argb[(size_t)g * C + c];
    }
    free(A); free(Argb);
}

int orc_abi_version(void) { return 1; }
