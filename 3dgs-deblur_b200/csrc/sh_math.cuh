// Spherical-harmonics basis shared by sh.cu and fused.cu (semantics of the reference's sh.cuh:54-340).
#pragma once
#include "common.cuh"

namespace b200 {

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

}  // namespace b200
