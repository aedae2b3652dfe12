// Microbenchmark: does packed fma.rn.f32x2 (SASS FFMA2) free ISSUE slots on sm_100a?
// Round 1 measured FFMA2 alone: same FLOP rate as scalar FFMA (the FMA pipe is busy 2 cycles per FFMA2), i.e. no gain for
// an FMA-pipe-bound loop.  The blend kernels are bound by total instruction issue (1 warp-instruction per cycle per SM
// sub-partition) with the FMA and ALU pipes each ~42 % busy, so the question is the MIX: per iteration
//   scalar: 16 FFMA + 8 ALU-pipe ops (PRMT)  = 24 issue slots
//   packed:  8 FFMA2 + 8 ALU-pipe ops         = 16 issue slots, FMA pipe busy 16 cycles
// If FFMA2 leaves the issue port free while the FMA pipe works on its second half, the packed mix runs in ~16 cycles
// per iteration instead of ~24.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_mix_bench ffma2_mix_bench.cu && ./ffma2_mix_bench
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 x, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x)); }
// an ALU-pipe op ptxas cannot merge across iterations (two fminf become one PRMT3): byte permute, already 3-input
__device__ __forceinline__ float alu_min(float a, float b) { float r; asm volatile("prmt.b32 %0, %1, %2, 0x3614;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) kern(float *out, float seed, int iters) {
    float a = seed + threadIdx.x, b = seed * 0.5f, c = 1.0f;
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = a - i;
    if (MODE == 0) {
        float x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = a + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(b), "f"(c));
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = alu_min(y[i], x[2 * i]);  // ALU pipe (PRMT)
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += x[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) s += y[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        u64 x[8];
        const u64 b2 = pack(b, b), c2 = pack(c, c);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = pack(a + i, a - i);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = ffma2(x[i], b2, c2);
#pragma unroll
            for (int i = 0; i < 8; ++i) { float lo, hi; unpack(x[i], lo, hi); y[i] = alu_min(y[i], lo); }
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { float lo, hi; unpack(x[i], lo, hi); s += lo + hi + y[i]; }
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}

int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, blocks = 148 * 8;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) kern<0><<<blocks, 256>>>(out, 1.0f, iters); else kern<1><<<blocks, 256>>>(out, 1.0f, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double warp_iters = (double)blocks * 256 / 32 * iters;
            // 148 SMs x 4 sub-partitions issue in parallel: cycles per iteration per sub-partition at 1.965 GHz
            const double cyc = ms * 1e-3 * 1.965e9 / (warp_iters / (148.0 * 4.0));
            if (rep) printf("%s: %.3f ms  %.1f cycles per warp-iteration per sub-partition (%d issue slots)\n",
                            mode ? "8 FFMA2 + 8 PRMT " : "16 FFMA + 8 PRMT ", ms, cyc, mode ? 16 : 24);
        }
    }
    return 0;
}
