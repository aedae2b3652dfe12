// Microbenchmark: issue throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) kern(float *out, float seed, int iters) {
    float a = seed + threadIdx.x, b = seed * 0.5f, c = 1.0f;
    if (MODE == 0) {  // 8 independent scalar FFMA chains
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = a + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], b, c);
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += x[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {  // 8 independent packed chains (16 FMAs per iteration)
        u64 x[8];
        const u64 b2 = pack(b, b), c2 = pack(c, c);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = pack(a + i, a - i);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = ffma2(x[i], b2, c2);
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
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
            double instr = (double)blocks * 256 / 32 * iters * 8;  // warp instructions
            double fma = (double)blocks * 256 * iters * 8 * (mode ? 2 : 1);
            if (rep) printf("%s: %.3f ms  %.2f warp-instr/clk/SM (at 1.965 GHz)  %.1f TFLOP/s\n", mode ? "FFMA2" : "FFMA ", ms,
                            instr / (ms * 1e-3) / 148 / 1.965e9, 2 * fma / (ms * 1e-3) / 1e12);
        }
    }
    return 0;
}
