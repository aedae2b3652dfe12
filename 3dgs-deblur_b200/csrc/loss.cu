// Fused L1 photometric loss (the L1 term of splatfacto.py:957 `torch.abs(gt - pred).mean()`):
// one pass over the rendered image computes mean |pred - target| AND the cotangent sign(pred - target) / n that the
// blend backward consumes, instead of sub / abs / mean forward plus sign / mul / div / expand in autograd
// (7 image-sized passes).  Deterministic: per-block partial sums, the last block to finish adds them in block order.
#include "common.cuh"

namespace b200 {

constexpr int L1_THREADS = 256;
constexpr int L1_MAX_BLOCKS = 148 * 4;

__global__ void __launch_bounds__(L1_THREADS) l1_loss_kernel(const float *__restrict__ pred,
                                                             const float *__restrict__ target, long long n,
                                                             float inv_n, float *__restrict__ grad,
                                                             float *__restrict__ partial,
                                                             unsigned int *__restrict__ ticket,
                                                             float *__restrict__ loss) {
    float sum = 0.f;
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float4 *p4 = reinterpret_cast<const float4 *>(pred);
    const float4 *t4 = reinterpret_cast<const float4 *>(target);
    float4 *g4 = reinterpret_cast<float4 *>(grad);
    auto sgn = [inv_n](float d) { return d > 0.f ? inv_n : (d < 0.f ? -inv_n : 0.f); };  // torch.sign: 0 at 0
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 a = p4[i], b = __ldg(t4 + i);
        const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
        sum += (fabsf(d0) + fabsf(d1)) + (fabsf(d2) + fabsf(d3));
        if (grad) g4[i] = make_float4(sgn(d0), sgn(d1), sgn(d2), sgn(d3));
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {  // tail (n not a multiple of 4)
        const long long i = (n4 << 2) + threadIdx.x;
        const float d = pred[i] - target[i];
        sum += fabsf(d);
        if (grad) grad[i] = sgn(d);
    }
    sum = warp_sum(sum);
    __shared__ float s_part[L1_THREADS / 32];
    __shared__ bool s_last;
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float b = 0.f;
#pragma unroll
        for (int w = 0; w < L1_THREADS / 32; ++w) b += s_part[w];
        partial[blockIdx.x] = b;
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) {
        __threadfence();
        float t = 0.f;
        for (int k = threadIdx.x; k < (int)gridDim.x; k += 32) t += __ldcg(partial + k);
        t = warp_sum(t);
        if (threadIdx.x == 0) {
            *loss = t * inv_n;
            *ticket = 0u;  // ready for the next launch on this workspace
        }
    }
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200_l1_loss_ws_bytes(void) { return sizeof(float) * L1_MAX_BLOCKS + 256; }

extern "C" int b200_l1_loss(long long numel, const float *pred, const float *target, float *loss, float *grad, void *ws,
                            int ws_is_zeroed, void *stream) {
    B200_REQUIRE(numel >= 1, "numel must be >= 1");
    B200_REQUIRE(pred && target && loss && ws, "null pointer");
    B200_REQUIRE(aligned16(pred) && aligned16(target) && (!grad || aligned16(grad)) && aligned16(ws),
                 "pred / target / grad / ws must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    unsigned int *ticket = static_cast<unsigned int *>(ws);
    float *partial = reinterpret_cast<float *>(static_cast<char *>(ws) + 256);
    if (!ws_is_zeroed) B200_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), st));
    const long long n4 = numel >> 2;
    long long want = (n4 + L1_THREADS - 1) / L1_THREADS;
    const int blocks = (int)(want < 1 ? 1 : (want > L1_MAX_BLOCKS ? L1_MAX_BLOCKS : want));
    l1_loss_kernel<<<blocks, L1_THREADS, 0, st>>>(pred, target, numel, 1.0f / (float)numel, grad, partial, ticket, loss);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
