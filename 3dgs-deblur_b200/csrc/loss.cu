// Fused L1 photometric loss (the L1 term of splatfacto.py:957 `torch.abs(gt - pred).mean()`):
// one pass over the rendered image computes mean |pred - target| AND the cotangent sign(pred - target) / n that the
// blend backward consumes, instead of sub / abs / mean forward plus sign / mul / div / expand in autograd
// (7 image-sized passes).  Deterministic: per-block partial sums, the last block to finish adds them in block order.
#include "common.cuh"

namespace b200 {

constexpr int L1_THREADS = 256;
constexpr int L1_MAX_BLOCKS = 148 * 4;

// GAMMA: the caller's "linear -> gamma corrected" step (splatfacto.py:879-880, `clamp(rgb, max=1) ** (1 / gamma)`) is
// applied to `pred` on the fly: loss = mean |min(pred, 1)^(1/gamma) - target| and grad = the cotangent w.r.t. the LINEAR
// image, sign * (1/gamma) * min(pred, 1)^(1/gamma - 1) below the clamp and 0 above it (torch's clamp / pow backward,
// including the infinite slope at exactly 0).  Two image-sized elementwise passes forward and three backward disappear.
template <bool GAMMA>
__global__ void __launch_bounds__(L1_THREADS) l1_loss_kernel(const float *__restrict__ pred,
                                                             const float *__restrict__ target, long long n,
                                                             float inv_n, float inv_gamma, float *__restrict__ grad,
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
    // one element: |f(x) - t| and d/dx; f = identity or the clamped power
    auto one = [&](float x, float t, float &g) {
        if (!GAMMA) {
            const float d = x - t;
            g = sgn(d);
            return fabsf(d);
        }
        const float c = fminf(x, 1.0f);
        const float y = powf(c, inv_gamma);
        const float d = y - t;
        // d y / d x = (1/gamma) c^(1/gamma - 1) where x <= 1 (torch routes the gradient of clamp(max=1) to x at x == 1), 0 above
        g = (x <= 1.0f) ? sgn(d) * (inv_gamma * powf(c, inv_gamma - 1.0f)) : 0.f;
        return fabsf(d);
    };
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 a = p4[i], b = __ldg(t4 + i);
        float4 g;
        const float e0 = one(a.x, b.x, g.x), e1 = one(a.y, b.y, g.y), e2 = one(a.z, b.z, g.z), e3 = one(a.w, b.w, g.w);
        sum += (e0 + e1) + (e2 + e3);
        if (grad) g4[i] = g;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {  // tail (n not a multiple of 4)
        const long long i = (n4 << 2) + threadIdx.x;
        float g;
        sum += one(pred[i], target[i], g);
        if (grad) grad[i] = g;
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

static int run_l1(long long numel, const float *pred, const float *target, float gamma, float *loss, float *grad, void *ws,
                  int ws_is_zeroed, void *stream);

extern "C" int b200_l1_loss(long long numel, const float *pred, const float *target, float *loss, float *grad, void *ws,
                            int ws_is_zeroed, void *stream) {
    return run_l1(numel, pred, target, 0.f, loss, grad, ws, ws_is_zeroed, stream);
}

// L1 loss of the GAMMA-CORRECTED render against the target, from the linear render (see l1_loss_kernel<true>); gamma > 0.
extern "C" int b200_l1_loss_gamma(long long numel, const float *pred_linear, const float *target, float gamma, float *loss,
                                  float *grad_linear, void *ws, int ws_is_zeroed, void *stream) {
    B200_REQUIRE(gamma > 0.f, "gamma must be positive");
    return run_l1(numel, pred_linear, target, gamma, loss, grad_linear, ws, ws_is_zeroed, stream);
}

static int run_l1(long long numel, const float *pred, const float *target, float gamma, float *loss, float *grad, void *ws,
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
    if (gamma > 0.f)
        l1_loss_kernel<true><<<blocks, L1_THREADS, 0, st>>>(pred, target, numel, 1.0f / (float)numel, 1.0f / gamma, grad, partial, ticket, loss);
    else
        l1_loss_kernel<false><<<blocks, L1_THREADS, 0, st>>>(pred, target, numel, 1.0f / (float)numel, 1.0f, grad, partial, ticket, loss);
    B200_LAUNCH_CHECK();
    return B200_OK;
}
