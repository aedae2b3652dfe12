// Adam over a contiguous slice of the flat parameter buffer ("one fused multi-tensor Adam over the
// flat 59*N buffer", the largest HBM stream of the train step: 1652 B per Gaussian).  The reference steps one
// torch.optim.Adam per parameter group (nerfstudio/engine/optimizers.py:158-171; Adam, eps 1e-15, no weight decay,
// splatfacto.py:1063-1100).  Here the trainer owns flat {param, grad, exp_avg, exp_avg_sq} buffers and this kernel
// updates any [begin, end) slice of them in one pass: 16 B read-modify-write per moment, 16-byte vector accesses, the
// gradient slice is optionally cleared on the way out (replaces the separate zero_grad fill), and an optional scale
// folds a 1/world average into the same pass.  Pure streaming, HBM-bound: 28 (+4 when clearing) bytes per element.
#include "common.cuh"

namespace b200 {

struct AdamArgs {  // every derived constant is formed in double on the host, like the Python optimizer does
    float lr_over_bc1, inv_sqrt_bc2, one_minus_beta1, beta2, one_minus_beta2, eps, grad_scale;
    int zero_grad;
};

__device__ __forceinline__ void adam_one(float &p, float &g, float &m, float &v, const AdamArgs &a) {
    const float gs = g * a.grad_scale;
    m = m + a.one_minus_beta1 * (gs - m);                 // exp_avg.lerp_(grad, 1 - beta1)
    v = a.beta2 * v + a.one_minus_beta2 * (gs * gs);      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) * a.inv_sqrt_bc2 + a.eps;
    p = p - a.lr_over_bc1 * (m / denom);
    if (a.zero_grad) g = 0.f;
}

constexpr int ADAM_THREADS = 256;

// Device-resident optimizer state for a step the host never synchronises with (CUDA-graph replays, optimistic list
// capacities): the step count and the bias corrections live on the device, and a step whose render was incomplete is
// vetoed there -- parameters and moments stay, the gradient slice is still cleared.
struct AdamState {
    float lr_over_bc1, inv_sqrt_bc2;
    int step;          // optimizer steps applied so far
    int skip;          // 1 = the current step is vetoed
    int seen;          // steps prepared so far (applied + vetoed)
    int vetoed;        // steps vetoed so far
    int ring[10];      // `seen` index (0-based) of the last 10 vetoed steps, slot = vetoed % 10
};
static_assert(sizeof(AdamState) == 64, "AdamState is 64 bytes (B200_ADAM_STATE_BYTES)");

__global__ void adam_prepare_kernel(AdamState *st, int32_t *veto_flag, double lr, double beta1, double beta2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const bool veto = veto_flag && *veto_flag != 0;
    if (veto_flag) *veto_flag = 0;  // consumed: the next step starts clean
    const int idx = st->seen;
    st->seen = idx + 1;
    st->skip = veto ? 1 : 0;
    if (veto) {
        st->ring[st->vetoed % 10] = idx;
        st->vetoed += 1;
        return;
    }
    const int step = st->step + 1;
    st->step = step;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    st->lr_over_bc1 = (float)(lr / bc1);
    st->inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
}

__global__ void __launch_bounds__(ADAM_THREADS) adam_state_kernel(long long n, float *__restrict__ p, float *__restrict__ g,
                                                                  float *__restrict__ m, float *__restrict__ v, AdamArgs a,
                                                                  const AdamState *__restrict__ st) {
    a.lr_over_bc1 = st->lr_over_bc1; a.inv_sqrt_bc2 = st->inv_sqrt_bc2;
    const bool skip = st->skip != 0;
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    float4 *p4 = reinterpret_cast<float4 *>(p), *g4 = reinterpret_cast<float4 *>(g);
    float4 *m4 = reinterpret_cast<float4 *>(m), *v4 = reinterpret_cast<float4 *>(v);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        if (skip) {
            if (a.zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        float4 P = p4[i], G = g4[i], M = m4[i], V = v4[i];
        adam_one(P.x, G.x, M.x, V.x, a); adam_one(P.y, G.y, M.y, V.y, a);
        adam_one(P.z, G.z, M.z, V.z, a); adam_one(P.w, G.w, M.w, V.w, a);
        p4[i] = P; m4[i] = M; v4[i] = V;
        if (a.zero_grad) g4[i] = G;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        if (skip) { if (a.zero_grad) g[i] = 0.f; }
        else adam_one(p[i], g[i], m[i], v[i], a);
    }
}

__global__ void __launch_bounds__(ADAM_THREADS) adam_kernel(long long n, float *__restrict__ p, float *__restrict__ g,
                                                            float *__restrict__ m, float *__restrict__ v, AdamArgs a) {
    // one 16-byte vector per thread, many short CTAs: a collective kernel of the next gradient slice (NCCL) can take
    // SMs as they free up instead of waiting behind persistent grid-stride CTAs
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    float4 *p4 = reinterpret_cast<float4 *>(p), *g4 = reinterpret_cast<float4 *>(g);
    float4 *m4 = reinterpret_cast<float4 *>(m), *v4 = reinterpret_cast<float4 *>(v);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = p4[i], G = g4[i], M = m4[i], V = v4[i];
        adam_one(P.x, G.x, M.x, V.x, a); adam_one(P.y, G.y, M.y, V.y, a);
        adam_one(P.z, G.z, M.z, V.z, a); adam_one(P.w, G.w, M.w, V.w, a);
        p4[i] = P; m4[i] = M; v4[i] = V;
        if (a.zero_grad) g4[i] = G;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        adam_one(p[i], g[i], m[i], v[i], a);
    }
}

// head elements before the first 16-byte boundary of an arbitrarily offset slice (at most 3)
__global__ void adam_head_kernel(int n, float *p, float *g, float *m, float *v, AdamArgs a) {
    if ((int)threadIdx.x < n) adam_one(p[threadIdx.x], g[threadIdx.x], m[threadIdx.x], v[threadIdx.x], a);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_adam_step(long long numel, float *param, float *grad, float *exp_avg, float *exp_avg_sq, int step,
                              double lr, double beta1, double beta2, double eps, double grad_scale, int zero_grad,
                              void *stream) {
    B200_REQUIRE(numel >= 0, "numel must be >= 0");
    if (numel == 0) return B200_OK;
    B200_REQUIRE(param && grad && exp_avg && exp_avg_sq, "null pointer");
    B200_REQUIRE(step >= 1, "step counts from 1");
    B200_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0, "betas must lie in [0, 1)");
    const uintptr_t mis = reinterpret_cast<uintptr_t>(param) & 15u;
    B200_REQUIRE((reinterpret_cast<uintptr_t>(grad) & 15u) == mis && (reinterpret_cast<uintptr_t>(exp_avg) & 15u) == mis &&
                     (reinterpret_cast<uintptr_t>(exp_avg_sq) & 15u) == mis && (mis & 3u) == 0,
                 "the four buffers must share their alignment modulo 16 bytes (same offset into equally aligned flat buffers)");
    AdamArgs a;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    a.lr_over_bc1 = (float)(lr / bc1);
    a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    a.one_minus_beta1 = (float)(1.0 - beta1); a.beta2 = (float)beta2; a.one_minus_beta2 = (float)(1.0 - beta2);
    a.eps = (float)eps; a.grad_scale = (float)grad_scale; a.zero_grad = zero_grad ? 1 : 0;
    cudaStream_t st = as_stream(stream);
    long long head = mis ? (long long)((16u - mis) >> 2) : 0;
    if (head > numel) head = numel;
    if (head) {
        adam_head_kernel<<<1, 32, 0, st>>>((int)head, param, grad, exp_avg, exp_avg_sq, a);
        B200_LAUNCH_CHECK();
        param += head; grad += head; exp_avg += head; exp_avg_sq += head; numel -= head;
        if (numel == 0) return B200_OK;
    }
    const long long n4 = numel >> 2;
    long long want = (n4 + ADAM_THREADS - 1) / ADAM_THREADS;
    const long long cap = 1ll << 30;
    const int blocks = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    adam_kernel<<<blocks, ADAM_THREADS, 0, st>>>(numel, param, grad, exp_avg, exp_avg_sq, a);
    B200_LAUNCH_CHECK();
    return B200_OK;
}


extern "C" size_t b200_adam_state_bytes(void) { return sizeof(AdamState); }

// Once per optimisation step, before the first b200_adam_step_state of that step.  `state` = DEVICE, b200_adam_state_bytes()
// bytes, zero-initialised by the caller before the first step.  veto_flag (DEVICE int32, may be null): non-zero = this
// step's render was incomplete (b200_bin_cull_emit_capacity status[0], reduced with MAX across ranks by a data-parallel
// caller): the step is skipped and the flag cleared.  state int32 words [2] step, [4] steps seen, [5] steps vetoed,
// [6..15] ring of vetoed step indices let the host find out later which images to repeat.
extern "C" int b200_adam_prepare(void *state, int32_t *veto_flag, double lr, double beta1, double beta2, void *stream) {
    B200_REQUIRE(state, "null state");
    B200_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0, "betas must lie in [0, 1)");
    adam_prepare_kernel<<<1, 32, 0, as_stream(stream)>>>(static_cast<AdamState *>(state), veto_flag, lr, beta1, beta2);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

// b200_adam_step with the step count / bias corrections / veto read from the device state prepared above.
static int adam_state_launch(long long numel, float *param, float *grad, float *exp_avg, float *exp_avg_sq, const void *state,
                             double beta1, double beta2, double eps, double grad_scale, int zero_grad, long long max_blocks,
                             void *stream) {
    B200_REQUIRE(numel >= 0, "numel must be >= 0");
    if (numel == 0) return B200_OK;
    B200_REQUIRE(param && grad && exp_avg && exp_avg_sq && state, "null pointer");
    const uintptr_t mis = reinterpret_cast<uintptr_t>(param) & 15u;
    B200_REQUIRE(mis == 0 && (reinterpret_cast<uintptr_t>(grad) & 15u) == 0 && (reinterpret_cast<uintptr_t>(exp_avg) & 15u) == 0 &&
                     (reinterpret_cast<uintptr_t>(exp_avg_sq) & 15u) == 0,
                 "the slice must start on a 16-byte boundary of the four buffers");
    AdamArgs a;
    a.lr_over_bc1 = 0.f; a.inv_sqrt_bc2 = 0.f;
    a.one_minus_beta1 = (float)(1.0 - beta1); a.beta2 = (float)beta2; a.one_minus_beta2 = (float)(1.0 - beta2);
    a.eps = (float)eps; a.grad_scale = (float)grad_scale; a.zero_grad = zero_grad ? 1 : 0;
    const long long n4 = numel >> 2;
    long long want = (n4 + ADAM_THREADS - 1) / ADAM_THREADS;
    const int blocks = (int)(want < 1 ? 1 : (want > max_blocks ? max_blocks : want));
    adam_state_kernel<<<blocks, ADAM_THREADS, 0, as_stream(stream)>>>(numel, param, grad, exp_avg, exp_avg_sq, a,
                                                                     static_cast<const AdamState *>(state));
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_adam_step_state(long long numel, float *param, float *grad, float *exp_avg, float *exp_avg_sq,
                                    const void *state, double beta1, double beta2, double eps, double grad_scale,
                                    int zero_grad, void *stream) {
    return adam_state_launch(numel, param, grad, exp_avg, exp_avg_sq, state, beta1, beta2, eps, grad_scale, zero_grad, 1ll << 30,
                             stream);
}

// The same update as a BACKGROUND kernel: at most ctas_per_sm resident CTAs per SM looping over the slice, so that
// kernels launched beside it on other streams (whatever their priority; memset nodes of a captured graph carry none)
// always find free thread slots instead of queueing behind a grid of thousands of short CTAs.
extern "C" int b200_adam_step_state_background(long long numel, float *param, float *grad, float *exp_avg, float *exp_avg_sq,
                                               const void *state, double beta1, double beta2, double eps, double grad_scale,
                                               int zero_grad, int ctas_per_sm, void *stream) {
    B200_REQUIRE(ctas_per_sm >= 1 && ctas_per_sm <= 32, "ctas_per_sm must be in [1, 32]");
    int dev = 0, sms = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    return adam_state_launch(numel, param, grad, exp_avg, exp_avg_sq, state, beta1, beta2, eps, grad_scale, zero_grad,
                             (long long)sms * ctas_per_sm, stream);
}
