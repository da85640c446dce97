// clip_grad_norm_(max_norm) + SGD(momentum, weight_decay) over ONE flat fp32 parameter arena
// (reference tools/train_ssl_reg.py:239-241, 174-175), and ExponentialMovingAverage.update
// (regda/utils/ema.py:46-51).  HBM-bound streaming kernels: 256-bit accesses, grid sized to the
// SM count, fixed-order two-stage reduction for the norm.
#include <cuda_bf16.h>

#include "common.cuh"

namespace regda {
namespace {

constexpr int kOptThreads = 256;

__global__ void __launch_bounds__(kOptThreads)
sumsq_partial_kernel(const float *__restrict__ x, long long n, float *__restrict__ partial) {
    float s = 0.f;
    const long long n4 = n / 4;
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    for (long long i = static_cast<long long>(blockIdx.x) * kOptThreads + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * kOptThreads) {
        const float4 v = x4[i];
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < n - n4 * 4) { const float v = x[n4 * 4 + threadIdx.x]; s += v * v; }
    __shared__ float red[kOptThreads / 32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < kOptThreads / 32; ++i) t += red[i];
        partial[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256)
sumsq_final_kernel(const float *__restrict__ partial, int n, float *__restrict__ out, int accumulate) {
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += static_cast<double>(partial[i]);
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + static_cast<float>(red[0]);
}

struct SgdArgs {
    float *param;
    const float *grad;
    float *buf;
    __nv_bfloat16 *shadow;
    const float *sumsq;
    const float *lr_dev;
    long long n;
    float max_norm, lr, momentum, wd, grad_scale;
    int first_step;
};

__device__ __forceinline__ float sgd_one(float p, float g, float &b, float coef, float lr, const SgdArgs &a) {
    g = g * coef + a.wd * p;                       // clip_grad_norm_ scaling, then weight decay (torch SGD: d_p += wd*p)
    b = a.first_step ? g : a.momentum * b + g;     // momentum buffer (dampening 0)
    return p - lr * b;
}

__global__ void __launch_bounds__(kOptThreads)
sgd_kernel(const SgdArgs a) {
    // total_norm = sqrt(sum g^2) of the (already grad_scale-d) gradient; coef = min(1, max_norm/(norm+1e-6))
    float coef = a.grad_scale;
    if (a.sumsq != nullptr) {
        const float norm = sqrtf(a.sumsq[0]) * a.grad_scale;
        coef = a.grad_scale * fminf(1.0f, a.max_norm / (norm + 1e-6f));
    }
    const float lr = a.lr_dev != nullptr ? a.lr_dev[0] : a.lr;
    const long long n4 = a.n / 4;
    for (long long i = static_cast<long long>(blockIdx.x) * kOptThreads + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * kOptThreads) {
        float4 p = reinterpret_cast<float4 *>(a.param)[i];
        const float4 g = reinterpret_cast<const float4 *>(a.grad)[i];
        float4 b = a.first_step ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4 *>(a.buf)[i];
        p.x = sgd_one(p.x, g.x, b.x, coef, lr, a); p.y = sgd_one(p.y, g.y, b.y, coef, lr, a);
        p.z = sgd_one(p.z, g.z, b.z, coef, lr, a); p.w = sgd_one(p.w, g.w, b.w, coef, lr, a);
        reinterpret_cast<float4 *>(a.param)[i] = p;
        reinterpret_cast<float4 *>(a.buf)[i] = b;
        if (a.shadow != nullptr) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
            uint2 pk;
            pk.x = *reinterpret_cast<unsigned *>(&lo); pk.y = *reinterpret_cast<unsigned *>(&hi);
            reinterpret_cast<uint2 *>(a.shadow)[i] = pk;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < a.n - n4 * 4) {
        const long long i = n4 * 4 + threadIdx.x;
        float b = a.first_step ? 0.f : a.buf[i];
        const float p = sgd_one(a.param[i], a.grad[i], b, coef, lr, a);
        a.param[i] = p; a.buf[i] = b;
        if (a.shadow != nullptr) a.shadow[i] = __float2bfloat16_rn(p);
    }
}

__global__ void __launch_bounds__(kOptThreads)
ema_kernel(float *__restrict__ shadow, const float *__restrict__ param, long long n, float one_minus_decay, float decay) {
    for (long long i = static_cast<long long>(blockIdx.x) * kOptThreads + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * kOptThreads)
        shadow[i] = one_minus_decay * param[i] + decay * shadow[i];      // ema.py:49-50
}

int opt_blocks(long long n) {
    const long long want = (n / 4 + kOptThreads - 1) / kOptThreads;
    const long long cap = 8ll * sm_count();
    return static_cast<int>(std::max<long long>(1, std::min(want, cap)));
}

}  // namespace
}  // namespace regda

using namespace regda;

extern "C" size_t regda_sumsq_workspace_bytes(int64_t n) {
    (void)n;
    return align_up(static_cast<size_t>(8 * sm_count()) * 4, 256);
}

extern "C" int regda_sumsq(const float *x, int64_t n, float *sumsq_out, int accumulate, void *workspace, size_t workspace_bytes, void *stream) {
    if (n < 0 || !sumsq_out) return fail(REGDA_ERR_INVALID_ARG, "sumsq: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) { if (!accumulate) REGDA_CUDA_CHECK(cudaMemsetAsync(sumsq_out, 0, 4, st)); return REGDA_OK; }
    if (!x || (reinterpret_cast<uintptr_t>(x) & 15)) return fail(REGDA_ERR_INVALID_ARG, "sumsq: input must be 16-byte aligned");
    if (!workspace || workspace_bytes < regda_sumsq_workspace_bytes(n)) return fail(REGDA_ERR_WORKSPACE, "sumsq: workspace too small");
    const int blocks = opt_blocks(n);
    sumsq_partial_kernel<<<blocks, kOptThreads, 0, st>>>(x, n, static_cast<float *>(workspace));
    REGDA_LAUNCH_CHECK();
    sumsq_final_kernel<<<1, 256, 0, st>>>(static_cast<float *>(workspace), blocks, sumsq_out, accumulate);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

extern "C" int regda_sgd_step(float *param, const float *grad, float *momentum_buf, void *param_bf16, int64_t n,
                              const float *sumsq, double max_norm, double grad_scale, double lr, const float *lr_device,
                              double momentum, double weight_decay, int first_step, void *stream) {
    if (n < 0) return fail(REGDA_ERR_INVALID_ARG, "sgd_step: bad size");
    if (n == 0) return REGDA_OK;
    if (!param || !grad || !momentum_buf) return fail(REGDA_ERR_INVALID_ARG, "sgd_step: null pointer");
    if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(momentum_buf)) & 15)
        return fail(REGDA_ERR_INVALID_ARG, "sgd_step: arenas must be 16-byte aligned");
    if (param_bf16 && (reinterpret_cast<uintptr_t>(param_bf16) & 7)) return fail(REGDA_ERR_INVALID_ARG, "sgd_step: bf16 shadow must be 8-byte aligned");
    SgdArgs a;
    a.param = param; a.grad = grad; a.buf = momentum_buf; a.shadow = static_cast<__nv_bfloat16 *>(param_bf16);
    a.sumsq = sumsq; a.lr_dev = lr_device; a.n = n; a.max_norm = static_cast<float>(max_norm); a.lr = static_cast<float>(lr);
    a.momentum = static_cast<float>(momentum); a.wd = static_cast<float>(weight_decay); a.grad_scale = static_cast<float>(grad_scale);
    a.first_step = first_step;
    sgd_kernel<<<opt_blocks(n), kOptThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

extern "C" int regda_ema_update(float *shadow, const float *param, int64_t n, double decay, void *stream) {
    if (n < 0) return fail(REGDA_ERR_INVALID_ARG, "ema_update: bad size");
    if (n == 0) return REGDA_OK;
    if (!shadow || !param) return fail(REGDA_ERR_INVALID_ARG, "ema_update: null pointer");
    ema_kernel<<<opt_blocks(n), kOptThreads, 0, static_cast<cudaStream_t>(stream)>>>(shadow, param, n, static_cast<float>(1.0 - decay), static_cast<float>(decay));
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}
