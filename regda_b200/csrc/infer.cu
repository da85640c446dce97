// Inference-mode pieces of the model (the offline teacher pass gener_target_pseudo / pre_slide / tta_predict,
// regda/gast/pseudo_generation.py:96-141, regda/utils/tools.py:61-152, and evaluate(), regda/utils/eval.py:14-56):
//   * eval-mode BatchNorm2d (+ residual + ReLU) over channels-last bf16 with the RUNNING statistics
//     (regda/_resnets.py:92-112 in eval mode), one pass;
//   * the tail of Deeplabv2.forward in eval mode (regda/models/Encoder.py:152-155): both heads' logits upsampled bilinearly
//     (align_corners=True) to the input size, softmax over classes, mean of the two heads -- one pass, the four
//     [b,c,H,W] intermediates of the reference are never written.
// Both HBM-bound streaming kernels.
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"

namespace regda {
namespace {

struct alignas(16) bf16x8i { __nv_bfloat162 v[4]; };

template <bool RELU, bool RES>
__global__ void __launch_bounds__(256)
bn_inference_kernel(const __nv_bfloat16 *__restrict__ y, const __nv_bfloat16 *__restrict__ res, __nv_bfloat16 *__restrict__ out, long long total,
                    int c, const float *__restrict__ gamma, const float *__restrict__ beta, const float *__restrict__ rmean,
                    const float *__restrict__ rvar, float eps) {
    const long long stride = static_cast<long long>(gridDim.x) * 256 * 8;
    long long e = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 8;
    if (e >= total) return;
    const int ch = static_cast<int>(e % c);           // invariant: stride % c == 0 (c divides 2048)
    // per-channel constants of this thread's 8 channels: 16-byte loads where the pointers allow (a warp then reads whole lines; the
    // scalar form touches 32 sectors per instruction -- csrc/norm.cu load8f has the measurement)
    const bool vec = ((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(rmean) |
                       reinterpret_cast<uintptr_t>(rvar)) & 15) == 0 && (c & 3) == 0;
    auto load8 = [&](const float *p, float (&v)[8]) {
        if (vec) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(p + ch)), b = __ldg(reinterpret_cast<const float4 *>(p + ch) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = p[ch + i];
        }
    };
    float sc[8], sh[8], gm[8], bt[8], mu[8], vr[8];
    if (gamma) load8(gamma, gm);
    if (beta) load8(beta, bt);
    load8(rmean, mu);
    load8(rvar, vr);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sc[i] = (gamma ? gm[i] : 1.f) * rsqrtf(vr[i] + eps);
        sh[i] = fmaf(-mu[i], sc[i], beta ? bt[i] : 0.f);
    }
    for (; e < total; e += stride) {
        const bf16x8i yv = *reinterpret_cast<const bf16x8i *>(y + e);
        bf16x8i rv = yv;
        if (RES) rv = *reinterpret_cast<const bf16x8i *>(res + e);
        bf16x8i o;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(yv.v[i]);
            float a = fmaf(f.x, sc[2 * i], sh[2 * i]), b = fmaf(f.y, sc[2 * i + 1], sh[2 * i + 1]);
            if (RES) {
                const float2 r = __bfloat1622float2(rv.v[i]);
                a += r.x; b += r.y;
            }
            if (RELU) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
            o.v[i] = __floats2bfloat162_rn(a, b);
        }
        *reinterpret_cast<bf16x8i *>(out + e) = o;
    }
}

// one thread per output pixel: bilinear taps of all classes from the (L2-resident) low-resolution logits, softmax per head
template <int CMAX>
__global__ void __launch_bounds__(256)
upsample_softmax_mean_kernel(const float *__restrict__ x1, const float *__restrict__ x2, float *__restrict__ out, int b, int c, int h, int w,
                             int H, int W, float sy, float sx) {
    const long long npx = static_cast<long long>(b) * H * W;
    const int plane = h * w;
    for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < npx; p += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int X = static_cast<int>(p % W);
        const int Y = static_cast<int>((p / W) % H);
        const int img = static_cast<int>(p / (static_cast<long long>(W) * H));
        const float fy = sy * static_cast<float>(Y), fx = sx * static_cast<float>(X);
        const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
        const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1i = x0 + (x0 < w - 1 ? 1 : 0);
        const float wy1 = fy - static_cast<float>(y0), wy0 = 1.0f - wy1, wx1 = fx - static_cast<float>(x0), wx0 = 1.0f - wx1;
        float acc[CMAX];
#pragma unroll
        for (int j = 0; j < CMAX; ++j) acc[j] = 0.f;
        const int heads = x2 != nullptr ? 2 : 1;
        for (int hd = 0; hd < heads; ++hd) {
            const float *src = (hd == 0 ? x1 : x2) + static_cast<size_t>(img) * c * plane;
            float z[CMAX];
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < CMAX; ++j)
                if (j < c) {
                    const float *q = src + static_cast<size_t>(j) * plane;
                    z[j] = wy0 * (wx0 * q[y0 * w + x0] + wx1 * q[y0 * w + x1i]) + wy1 * (wx0 * q[y1 * w + x0] + wx1 * q[y1 * w + x1i]);
                    m = fmaxf(m, z[j]);
                }
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < CMAX; ++j)
                if (j < c) { z[j] = expf(z[j] - m); s += z[j]; }
            const float inv = 1.0f / s;
#pragma unroll
            for (int j = 0; j < CMAX; ++j)
                if (j < c) acc[j] += z[j] * inv;
        }
        const float scale = heads == 2 ? 0.5f : 1.0f;
        float *o = out + static_cast<size_t>(img) * c * H * W + static_cast<size_t>(Y) * W + X;
#pragma unroll
        for (int j = 0; j < CMAX; ++j)
            if (j < c) o[static_cast<size_t>(j) * H * W] = acc[j] * scale;
    }
}

}  // namespace
}  // namespace regda

using namespace regda;

extern "C" int regda_bn_inference_bf16(const void *y, const void *residual, void *out, int64_t npix, int c, const float *gamma,
                                       const float *beta, const float *running_mean, const float *running_var, double eps, int relu,
                                       void *stream) {
    if (npix <= 0 || c < 8 || c % 8 != 0 || 2048 % c != 0) return fail(REGDA_ERR_UNSUPPORTED, "bn_inference: channels must divide 2048 and be a multiple of 8");
    if (!y || !out || !running_mean || !running_var) return fail(REGDA_ERR_INVALID_ARG, "bn_inference: null pointer");
    const long long total = static_cast<long long>(npix) * c;
    const long long steps = (total + 2047) / 2048;
    const int blocks = static_cast<int>(std::min<long long>(steps, 8ll * sm_count()));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const __nv_bfloat16 *yy = static_cast<const __nv_bfloat16 *>(y), *rr = static_cast<const __nv_bfloat16 *>(residual);
    __nv_bfloat16 *oo = static_cast<__nv_bfloat16 *>(out);
    const float e = static_cast<float>(eps);
#define REGDA_BN_INF(R, S) bn_inference_kernel<R, S><<<blocks, 256, 0, st>>>(yy, rr, oo, total, c, gamma, beta, running_mean, running_var, e)
    if (relu) { if (rr) REGDA_BN_INF(true, true); else REGDA_BN_INF(true, false); }
    else { if (rr) REGDA_BN_INF(false, true); else REGDA_BN_INF(false, false); }
#undef REGDA_BN_INF
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

extern "C" int regda_upsample_softmax_mean(const float *x1, const float *x2, float *out, int b, int c, int h, int w, int H, int W,
                                           void *stream) {
    if (b < 0 || c < 1 || h < 1 || w < 1 || H < 1 || W < 1) return fail(REGDA_ERR_INVALID_ARG, "upsample_softmax_mean: bad shape");
    if (c > 16) return fail(REGDA_ERR_UNSUPPORTED, "upsample_softmax_mean: at most 16 classes");
    if (b == 0) return REGDA_OK;
    if (!x1 || !out) return fail(REGDA_ERR_INVALID_ARG, "upsample_softmax_mean: null pointer");
    const float sy = H > 1 ? static_cast<float>(h - 1) / static_cast<float>(H - 1) : 0.f;
    const float sx = W > 1 ? static_cast<float>(w - 1) / static_cast<float>(W - 1) : 0.f;
    const long long npx = static_cast<long long>(b) * H * W;
    const int blocks = static_cast<int>(std::min<long long>((npx + 255) / 256, 16ll * sm_count()));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (c <= 8) upsample_softmax_mean_kernel<8><<<blocks, 256, 0, st>>>(x1, x2, out, b, c, h, w, H, W, sy, sx);
    else upsample_softmax_mean_kernel<16><<<blocks, 256, 0, st>>>(x1, x2, out, b, c, h, w, H, W, sy, sx);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}
