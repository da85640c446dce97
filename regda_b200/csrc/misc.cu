// Small kernels around the convolutions of the PPM head and the strided layers:
//
//   regda_zero_insert2_bf16   dY of a stride-2 convolution -> the zero-inserted map whose stride-1 data gradient is the
//                             stride-2 one (regda/_resnets.py:92-112: layer2.0 / layer3.0 conv2 and downsample)
//   regda_dropout2d_mask      Dropout2d(p) keep/scale factors per (image, channel) (regda/models/Encoder.py:39), counter-based
//                             generator whose state lives in device memory (CUDA-graph replay draws fresh masks)
//   regda_classifier_fwd/bwd  Dropout2d -> Conv2d(512, C, 1) + bias, the last two layers of the PPM head
//                             (regda/models/Encoder.py:39-40): C is 6 or 7, so this is an HBM-bound streaming pass over the
//                             512-channel activation (one warp per pixel), not a tensor-core tile
#include <cuda_bf16.h>

#include "common.cuh"

namespace regda {
namespace {

struct alignas(16) bf16x8 { __nv_bfloat162 v[4]; };

__device__ __forceinline__ void load8(const __nv_bfloat16 *p, float (&f)[8]) {
    const bf16x8 v = *reinterpret_cast<const bf16x8 *>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(v.v[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ void load8(const float *p, float (&f)[8]) {
    const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8(__nv_bfloat16 *p, const float (&f)[8]) {
    bf16x8 v;
#pragma unroll
    for (int i = 0; i < 4; ++i) v.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<bf16x8 *>(p) = v;
}
__device__ __forceinline__ void store8(float *p, const float (&f)[8]) {
    *reinterpret_cast<float4 *>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4 *>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

// ---- zero insertion ------------------------------------------------------------------------------------------------
// dst [n][OH][OW][c] (already zeroed) : dst[n][2i][2j][:] = src[n][i][j][:]
__global__ void __launch_bounds__(256)
zero_insert2_kernel(const __nv_bfloat16 *__restrict__ src, __nv_bfloat16 *__restrict__ dst, int n, int oh, int ow, int OH, int OW, int c) {
    const int octs = c >> 3;
    const long long total = static_cast<long long>(n) * oh * ow * octs;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int o = static_cast<int>(i % octs);
        long long p = i / octs;
        const int x = static_cast<int>(p % ow); p /= ow;
        const int y = static_cast<int>(p % oh);
        const int img = static_cast<int>(p / oh);
        const uint4 v = *reinterpret_cast<const uint4 *>(src + i * 8);
        *reinterpret_cast<uint4 *>(dst + ((static_cast<long long>(img) * OH + 2 * y) * OW + 2 * x) * c + o * 8) = v;
    }
}

// ---- Dropout2d mask --------------------------------------------------------------------------------------------------
// state[0] = seed, state[1] = draw counter (advanced by the kernel: a replayed CUDA graph draws a new mask every step).
// Philox-style counter hash: two rounds of a 64-bit mix (splitmix64 finaliser) over (seed, counter, index).
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
dropout2d_mask_kernel(unsigned long long *__restrict__ state, float p, float *__restrict__ keep_scale, int n) {
    const unsigned long long seed = state[0], ctr = state[1];
    const float scale = 1.f / (1.f - p);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long h = mix64(mix64(seed + 0x9E3779B97F4A7C15ull * (ctr + 1)) ^ (static_cast<unsigned long long>(i) * 0xD1342543DE82EF95ull + 1));
        const float u = static_cast<float>(h >> 40) * (1.f / 16777216.f);         // 24 uniform bits in [0, 1)
        keep_scale[i] = u < p ? 0.f : scale;
    }
    __syncthreads();
    if (threadIdx.x == 0) state[1] = ctr + 1;
}

// ---- classifier ------------------------------------------------------------------------------------------------------
constexpr int kMaxClasses = 8;
constexpr int kClsThreads = 256;

// out[img][k][px] = bias[k] + sum_c W[k][c] * keep[img][c] * y[img][px][c].   grid (blocks per image, images); one warp per pixel.
template <typename T>
__global__ void __launch_bounds__(kClsThreads)
classifier_fwd_kernel(const T *__restrict__ y, const float *__restrict__ w, const float *__restrict__ bias, const float *__restrict__ keep,
                      float *__restrict__ out, int hw, int cin, int ncls) {
    extern __shared__ float sw[];                       // [ncls][cin], dropout factors folded in
    const int img = blockIdx.y;
    for (int i = threadIdx.x; i < ncls * cin; i += kClsThreads) {
        const int c = i % cin;
        sw[i] = w[i] * (keep != nullptr ? keep[static_cast<size_t>(img) * cin + c] : 1.f);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const T *yi = y + static_cast<size_t>(img) * hw * cin;
    for (int px = blockIdx.x * (kClsThreads / 32) + warp; px < hw; px += gridDim.x * (kClsThreads / 32)) {
        float acc[kMaxClasses];
#pragma unroll
        for (int k = 0; k < kMaxClasses; ++k) acc[k] = 0.f;
        for (int c = lane * 8; c < cin; c += 256) {
            float f[8];
            load8(yi + static_cast<size_t>(px) * cin + c, f);
#pragma unroll
            for (int k = 0; k < kMaxClasses; ++k) {
                if (k < ncls) {
                    const float4 a = *reinterpret_cast<const float4 *>(sw + k * cin + c), b = *reinterpret_cast<const float4 *>(sw + k * cin + c + 4);
                    acc[k] += f[0] * a.x + f[1] * a.y + f[2] * a.z + f[3] * a.w + f[4] * b.x + f[5] * b.y + f[6] * b.z + f[7] * b.w;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kMaxClasses; ++k) acc[k] = warp_sum(acc[k]);
        if (lane < ncls) {
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < kMaxClasses; ++k) v = lane == k ? acc[k] : v;
            out[(static_cast<size_t>(img) * ncls + lane) * hw + px] = v + (bias != nullptr ? bias[lane] : 0.f);
        }
    }
}

// dy[img][px][c] = keep[img][c] * sum_k W[k][c] * dout[img][k][px];  dW[k][c] += sum_px dout[k][px] * keep[c] * y[px][c];
// dbias[k] += sum_px dout[k][px].   grid (blocks per image, images); a thread owns 8 channels of every pixel it visits:
// threads [0, cin/8) of a "row group" cover one pixel, the block's kClsThreads / (cin/8) row groups walk the pixels.
template <typename T>
__global__ void __launch_bounds__(kClsThreads)
classifier_bwd_kernel(const T *__restrict__ y, const float *__restrict__ w, const float *__restrict__ keep, const float *__restrict__ dout,
                      T *__restrict__ dy, float *__restrict__ dw, float *__restrict__ dbias, int hw, int cin, int ncls, int px_per_block) {
    extern __shared__ float sm[];
    float *sd = sm;                                     // [px_per_block][kMaxClasses]: this block's slice of dout
    const int img = blockIdx.y;
    const int px0 = blockIdx.x * px_per_block;
    const int npx = min(px_per_block, hw - px0);
    for (int i = threadIdx.x; i < npx * kMaxClasses; i += kClsThreads) {
        const int p = i / kMaxClasses, k = i % kMaxClasses;
        sd[i] = k < ncls ? dout[(static_cast<size_t>(img) * ncls + k) * hw + px0 + p] : 0.f;
    }
    __syncthreads();
    const int octs = cin >> 3;
    const int groups = kClsThreads / octs;              // pixel rows in flight
    const int o = threadIdx.x % octs, grp = threadIdx.x / octs;
    float wk[kMaxClasses][8], kp[8], gw[kMaxClasses][8];
    {
        // this thread's 8 channels of the dropout factors and of every class's weight row: 16-byte loads (cin is a multiple of 8 and
        // the tensors come 16-byte aligned; the scalar form issues 72 loads per thread at a lane stride of 32 bytes)
        const bool vec = ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(keep)) & 15) == 0;
        auto load8w = [&](const float *p, float (&v)[8]) {
            if (vec) {
                const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b4 = __ldg(reinterpret_cast<const float4 *>(p) + 1);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b4.x; v[5] = b4.y; v[6] = b4.z; v[7] = b4.w;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = p[j];
            }
        };
#pragma unroll
        for (int j = 0; j < 8; ++j) kp[j] = 1.f;
        if (keep != nullptr) load8w(keep + static_cast<size_t>(img) * cin + o * 8, kp);
#pragma unroll
        for (int k = 0; k < kMaxClasses; ++k) {
            float wr[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) wr[j] = 0.f;
            if (k < ncls) load8w(w + static_cast<size_t>(k) * cin + o * 8, wr);
#pragma unroll
            for (int j = 0; j < 8; ++j) { wk[k][j] = wr[j] * kp[j]; gw[k][j] = 0.f; }
        }
    }
    {
        const T *yi = y + (static_cast<size_t>(img) * hw + px0) * cin + o * 8;
        T *dyi = dy + (static_cast<size_t>(img) * hw + px0) * cin + o * 8;
        for (int p = grp; p < npx; p += groups) {
            float f[8], g[8];
            load8(yi + static_cast<size_t>(p) * cin, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = 0.f;
#pragma unroll
            for (int k = 0; k < kMaxClasses; ++k) {
                const float d = sd[p * kMaxClasses + k];
#pragma unroll
                for (int j = 0; j < 8; ++j) { g[j] = fmaf(d, wk[k][j], g[j]); gw[k][j] = fmaf(d, f[j], gw[k][j]); }
            }
            store8(dyi + static_cast<size_t>(p) * cin, g);
        }
    }
    __syncthreads();
    if (dbias != nullptr && threadIdx.x < ncls) {
        float t = 0.f;
        for (int p = 0; p < npx; ++p) t += sd[p * kMaxClasses + threadIdx.x];
        atomicAdd(dbias + threadIdx.x, t);
    }
    // weight gradient: the block's `groups` row groups hold partial sums for the same (class, channel) -- fold them in shared
    // memory first (the staged dout slice is dead by now), then ONE global atomic per (class, channel) per block
    __syncthreads();
    float *sg = sm;                                     // [groups][ncls][cin] would not fit: fold class by class
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k) {
        if (k >= ncls) break;                           // (block-uniform: the barriers below stay convergent)
#pragma unroll
        for (int j = 0; j < 8; ++j) sg[grp * cin + o * 8 + j] = gw[k][j] * kp[j];
        __syncthreads();
        for (int c = threadIdx.x; c < cin; c += kClsThreads) {
            float t = 0.f;
            for (int gi = 0; gi < groups; ++gi) t += sg[gi * cin + c];
            atomicAdd(dw + k * cin + c, t);
        }
        __syncthreads();
    }
}

}  // namespace
}  // namespace regda

using namespace regda;

// src bf16 [n][oh][ow][c] -> dst bf16 [n][OH][OW][c], dst[n][2i][2j] = src[n][i][j], zero elsewhere (OH >= 2*oh-1, OW >= 2*ow-1)
extern "C" int regda_zero_insert2_bf16(const void *src, void *dst, int n, int oh, int ow, int OH, int OW, int c, void *stream) {
    if (!src || !dst || n < 1 || oh < 1 || ow < 1 || c < 8 || c % 8 || OH < 2 * oh - 1 || OW < 2 * ow - 1)
        return fail(REGDA_ERR_INVALID_ARG, "zero_insert2: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    REGDA_CUDA_CHECK(cudaMemsetAsync(dst, 0, static_cast<size_t>(n) * OH * OW * c * 2, st));
    const long long total = static_cast<long long>(n) * oh * ow * (c / 8);
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 8ll * sm_count()));
    zero_insert2_kernel<<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16 *>(src), static_cast<__nv_bfloat16 *>(dst), n, oh, ow, OH, OW, c);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

// keep_scale float32 [n]: 0 with probability p, else 1/(1-p).  state: uint64[2] in device memory = {seed, draw counter}.
extern "C" int regda_dropout2d_mask(void *state_u64x2, double p, float *keep_scale, int n, void *stream) {
    if (!state_u64x2 || !keep_scale || n < 1 || !(p >= 0.0 && p < 1.0)) return fail(REGDA_ERR_INVALID_ARG, "dropout2d_mask: bad arguments");
    dropout2d_mask_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<unsigned long long *>(state_u64x2), static_cast<float>(p),
                                                                          keep_scale, n);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

namespace {
int cls_args_ok(const void *y, const float *w, int b, int hw, int cin, int ncls) {
    if (!y || !w || b < 1 || hw < 1 || ncls < 1 || ncls > kMaxClasses) return 0;
    if (cin < 8 || cin % 8) return 0;
    const int octs = cin / 8;                       // threads per pixel row in the backward; the forward keeps [ncls][cin] floats in smem
    return octs <= kClsThreads && kClsThreads % octs == 0 && static_cast<size_t>(ncls) * cin * sizeof(float) <= 48 * 1024;
}
}  // namespace

// y [b][hw][cin] (bf16, or float32 when y_is_f32), w float32 [ncls][cin], bias float32 [ncls] or NULL, keep float32 [b][cin] or NULL
// -> out float32 [b][ncls][hw] (the reference's NCHW logits)
extern "C" int regda_classifier_fwd(const void *y, int y_is_f32, const float *w, const float *bias, const float *keep, float *out, int b,
                                    int hw, int cin, int ncls, void *stream) {
    if (!cls_args_ok(y, w, b, hw, cin, ncls) || !out) return fail(REGDA_ERR_INVALID_ARG, "classifier_fwd: bad arguments (cin / 8 must divide 256, <= 8 classes)");
    // ~8 resident blocks per SM: a block's prologue (weights x dropout factors into shared memory) is a chain of global-load latencies,
    // and a warp then walks its pixels one after the other -- many short blocks hide both (2 blocks per SM: 27 us per head)
    const int per_img = std::max(1, std::min((hw + 7) / 8, 8 * sm_count() / b + 1));
    const size_t smem = static_cast<size_t>(ncls) * cin * sizeof(float);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (y_is_f32)
        classifier_fwd_kernel<float><<<dim3(per_img, b), kClsThreads, smem, st>>>(static_cast<const float *>(y), w, bias, keep, out, hw, cin, ncls);
    else
        classifier_fwd_kernel<__nv_bfloat16><<<dim3(per_img, b), kClsThreads, smem, st>>>(static_cast<const __nv_bfloat16 *>(y), w, bias, keep, out, hw, cin, ncls);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

// dout float32 [b][ncls][hw] -> dy [b][hw][cin] (same type as y, written), dw float32 [ncls][cin] and dbias float32 [ncls] ACCUMULATED
extern "C" int regda_classifier_bwd(const void *y, int y_is_f32, const float *w, const float *keep, const float *dout, void *dy, float *dw,
                                    float *dbias, int b, int hw, int cin, int ncls, void *stream) {
    if (!cls_args_ok(y, w, b, hw, cin, ncls) || !dout || !dy || !dw) return fail(REGDA_ERR_INVALID_ARG, "classifier_bwd: bad arguments");
    const int px_per_block = 128;         // (measured: 32 pixels per block -- 512 blocks instead of 128 -- is slower, 42 vs 31 us: four times the weight-gradient atomics)
    const int per_img = (hw + px_per_block - 1) / px_per_block;
    const size_t smem = std::max(static_cast<size_t>(px_per_block) * kMaxClasses, static_cast<size_t>(kClsThreads / (cin / 8)) * cin) * sizeof(float);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (y_is_f32)
        classifier_bwd_kernel<float><<<dim3(per_img, b), kClsThreads, smem, st>>>(static_cast<const float *>(y), w, keep, dout, static_cast<float *>(dy), dw,
                                                                                   dbias, hw, cin, ncls, px_per_block);
    else
        classifier_bwd_kernel<__nv_bfloat16><<<dim3(per_img, b), kClsThreads, smem, st>>>(static_cast<const __nv_bfloat16 *>(y), w, keep, dout,
                                                                                           static_cast<__nv_bfloat16 *>(dy), dw, dbias, hw, cin, ncls, px_per_block);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}
