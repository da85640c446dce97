// Pyramid pooling front end of the PPM heads (reference regda/models/Encoder.py:43-52): for every
// pool scale s in (1,2,3,6), AdaptiveAvgPool2d(s) of the feature map and, after the tiny per-branch
// 1x1 conv + BN + ReLU, bilinear upsampling (align_corners=False) back to the feature-map size and
// concatenation with the feature map itself.  Replaces the ATen adaptive_avg_pool2d /
// upsample_bilinear2d / cat kernels (34 % of the step's GPU time in the eager profile,
// profiles/step_kernel_table_round1_cudnn_eager.txt) with four HBM-bound NHWC kernels:
//
//   ppm_pool_fwd   feat bf16 [b][h][w][c]            -> pooled f32 [b][ncell][c]   (ALL scales in one launch)
//   ppm_pool_bwd   dpooled f32 [b][ncell][c]          -> dfeat  bf16 [b][h][w][c]
//   ppm_upcat_fwd  feat + branch maps bf16 [b][s][s][cb] -> cat bf16 [b][h][w][c + nscale*cb]
//   ppm_upcat_bwd  dcat -> dbranch f32 [b][s][s][cb] per scale   (d feat is the first c channels of dcat)
//
// ncell = sum s^2 (50 for (1,2,3,6)); cells of scale k start at off_k = sum_{j<k} s_j^2, row-major.
// A thread owns 8 consecutive channels (16-byte accesses); one block per (image, feature-map row).
#include <cuda_bf16.h>

#include "common.cuh"

namespace regda {
namespace {

constexpr int kMaxScales = 4;
struct PpmScales {
    int n;
    int s[kMaxScales];
    int off[kMaxScales];
    int ncell;
};

struct alignas(16) bf16x8 { __nv_bfloat162 v[4]; };
__device__ __forceinline__ void unpack(const bf16x8 &v, float (&f)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(v.v[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ bf16x8 pack(const float (&f)[8]) {
    bf16x8 v;
#pragma unroll
    for (int i = 0; i < 4; ++i) v.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return v;
}
__device__ __forceinline__ void atomic_add8(float *p, const float (&f)[8], float w) {
    atomicAdd(reinterpret_cast<float4 *>(p), make_float4(f[0] * w, f[1] * w, f[2] * w, f[3] * w));
    atomicAdd(reinterpret_cast<float4 *>(p) + 1, make_float4(f[4] * w, f[5] * w, f[6] * w, f[7] * w));
}

// adaptive pooling window of output index i: [floor(i*n/s), ceil((i+1)*n/s))
__device__ __forceinline__ int win_lo(int i, int n, int s) { return (i * n) / s; }
__device__ __forceinline__ int win_hi(int i, int n, int s) { return ((i + 1) * n + s - 1) / s; }

// grid (h, b), block c/8 threads (<= 1024; larger c loops)
__global__ void __launch_bounds__(256)
ppm_pool_fwd_kernel(const __nv_bfloat16 *__restrict__ feat, float *__restrict__ pooled, int h, int w, int c, const PpmScales sc) {
    const int y = blockIdx.x, img = blockIdx.y;
    const __nv_bfloat16 *row = feat + (static_cast<size_t>(img) * h + y) * w * c;
    float *pimg = pooled + static_cast<size_t>(img) * sc.ncell * c;
    for (int ch = threadIdx.x * 8; ch < c; ch += blockDim.x * 8) {
        for (int k = 0; k < sc.n; ++k) {
            const int s = sc.s[k];
            // the row-cells this feature-map row belongs to (windows may overlap by one row)
            const int i0 = (y * s) / h;
            for (int j = 0; j < s; ++j) {
                const int x0 = win_lo(j, w, s), x1 = win_hi(j, w, s);
                float acc[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) acc[t] = 0.f;
                for (int x = x0; x < x1; ++x) {
                    float f[8];
                    unpack(*reinterpret_cast<const bf16x8 *>(row + static_cast<size_t>(x) * c + ch), f);
#pragma unroll
                    for (int t = 0; t < 8; ++t) acc[t] += f[t];
                }
                for (int i = max(i0 - 1, 0); i <= min(i0 + 1, s - 1); ++i) {
                    const int y0 = win_lo(i, h, s), y1 = win_hi(i, h, s);
                    if (y < y0 || y >= y1) continue;
                    const float wgt = 1.f / static_cast<float>((y1 - y0) * (x1 - x0));
                    atomic_add8(pimg + static_cast<size_t>(sc.off[k] + i * s + j) * c + ch, acc, wgt);
                }
            }
        }
    }
}

// grid (h, b, c / kPoolBwdChunk): for one feature-map row y and a 512-channel chunk, phase 1 folds the (at most two) row
// cells of every scale that contain y into per-column vectors R[column cell][channel] (12 x 512 floats in shared memory,
// 1/area included), phase 2 writes each pixel as the sum of the <= 2 column vectors per scale that contain x.
constexpr int kPoolBwdChunk = 512;
constexpr int kMaxColCells = 16;           // sum of the pool scales (1+2+3+6 = 12)
__global__ void __launch_bounds__(256)
ppm_pool_bwd_kernel(const float *__restrict__ dpooled, __nv_bfloat16 *__restrict__ dfeat, int h, int w, int c, const PpmScales sc) {
    __shared__ float R[kMaxColCells][kPoolBwdChunk];
    const int y = blockIdx.x, img = blockIdx.y, c0 = blockIdx.z * kPoolBwdChunk;
    const int cw = min(kPoolBwdChunk, c - c0);
    const float *pimg = dpooled + static_cast<size_t>(img) * sc.ncell * c + c0;
    int col = 0;
    for (int k = 0; k < sc.n; ++k) {
        const int s = sc.s[k];
        const int i0 = (y * s) / h;
        for (int j = 0; j < s; ++j, ++col) {
            const int x0 = win_lo(j, w, s), x1 = win_hi(j, w, s);
            for (int ch = threadIdx.x; ch < cw; ch += blockDim.x) {
                float acc = 0.f;
                for (int i = max(i0 - 1, 0); i <= min(i0 + 1, s - 1); ++i) {
                    const int y0 = win_lo(i, h, s), y1 = win_hi(i, h, s);
                    if (y < y0 || y >= y1) continue;
                    acc += pimg[static_cast<size_t>(sc.off[k] + i * s + j) * c + ch] / static_cast<float>((y1 - y0) * (x1 - x0));
                }
                R[col][ch] = acc;
            }
        }
    }
    __syncthreads();
    __nv_bfloat16 *row = dfeat + (static_cast<size_t>(img) * h + y) * w * c + c0;
    const int octs = cw >> 3;
    for (int item = threadIdx.x; item < w * octs; item += blockDim.x) {
        const int x = item / octs, ch = (item - x * octs) * 8;
        float acc[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] = 0.f;
        int base = 0;
        for (int k = 0; k < sc.n; ++k) {
            const int s = sc.s[k];
            const int j0 = (x * s) / w;
            for (int j = max(j0 - 1, 0); j <= min(j0 + 1, s - 1); ++j) {
                if (x < win_lo(j, w, s) || x >= win_hi(j, w, s)) continue;
                const float4 a = *reinterpret_cast<const float4 *>(&R[base + j][ch]);
                const float4 b = *reinterpret_cast<const float4 *>(&R[base + j][ch + 4]);
                acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
                acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
            }
            base += s;
        }
        *reinterpret_cast<bf16x8 *>(row + static_cast<size_t>(x) * c + ch) = pack(acc);
    }
}

// PyTorch's area_pixel_compute_source_index for align_corners=False: src = (dst + 0.5) * in/out - 0.5, clamped at 0
__device__ __forceinline__ void bilinear_src(int dst, int in_size, int out_size, int &i0, int &i1, float &l1) {
    const float scale = static_cast<float>(in_size) / static_cast<float>(out_size);
    float src = (static_cast<float>(dst) + 0.5f) * scale - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = static_cast<int>(src);
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = src - static_cast<float>(i0);
}

struct BranchPtrs { const __nv_bfloat16 *p[kMaxScales]; };
struct BranchGradPtrs { float *p[kMaxScales]; };

// grid (h, b, 1 + nscales): segment 0 copies the feature row into cat[..., 0..c); segment 1+k writes
// cat[..., c + k*cb + q] = bilinear(branch_k)[q] (the row's two source rows and weights are block-uniform)
__global__ void __launch_bounds__(256)
ppm_upcat_fwd_kernel(const __nv_bfloat16 *__restrict__ feat, const BranchPtrs br, __nv_bfloat16 *__restrict__ cat, int h, int w, int c,
                     int cb, const PpmScales sc) {
    const int y = blockIdx.x, img = blockIdx.y, seg = blockIdx.z;
    const int ctot = c + sc.n * cb;
    __nv_bfloat16 *crow = cat + (static_cast<size_t>(img) * h + y) * w * ctot;
    if (seg == 0) {
        const __nv_bfloat16 *frow = feat + (static_cast<size_t>(img) * h + y) * w * c;
        const int octs = c >> 3;
        for (int item = threadIdx.x; item < w * octs; item += blockDim.x) {
            const int x = item / octs, ch = (item - x * octs) * 8;
            *reinterpret_cast<bf16x8 *>(crow + static_cast<size_t>(x) * ctot + ch) =
                *reinterpret_cast<const bf16x8 *>(frow + static_cast<size_t>(x) * c + ch);
        }
        return;
    }
    const int k = seg - 1;
    int s = sc.s[0];
    const __nv_bfloat16 *bp = br.p[0];
    if (k == 1) { s = sc.s[1]; bp = br.p[1]; }
    if (k == 2) { s = sc.s[2]; bp = br.p[2]; }
    if (k == 3) { s = sc.s[3]; bp = br.p[3]; }
    int y0, y1;
    float ly;
    bilinear_src(y, s, h, y0, y1, ly);
    const __nv_bfloat16 *r0 = bp + (static_cast<size_t>(img) * s + y0) * s * cb;
    const __nv_bfloat16 *r1 = bp + (static_cast<size_t>(img) * s + y1) * s * cb;
    const int octs = cb >> 3;
    for (int item = threadIdx.x; item < w * octs; item += blockDim.x) {
        const int x = item / octs, q = (item - x * octs) * 8;
        int x0, x1;
        float lx;
        bilinear_src(x, s, w, x0, x1, lx);
        float f00[8], f01[8], f10[8], f11[8], o[8];
        unpack(*reinterpret_cast<const bf16x8 *>(r0 + static_cast<size_t>(x0) * cb + q), f00);
        unpack(*reinterpret_cast<const bf16x8 *>(r0 + static_cast<size_t>(x1) * cb + q), f01);
        unpack(*reinterpret_cast<const bf16x8 *>(r1 + static_cast<size_t>(x0) * cb + q), f10);
        unpack(*reinterpret_cast<const bf16x8 *>(r1 + static_cast<size_t>(x1) * cb + q), f11);
        const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = w00 * f00[t] + w01 * f01[t] + w10 * f10[t] + w11 * f11[t];
        *reinterpret_cast<bf16x8 *>(crow + static_cast<size_t>(x) * ctot + c + k * cb + q) = pack(o);
    }
}

// grid (h, b), one thread per (branch channel octet): for every source column j, the weighted sum of this row's
// gradients that interpolate from column j, then atomically into the (up to two) source rows.
__global__ void __launch_bounds__(256)
ppm_upcat_bwd_kernel(const __nv_bfloat16 *__restrict__ dcat, const BranchGradPtrs dbr, int h, int w, int c, int cb, const PpmScales sc) {
    const int y = blockIdx.x, img = blockIdx.y;
    const int ctot = c + sc.n * cb;
    const __nv_bfloat16 *drow = dcat + (static_cast<size_t>(img) * h + y) * w * ctot + c;
    for (int ch = threadIdx.x * 8; ch < sc.n * cb; ch += blockDim.x * 8) {
        const int k = ch / cb, q = ch - k * cb;
        const int s = sc.s[k];
        int y0, y1;
        float ly;
        bilinear_src(y, s, h, y0, y1, ly);
        float *g = dbr.p[k] + static_cast<size_t>(img) * s * s * cb + q;
        for (int j = 0; j < s; ++j) {
            float acc[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) acc[t] = 0.f;
            // destination columns whose source index is in (j-1, j+1): conservative bounds, exact test inside
            const float inv = static_cast<float>(w) / static_cast<float>(s);
            const int xa = max(0, static_cast<int>(floorf((static_cast<float>(j) - 0.5f) * inv - 0.5f)) - 1);
            const int xb = min(w - 1, static_cast<int>(ceilf((static_cast<float>(j) + 1.5f) * inv - 0.5f)) + 1);
            for (int x = xa; x <= xb; ++x) {
                int x0, x1;
                float lx;
                bilinear_src(x, s, w, x0, x1, lx);
                float wgt = 0.f;
                if (x0 == j) wgt += 1.f - lx;
                if (x1 == j) wgt += lx;
                if (wgt == 0.f) continue;
                float f[8];
                unpack(*reinterpret_cast<const bf16x8 *>(drow + static_cast<size_t>(x) * ctot + ch), f);
#pragma unroll
                for (int t = 0; t < 8; ++t) acc[t] = fmaf(f[t], wgt, acc[t]);
            }
            if (y0 == y1) {
                atomic_add8(g + static_cast<size_t>(y0 * s + j) * cb, acc, 1.f);
            } else {
                atomic_add8(g + static_cast<size_t>(y0 * s + j) * cb, acc, 1.f - ly);
                atomic_add8(g + static_cast<size_t>(y1 * s + j) * cb, acc, ly);
            }
        }
    }
}

int make_scales(const int *scales, int nscales, PpmScales *out) {
    if (nscales < 1 || nscales > kMaxScales || !scales) return fail(REGDA_ERR_INVALID_ARG, "ppm: 1..4 pool scales");
    out->n = nscales;
    int off = 0;
    for (int k = 0; k < kMaxScales; ++k) {
        out->s[k] = k < nscales ? scales[k] : 1;
        out->off[k] = off;
        if (k < nscales) {
            if (scales[k] < 1 || scales[k] > 64) return fail(REGDA_ERR_INVALID_ARG, "ppm: pool scale out of range");
            off += scales[k] * scales[k];
        }
    }
    out->ncell = off;
    return REGDA_OK;
}

}  // namespace
}  // namespace regda

using namespace regda;

extern "C" int regda_ppm_pool_fwd(const void *feat, float *pooled, int b, int h, int w, int c, const int *scales_host, int nscales,
                                  void *stream) {
    PpmScales sc;
    const int rc = make_scales(scales_host, nscales, &sc);
    if (rc) return rc;
    if (!feat || !pooled || b < 1 || h < 1 || w < 1 || c < 8 || c % 8) return fail(REGDA_ERR_INVALID_ARG, "ppm_pool_fwd: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    REGDA_CUDA_CHECK(cudaMemsetAsync(pooled, 0, static_cast<size_t>(b) * sc.ncell * c * sizeof(float), st));
    const int threads = std::min(256, (c / 8 + 31) / 32 * 32);
    ppm_pool_fwd_kernel<<<dim3(h, b), threads, 0, st>>>(static_cast<const __nv_bfloat16 *>(feat), pooled, h, w, c, sc);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

extern "C" int regda_ppm_pool_bwd(const float *dpooled, void *dfeat, int b, int h, int w, int c, const int *scales_host, int nscales,
                                  void *stream) {
    PpmScales sc;
    const int rc = make_scales(scales_host, nscales, &sc);
    if (rc) return rc;
    if (!dpooled || !dfeat || b < 1 || h < 1 || w < 1 || c < 8 || c % 8) return fail(REGDA_ERR_INVALID_ARG, "ppm_pool_bwd: bad arguments");
    int cols = 0;
    for (int k = 0; k < sc.n; ++k) cols += sc.s[k];
    if (cols > kMaxColCells) return fail(REGDA_ERR_UNSUPPORTED, "ppm_pool_bwd: pool scales sum to more than 16");
    ppm_pool_bwd_kernel<<<dim3(h, b, (c + kPoolBwdChunk - 1) / kPoolBwdChunk), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dpooled, static_cast<__nv_bfloat16 *>(dfeat), h, w, c, sc);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

extern "C" int regda_ppm_upcat_fwd(const void *feat, const void *br0, const void *br1, const void *br2, const void *br3, void *cat,
                                   int b, int h, int w, int c, int cb, const int *scales_host, int nscales, void *stream) {
    PpmScales sc;
    const int rc = make_scales(scales_host, nscales, &sc);
    if (rc) return rc;
    if (!feat || !cat || b < 1 || h < 1 || w < 1 || c % 8 || cb % 8 || c < 8 || cb < 8) return fail(REGDA_ERR_INVALID_ARG, "ppm_upcat_fwd: bad arguments");
    BranchPtrs bp;
    const void *ps[4] = {br0, br1, br2, br3};
    for (int k = 0; k < kMaxScales; ++k) {
        bp.p[k] = static_cast<const __nv_bfloat16 *>(ps[k]);
        if (k < nscales && !ps[k]) return fail(REGDA_ERR_INVALID_ARG, "ppm_upcat_fwd: null branch pointer");
    }
    ppm_upcat_fwd_kernel<<<dim3(h, b, 1 + nscales), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16 *>(feat), bp,
                                                                                      static_cast<__nv_bfloat16 *>(cat), h, w, c, cb, sc);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

extern "C" int regda_ppm_upcat_bwd(const void *dcat, float *dbr0, float *dbr1, float *dbr2, float *dbr3, int b, int h, int w, int c,
                                   int cb, const int *scales_host, int nscales, void *stream) {
    PpmScales sc;
    const int rc = make_scales(scales_host, nscales, &sc);
    if (rc) return rc;
    if (!dcat || b < 1 || h < 1 || w < 1 || c % 8 || cb % 8 || c < 8 || cb < 8) return fail(REGDA_ERR_INVALID_ARG, "ppm_upcat_bwd: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    BranchGradPtrs gp;
    float *ps[4] = {dbr0, dbr1, dbr2, dbr3};
    for (int k = 0; k < kMaxScales; ++k) {
        gp.p[k] = ps[k];
        if (k < nscales) {
            if (!ps[k]) return fail(REGDA_ERR_INVALID_ARG, "ppm_upcat_bwd: null branch gradient pointer");
            REGDA_CUDA_CHECK(cudaMemsetAsync(ps[k], 0, static_cast<size_t>(b) * sc.s[k] * sc.s[k] * cb * sizeof(float), st));
        }
    }
    const int threads = std::min(256, (nscales * cb / 8 + 31) / 32 * 32);
    ppm_upcat_bwd_kernel<<<dim3(h, b), threads, 0, st>>>(static_cast<const __nv_bfloat16 *>(dcat), gp, h, w, c, cb, sc);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

// =========================================================================================================
// Folded PPM fuse convolution (regda/models/Encoder.py:43-52 + the 3x3 conv of :33-34).
//
// The reference concatenates the feature map with four bilinearly UPSAMPLED s x s branch maps (s = 1, 2, 3, 6) and runs a 3x3
// convolution over all 4096 channels: 42.7 % of the model's multiply-adds, half of them over channels that carry only
// 1 + 4 + 9 + 36 = 50 distinct values per image.  Upsampling and convolution are both linear, so for the branch channels
//     conv(up(p_k))[px, o] = sum_tap sum_j B_k[px + tap, j] * G_k[j, tap, o],      G_k[j, tap, o] = sum_c p_k[j, c] W[o, tap, c_k + c]
// (B_k = bilinear interpolation weights, zero outside the map = the convolution's zero padding).  G is a tiny GEMM (50 cells per
// image), and the sum over (tap, cell) a [pixels x 450] . [450 x 512] GEMM per image with a CONSTANT left factor: both run on
// the tcgen05 convolution kernels as 1x1 convolutions (host side: regda_b200/ops/ppm_fold.py), the result is handed to the
// 3x3 convolution over the 2048 feature channels as its epilogue addend.  The kernels below are the layout glue:
//   regda_ppm_gather_weights   bf16 OHWI weight -> the dense feature-part weight and the four branch weights [(tap, o)][c]
//   regda_ppm_scatter_wgrad    their float32 gradients -> added into the OHWI gradient
//   regda_ppm_g_pack / unpack  G_k [(img, cell)][(tap, o)]  <->  GT [img][o][(cell, tap)] (K-major operand of the second GEMM)
//   regda_transpose_bf16       batched [r][c] -> [c][r]
// =========================================================================================================
namespace regda {
namespace {

struct GPtrs { __nv_bfloat16 *p[kMaxScales]; };

// grid (taps * o, 1 + nb): row (o, tap) of the OHWI weight -> segment 0: wmain[o][tap][0..cf), segment 1+k: wb[k][tap*O + o][0..cb)
__global__ void __launch_bounds__(256)
ppm_gather_weights_kernel(const __nv_bfloat16 *__restrict__ w, __nv_bfloat16 *__restrict__ wmain, const GPtrs wb, int O, int T, int ct, int cf, int cb) {
    const int row = blockIdx.x;                        // o * T + tap
    const int o = row / T, tap = row - o * T;
    const __nv_bfloat16 *src = w + static_cast<size_t>(row) * ct;
    if (blockIdx.y == 0) {
        if (wmain == nullptr) return;                   // the feature part is read in place (weight channel stride of the conv kernels)
        __nv_bfloat16 *dst = wmain + static_cast<size_t>(row) * cf;
        for (int c = threadIdx.x * 8; c < cf; c += blockDim.x * 8) *reinterpret_cast<uint4 *>(dst + c) = *reinterpret_cast<const uint4 *>(src + c);
    } else {
        const int k = blockIdx.y - 1;
        __nv_bfloat16 *dst = wb.p[k] + (static_cast<size_t>(tap) * O + o) * cb;
        const __nv_bfloat16 *s2 = src + cf + static_cast<size_t>(k) * cb;
        for (int c = threadIdx.x * 8; c < cb; c += blockDim.x * 8) *reinterpret_cast<uint4 *>(dst + c) = *reinterpret_cast<const uint4 *>(s2 + c);
    }
}

struct FPtrs { const float *p[kMaxScales]; };

__global__ void __launch_bounds__(256)
ppm_scatter_wgrad_kernel(const float *__restrict__ gmain, const FPtrs gwb, float *__restrict__ gw, int O, int T, int ct, int cf, int cb) {
    const int row = blockIdx.x;
    const int o = row / T, tap = row - o * T;
    float *dst = gw + static_cast<size_t>(row) * ct;
    if (blockIdx.y == 0) {
        if (gmain == nullptr) return;                   // the feature part was accumulated in place by the weight-gradient kernel
        const float *src = gmain + static_cast<size_t>(row) * cf;
        for (int c = threadIdx.x * 4; c < cf; c += blockDim.x * 4) {
            float4 d = *reinterpret_cast<float4 *>(dst + c);
            const float4 s4 = *reinterpret_cast<const float4 *>(src + c);
            d.x += s4.x; d.y += s4.y; d.z += s4.z; d.w += s4.w;
            *reinterpret_cast<float4 *>(dst + c) = d;
        }
    } else {
        const int k = blockIdx.y - 1;
        const float *src = gwb.p[k] + (static_cast<size_t>(tap) * O + o) * cb;
        float *d2 = dst + cf + static_cast<size_t>(k) * cb;
        for (int c = threadIdx.x * 4; c < cb; c += blockDim.x * 4) {
            float4 d = *reinterpret_cast<float4 *>(d2 + c);
            const float4 s4 = *reinterpret_cast<const float4 *>(src + c);
            d.x += s4.x; d.y += s4.y; d.z += s4.z; d.w += s4.w;
            *reinterpret_cast<float4 *>(d2 + c) = d;
        }
    }
}

// GT[img][o][kappa], kappa = (cell_global * T + tap) < ncell * T, zero up to kp.  One block per (img, o): the 450 source
// elements G_k[(img, cell)][tap * O + o] are a strided gather (2 B each, L2-resident: G is ~7 MB), the write is one dense row.
// PACK: G -> GT;  !PACK: GT -> G (the gradient's way back).
template <bool PACK>
__global__ void __launch_bounds__(128)
ppm_g_pack_kernel(const GPtrs g, __nv_bfloat16 *__restrict__ gt, int O, int T, int kp, const PpmScales sc) {
    const int o = blockIdx.x, img = blockIdx.y;
    __nv_bfloat16 *row = gt + (static_cast<size_t>(img) * O + o) * kp;
    const int nk = sc.ncell * T;
    for (int kappa = threadIdx.x; kappa < kp; kappa += blockDim.x) {
        if (kappa >= nk) {
            if (PACK) row[kappa] = __float2bfloat16(0.f);
            continue;
        }
        const int cell = kappa / T, tap = kappa - cell * T;
        int k = 0;
#pragma unroll
        for (int i = 1; i < kMaxScales; ++i) if (i < sc.n && cell >= sc.off[i]) k = i;
        const int j = cell - sc.off[k], s2 = sc.s[k] * sc.s[k];
        __nv_bfloat16 *e = g.p[k] + (static_cast<size_t>(img) * s2 + j) * (static_cast<size_t>(T) * O) + static_cast<size_t>(tap) * O + o;
        if (PACK) row[kappa] = *e; else *e = row[kappa];
    }
}

// dst[b][c][r] = src[b][r][c], 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16 *__restrict__ src, __nv_bfloat16 *__restrict__ dst, int rows, int cols) {
    __shared__ unsigned short tile[32][33];
    const size_t boff = static_cast<size_t>(blockIdx.z) * rows * cols;
    const unsigned short *s = reinterpret_cast<const unsigned short *>(src) + boff;
    unsigned short *d = reinterpret_cast<unsigned short *>(dst) + boff;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int r = r0 + ty + i, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + i][tx] = s[static_cast<size_t>(r) * cols + c];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int c = c0 + ty + i, r = r0 + tx;
        if (r < rows && c < cols) d[static_cast<size_t>(c) * rows + r] = tile[tx][ty + i];
    }
}

}  // namespace
}  // namespace regda

// w bf16 [O][T][ct] (OHWI), ct = cf + nb*cb -> wmain bf16 [O][T][cf] (may be NULL: not wanted), wb_k bf16 [T*O][cb] (row = tap*O + o), k < nb <= 4
extern "C" int regda_ppm_gather_weights(const void *w, void *wmain, void *wb0, void *wb1, void *wb2, void *wb3, int O, int T, int ct,
                                        int cf, int cb, int nb, void *stream) {
    if (!w || O < 1 || T < 1 || nb < 1 || nb > kMaxScales || cf % 8 || cb % 8 || ct != cf + nb * cb)
        return fail(REGDA_ERR_INVALID_ARG, "ppm_gather_weights: bad arguments");
    GPtrs p;
    void *ps[4] = {wb0, wb1, wb2, wb3};
    for (int k = 0; k < kMaxScales; ++k) {
        p.p[k] = static_cast<__nv_bfloat16 *>(ps[k]);
        if (k < nb && !ps[k]) return fail(REGDA_ERR_INVALID_ARG, "ppm_gather_weights: null branch weight");
    }
    ppm_gather_weights_kernel<<<dim3(O * T, 1 + nb), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16 *>(w), static_cast<__nv_bfloat16 *>(wmain), p, O, T, ct, cf, cb);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

// gw float32 [O][T][ct] += (gmain float32 [O][T][cf] (may be NULL), gwb_k float32 [T*O][cb])
extern "C" int regda_ppm_scatter_wgrad(const float *gmain, const float *g0, const float *g1, const float *g2, const float *g3, float *gw,
                                       int O, int T, int ct, int cf, int cb, int nb, void *stream) {
    if (!gw || O < 1 || T < 1 || nb < 1 || nb > kMaxScales || cf % 4 || cb % 4 || ct != cf + nb * cb)
        return fail(REGDA_ERR_INVALID_ARG, "ppm_scatter_wgrad: bad arguments");
    FPtrs p;
    const float *ps[4] = {g0, g1, g2, g3};
    for (int k = 0; k < kMaxScales; ++k) {
        p.p[k] = ps[k];
        if (k < nb && !ps[k]) return fail(REGDA_ERR_INVALID_ARG, "ppm_scatter_wgrad: null branch gradient");
    }
    ppm_scatter_wgrad_kernel<<<dim3(O * T, 1 + nb), 256, 0, static_cast<cudaStream_t>(stream)>>>(gmain, p, gw, O, T, ct, cf, cb);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

// pack != 0: G_k bf16 [b][s_k*s_k][T*O] -> GT bf16 [b][O][kp] (kappa = cell*T + tap, zero padded);  pack == 0: the reverse
extern "C" int regda_ppm_g_pack(void *g0, void *g1, void *g2, void *g3, void *gt, int b, int O, int T, int kp, const int *scales_host,
                                int nscales, int pack, void *stream) {
    PpmScales sc;
    const int rc = make_scales(scales_host, nscales, &sc);
    if (rc) return rc;
    if (!gt || b < 1 || O < 1 || T < 1 || kp < sc.ncell * T) return fail(REGDA_ERR_INVALID_ARG, "ppm_g_pack: bad arguments");
    GPtrs p;
    void *ps[4] = {g0, g1, g2, g3};
    for (int k = 0; k < kMaxScales; ++k) {
        p.p[k] = static_cast<__nv_bfloat16 *>(ps[k]);
        if (k < nscales && !ps[k]) return fail(REGDA_ERR_INVALID_ARG, "ppm_g_pack: null branch pointer");
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (pack) ppm_g_pack_kernel<true><<<dim3(O, b), 128, 0, st>>>(p, static_cast<__nv_bfloat16 *>(gt), O, T, kp, sc);
    else ppm_g_pack_kernel<false><<<dim3(O, b), 128, 0, st>>>(p, static_cast<__nv_bfloat16 *>(gt), O, T, kp, sc);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

// dst bf16 [batch][cols][rows] = src bf16 [batch][rows][cols]
extern "C" int regda_transpose_bf16(const void *src, void *dst, int batch, int rows, int cols, void *stream) {
    if (!src || !dst || batch < 1 || rows < 1 || cols < 1 || batch > 65535) return fail(REGDA_ERR_INVALID_ARG, "transpose: bad arguments");
    transpose_bf16_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32, batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16 *>(src), static_cast<__nv_bfloat16 *>(dst), rows, cols);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}
