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

// Fast pooling kernels for pyramids with at most 16 column cells over all scales (1 + 2 + 3 + 6 = 12): the column cells of all
// scales side by side, cell `col` = column j[col] of scale k[col], covering pixels [lo[col], hi[col]) of a feature-map row.
constexpr int kMaxColCells = 16;
struct PpmCols {
    int n;
    int k[kMaxColCells], j[kMaxColCells], lo[kMaxColCells], hi[kMaxColCells];
};

// Backward: dfeat[y][x] = addend[y][x] + sum over the cells containing (y, x) of dpooled[cell] / area(cell).
// grid (h, b, c / 512): one feature-map row and a 512-channel chunk.  Phase 1 folds, per column cell, the (at most two) row cells
// of its scale that contain y into R[col][512 channels] in shared memory (every thread 2 channels, all loads independent);
// phase 2 writes the row: a thread handles 8 (pixel, 8-channel) items -- a warp = 32 consecutive channel octets of ONE pixel, so
// the column-cell membership test is warp-uniform -- with the optional bf16 addend (the gradient the feature map's other readers
// sent) prefetched for all 8 items before the first sum.
constexpr int kPoolBwdChunk = 512;
constexpr int kPoolMaxW = 512;            // widest feature-map row of the fast backward kernel
__global__ void __launch_bounds__(256, 3)
ppm_pool_bwd_cols_kernel(const float *__restrict__ dpooled, const __nv_bfloat16 *__restrict__ addend, __nv_bfloat16 *__restrict__ dfeat,
                         int h, int w, int c, const PpmScales sc, const PpmCols cols) {
    __shared__ __align__(16) float R[kMaxColCells][kPoolBwdChunk];
    __shared__ int cell_of[kMaxColCells][2];          // the row cells (index into dpooled's cell axis) of column cell col that contain y
    __shared__ float winv[kMaxColCells][2];           // 1 / area of that cell; 0 for an unused slot (which then points at cell 0)
    __shared__ unsigned colmask[kPoolMaxW];           // bit col = pixel x lies in column cell col
    const int y = blockIdx.x, img = blockIdx.y, c0 = blockIdx.z * kPoolBwdChunk;
    const int cw = min(kPoolBwdChunk, c - c0);
    if (threadIdx.x < kMaxColCells) {
        const int col = threadIdx.x;
        cell_of[col][0] = cell_of[col][1] = 0;
        winv[col][0] = winv[col][1] = 0.f;
        if (col < cols.n) {
            const int k = cols.k[col], s = sc.s[k], i0 = (y * s) / h;
            int e = 0;
            for (int i = max(i0 - 1, 0); i <= min(i0 + 1, s - 1); ++i) {
                const int y0 = win_lo(i, h, s), y1 = win_hi(i, h, s);
                if (y < y0 || y >= y1 || e == 2) continue;
                cell_of[col][e] = sc.off[k] + i * s + cols.j[col];
                winv[col][e] = 1.0f / static_cast<float>((y1 - y0) * (cols.hi[col] - cols.lo[col]));
                ++e;
            }
        }
    }
    for (int x = threadIdx.x; x < w; x += blockDim.x) {
        unsigned m = 0u;
#pragma unroll
        for (int col = 0; col < kMaxColCells; ++col)
            if (col < cols.n && x >= cols.lo[col] && x < cols.hi[col]) m |= 1u << col;
        colmask[x] = m;
    }
    __syncthreads();
    const float *pimg = dpooled + static_cast<size_t>(img) * sc.ncell * c + c0;
    for (int ch = threadIdx.x * 2; ch < cw; ch += blockDim.x * 2) {
        float2 v[kMaxColCells][2];
#pragma unroll
        for (int col = 0; col < kMaxColCells; ++col)          // 32 independent loads (unused slots read cell 0 with weight 0)
#pragma unroll
            for (int e = 0; e < 2; ++e) v[col][e] = __ldg(reinterpret_cast<const float2 *>(pimg + static_cast<size_t>(cell_of[col][e]) * c + ch));
#pragma unroll
        for (int col = 0; col < kMaxColCells; ++col) {
            const float w0 = winv[col][0], w1 = winv[col][1];
            *reinterpret_cast<float2 *>(&R[col][ch]) = make_float2(v[col][0].x * w0 + v[col][1].x * w1, v[col][0].y * w0 + v[col][1].y * w1);
        }
    }
    __syncthreads();
    const size_t row_off = (static_cast<size_t>(img) * h + y) * w * c + c0;
    const int octs = cw >> 3, items = w * octs;
    for (int it0 = threadIdx.x; it0 < items; it0 += blockDim.x * 4) {
        bf16x8 av[4];
        if (addend != nullptr) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int item = it0 + u * blockDim.x;
                if (item < items) {
                    const int x = item / octs, oc = item - x * octs;
                    av[u] = *reinterpret_cast<const bf16x8 *>(addend + row_off + static_cast<size_t>(x) * c + oc * 8);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int item = it0 + u * blockDim.x;
            if (item >= items) break;
            const int x = item / octs, oc = item - x * octs;
            const unsigned m = colmask[x];                     // warp-uniform: a warp is 32 channel octets of one pixel
            float acc[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) acc[t] = 0.f;
#pragma unroll
            for (int col = 0; col < kMaxColCells; ++col) {
                if (m & (1u << col)) {
                    const float4 a = *reinterpret_cast<const float4 *>(&R[col][oc * 8]);
                    const float4 b2 = *reinterpret_cast<const float4 *>(&R[col][oc * 8 + 4]);
                    acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
                    acc[4] += b2.x; acc[5] += b2.y; acc[6] += b2.z; acc[7] += b2.w;
                }
            }
            if (addend != nullptr) {
                float f[8];
                unpack(av[u], f);
#pragma unroll
                for (int t = 0; t < 8; ++t) acc[t] += f[t];
            }
            *reinterpret_cast<bf16x8 *>(dfeat + row_off + static_cast<size_t>(x) * c + oc * 8) = pack(acc);
        }
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

// column-cell table of a pyramid on rows of w pixels; false when it has more than kMaxColCells column cells
static bool make_cols(const PpmScales &sc, int w, PpmCols *out) {
    int n = 0;
    for (int k = 0; k < sc.n; ++k) n += sc.s[k];
    if (n > kMaxColCells) return false;
    memset(out, 0, sizeof(*out));
    out->n = n;
    int col = 0;
    for (int k = 0; k < sc.n; ++k)
        for (int j = 0; j < sc.s[k]; ++j, ++col) {
            out->k[col] = k;
            out->j[col] = j;
            out->lo[col] = (j * w) / sc.s[k];
            out->hi[col] = ((j + 1) * w + sc.s[k] - 1) / sc.s[k];
        }
    return true;
}

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

// dfeat = addend + pool_backward(dpooled); addend bf16 [b][h][w][c] or NULL (the gradient the feature map's other consumers sent:
// the pooling is then the one place where the feature map's gradient is assembled, no separate add kernels)
extern "C" int regda_ppm_pool_bwd_add(const float *dpooled, const void *addend, void *dfeat, int b, int h, int w, int c,
                                      const int *scales_host, int nscales, void *stream) {
    PpmScales sc;
    const int rc = make_scales(scales_host, nscales, &sc);
    if (rc) return rc;
    if (!dpooled || !dfeat || b < 1 || h < 1 || w < 1 || c < 8 || c % 8) return fail(REGDA_ERR_INVALID_ARG, "ppm_pool_bwd: bad arguments");
    if ((reinterpret_cast<uintptr_t>(dpooled) | reinterpret_cast<uintptr_t>(dfeat) | reinterpret_cast<uintptr_t>(addend)) & 15)
        return fail(REGDA_ERR_INVALID_ARG, "ppm_pool_bwd: tensors must be 16-byte aligned");
    PpmCols cols;
    if (!make_cols(sc, w, &cols)) return fail(REGDA_ERR_UNSUPPORTED, "ppm_pool_bwd: pool scales sum to more than 16");
    if (w > kPoolMaxW) return fail(REGDA_ERR_UNSUPPORTED, "ppm_pool_bwd: feature-map rows wider than 512 pixels");
    ppm_pool_bwd_cols_kernel<<<dim3(h, b, (c + kPoolBwdChunk - 1) / kPoolBwdChunk), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dpooled, static_cast<const __nv_bfloat16 *>(addend), static_cast<__nv_bfloat16 *>(dfeat), h, w, c, sc, cols);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

extern "C" int regda_ppm_pool_bwd(const float *dpooled, void *dfeat, int b, int h, int w, int c, const int *scales_host, int nscales,
                                  void *stream) {
    return regda_ppm_pool_bwd_add(dpooled, nullptr, dfeat, b, h, w, c, scales_host, nscales, stream);
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
// 3x3 convolution over the 2048 feature channels as its epilogue addend.  Every GEMM reads (and differentiates) its part of the
// OHWI weight IN PLACE through the convolution kernels' weight channel stride; the kernels below are the remaining layout glue:
//   regda_ppm_g_pack / unpack  G_k [(img, cell)][(o, tap)]  <->  GT [img][o][(cell, tap)] (K-major operand of the second GEMM)
//   regda_ppm_cells            pooled float32 cells <-> the bf16 s x s branch inputs
//   regda_transpose_bf16       batched [r][c] -> [c][r]
// =========================================================================================================
namespace regda {
namespace {

struct GPtrs { __nv_bfloat16 *p[kMaxScales]; };

// GT[img][o][kappa], kappa = (cell_global * T + tap) < ncell * T, zero up to kp  <->  G_k[(img, cell)][o * T + tap] (the branch
// GEMMs run against the OHWI weight in place, so their output columns come in the weight's (o, tap) row order).
// One block per (image, 32 output channels): for every cell the 32 * T source elements are ONE contiguous run; they pass through a
// [32][kp] shared-memory tile whose rows are the dense GT rows (16-byte accesses on the GT side).
// PACK: G -> GT;  !PACK: GT -> G (the gradient's way back).  grid (ceil(O/32), b), dynamic shared memory 32 * kp * 2 bytes.
template <bool PACK>
__global__ void __launch_bounds__(256)
ppm_g_pack_kernel(const GPtrs g, __nv_bfloat16 *__restrict__ gt, int O, int T, int kp, const PpmScales sc) {
    extern __shared__ __align__(16) unsigned short gtile[];        // [32][kp]
    const int o0 = blockIdx.x * 32, img = blockIdx.y;
    const int no = min(32, O - o0);
    const int nk = sc.ncell * T, run = no * T;
    unsigned short *gtu = reinterpret_cast<unsigned short *>(gt) + (static_cast<size_t>(img) * O + o0) * kp;
    const int row_q = kp >> 3;                                     // 16-byte chunks per GT row (kp is a multiple of 8)
    if (PACK) {
        for (int i = threadIdx.x; i < no * (kp - nk); i += blockDim.x) {
            const int oo = i / (kp - nk), kappa = nk + i - oo * (kp - nk);
            gtile[oo * kp + kappa] = 0;
        }
    } else {
        for (int i = threadIdx.x; i < no * row_q; i += blockDim.x) {
            const int oo = i / row_q, q = i - oo * row_q;
            *reinterpret_cast<uint4 *>(gtile + oo * kp + q * 8) = *reinterpret_cast<const uint4 *>(gtu + static_cast<size_t>(oo) * kp + q * 8);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < sc.ncell * run; i += blockDim.x) {
        const int cell = i / run, e = i - cell * run;
        const int oo = e / T, tap = e - oo * T;
        int k = 0;
#pragma unroll
        for (int q = 1; q < kMaxScales; ++q) if (q < sc.n && cell >= sc.off[q]) k = q;
        const int j = cell - sc.off[k], s2 = sc.s[k] * sc.s[k];
        unsigned short *src = reinterpret_cast<unsigned short *>(g.p[k]) + (static_cast<size_t>(img) * s2 + j) * (static_cast<size_t>(T) * O) +
                              static_cast<size_t>(o0) * T + e;
        if (PACK) gtile[oo * kp + cell * T + tap] = *src;
        else *src = gtile[oo * kp + cell * T + tap];
    }
    if (PACK) {
        __syncthreads();
        for (int i = threadIdx.x; i < no * row_q; i += blockDim.x) {
            const int oo = i / row_q, q = i - oo * row_q;
            *reinterpret_cast<uint4 *>(gtu + static_cast<size_t>(oo) * kp + q * 8) = *reinterpret_cast<const uint4 *>(gtile + oo * kp + q * 8);
        }
    }
}

// pooled f32 [b][ncell][c] <-> the per-scale branch inputs p_k bf16 [b][s_k][s_k][c] (dense NHWC, what the branch convolutions read).
// SPLIT: p_k[img][j][:] = bf16(pooled[img][off_k + j][:]);   !SPLIT (the gradient's way back): dpooled[img][off_k + j][:] = dp_k[img][j][:]
// grid (ncell, b), a thread moves 8 channels.
template <bool SPLIT>
__global__ void __launch_bounds__(256)
ppm_cells_kernel(float *__restrict__ pooled, const GPtrs p, int c, const PpmScales sc) {
    const int cell = blockIdx.x, img = blockIdx.y;
    int k = 0;
#pragma unroll
    for (int i = 1; i < kMaxScales; ++i) if (i < sc.n && cell >= sc.off[i]) k = i;
    const int j = cell - sc.off[k], s2 = sc.s[k] * sc.s[k];
    float *src = pooled + (static_cast<size_t>(img) * sc.ncell + cell) * c;
    __nv_bfloat16 *dst = p.p[k] + (static_cast<size_t>(img) * s2 + j) * c;
    for (int ch = threadIdx.x * 8; ch < c; ch += blockDim.x * 8) {
        if (SPLIT) {
            const float4 a = *reinterpret_cast<const float4 *>(src + ch), b2 = *reinterpret_cast<const float4 *>(src + ch + 4);
            const float f[8] = {a.x, a.y, a.z, a.w, b2.x, b2.y, b2.z, b2.w};
            *reinterpret_cast<bf16x8 *>(dst + ch) = pack(f);
        } else {
            float f[8];
            unpack(*reinterpret_cast<const bf16x8 *>(dst + ch), f);
            *reinterpret_cast<float4 *>(src + ch) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4 *>(src + ch + 4) = make_float4(f[4], f[5], f[6], f[7]);
        }
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

// pack != 0: G_k bf16 [b][s_k*s_k][O*T] (column = o*T + tap) -> GT bf16 [b][O][kp] (kappa = cell*T + tap, zero padded);  pack == 0: the reverse
extern "C" int regda_ppm_g_pack(void *g0, void *g1, void *g2, void *g3, void *gt, int b, int O, int T, int kp, const int *scales_host,
                                int nscales, int pack, void *stream) {
    PpmScales sc;
    const int rc = make_scales(scales_host, nscales, &sc);
    if (rc) return rc;
    if (!gt || b < 1 || b > 65535 || O < 1 || O > 65535 * 32 || T < 1 || kp < sc.ncell * T) return fail(REGDA_ERR_INVALID_ARG, "ppm_g_pack: bad arguments");
    GPtrs p;
    void *ps[4] = {g0, g1, g2, g3};
    for (int k = 0; k < kMaxScales; ++k) {
        p.p[k] = static_cast<__nv_bfloat16 *>(ps[k]);
        if (k < nscales && !ps[k]) return fail(REGDA_ERR_INVALID_ARG, "ppm_g_pack: null branch pointer");
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (kp % 8 || (reinterpret_cast<uintptr_t>(gt) & 15)) return fail(REGDA_ERR_INVALID_ARG, "ppm_g_pack: kp must be a multiple of 8 and GT 16-byte aligned");
    const size_t smem = static_cast<size_t>(32) * kp * 2;
    if (smem > 160 * 1024) return fail(REGDA_ERR_UNSUPPORTED, "ppm_g_pack: kp too large for the shared-memory tile");
    const dim3 grid((O + 31) / 32, b);
    if (pack) {
        if (smem > 48 * 1024) REGDA_CUDA_CHECK(cudaFuncSetAttribute(ppm_g_pack_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        ppm_g_pack_kernel<true><<<grid, 256, smem, st>>>(p, static_cast<__nv_bfloat16 *>(gt), O, T, kp, sc);
    } else {
        if (smem > 48 * 1024) REGDA_CUDA_CHECK(cudaFuncSetAttribute(ppm_g_pack_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        ppm_g_pack_kernel<false><<<grid, 256, smem, st>>>(p, static_cast<__nv_bfloat16 *>(gt), O, T, kp, sc);
    }
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

// split != 0: pooled float32 [b][ncell][c] -> p_k bf16 [b][s_k][s_k][c] for every scale (the branch convolutions' inputs);
// split == 0: the reverse with a float32 result (the gradient of the pooled cells from the branch inputs' gradients)
extern "C" int regda_ppm_cells(float *pooled, void *p0, void *p1, void *p2, void *p3, int b, int c, const int *scales_host, int nscales,
                               int split, void *stream) {
    PpmScales sc;
    const int rc = make_scales(scales_host, nscales, &sc);
    if (rc) return rc;
    if (!pooled || b < 1 || b > 65535 || c < 8 || c % 8) return fail(REGDA_ERR_INVALID_ARG, "ppm_cells: bad arguments");
    GPtrs p;
    void *ps[4] = {p0, p1, p2, p3};
    for (int k = 0; k < kMaxScales; ++k) {
        p.p[k] = static_cast<__nv_bfloat16 *>(ps[k]);
        if (k < nscales && (!ps[k] || (reinterpret_cast<uintptr_t>(ps[k]) & 15))) return fail(REGDA_ERR_INVALID_ARG, "ppm_cells: null / unaligned branch pointer");
    }
    if (reinterpret_cast<uintptr_t>(pooled) & 15) return fail(REGDA_ERR_INVALID_ARG, "ppm_cells: pooled must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int threads = std::min(256, c / 8);
    if (split) ppm_cells_kernel<true><<<dim3(sc.ncell, b), threads, 0, st>>>(pooled, p, c, sc);
    else ppm_cells_kernel<false><<<dim3(sc.ncell, b), threads, 0, st>>>(pooled, p, c, sc);
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
