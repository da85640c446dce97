// Stem convolution support (regda/_resnets.py:150-153: Conv2d(3, 64, 7, stride 2, padding 3)).
//
// Three input channels cannot feed the implicit-GEMM kernel directly (its K-blocks are 64 channels of one filter tap), so
// the stem is lowered to a GEMM explicitly: this kernel writes the patch matrix A [n*oh*ow][192] (k = r*24 + s*3 + c: every
// filter row r padded from 21 to 24 taps, zero up to 192 = 3 K-blocks) and the tcgen05 kernels then run it as a 1x1 convolution with 192 input
// channels -- forward (with the fused BatchNorm statistics) and weight gradient; the image needs no data gradient.
// HBM-bound: 384 B written per output pixel; the 3-channel image is read through L1/L2 (each element is used ~12 times).
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"

namespace regda {
namespace {

constexpr int kStemK = 192;

constexpr int kStemSeg = 64;                       // output pixels per block
constexpr int kStemRow = (2 * kStemSeg + 5) * 3 + 1; // staged input elements per filter row: 133 pixels x 3 channels (+1 pad)

// One block = 64 consecutive output pixels of one output row: the 7 input rows x 133 input pixels they read are staged in
// shared memory with coalesced loads (zero outside the image), then every thread emits 16-byte chunks of the patch matrix
// (fully coalesced 384-byte rows); tap (r, j) of pixel p is staged element r * kStemRow + j + 6*p.
// F32_NCHW: the image is the float32 [n][3][h][w] tensor the data loader delivers (regda/datasets: CHW float images); it is
// rounded to bf16 while it is staged, so the separate NCHW float32 -> NHWC bf16 conversion pass over the batch disappears.
template <bool F32_NCHW>
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const void *__restrict__ xv, __nv_bfloat16 *__restrict__ a, int n, int h, int w, int oh, int ow, int segs) {
    __shared__ __align__(16) unsigned short sin[7 * kStemRow + 8];      // (+8: the masked tail of the last chunk is still read)
    int b = blockIdx.x;
    const int seg = b % segs; b /= segs;
    const int oy = b % oh;
    const int img = b / oh;
    const int ox0 = seg * kStemSeg;
    const int ix_start = 2 * ox0 - 3;
    if (F32_NCHW) {
        // one plane row (r, ch) at a time, 16 bytes per load: the run starts one pixel left of the first staged pixel, where the row is
        // 16-byte aligned (ix_start - 1 = 2 * (ox0 - 2), ox0 a multiple of 64, w a multiple of 4); element (r, ch, px) -> staged slot
        // r*kStemRow + px*3 + ch
        const float *xi = static_cast<const float *>(xv) + static_cast<long long>(img) * 3 * h * w;
        constexpr int kPx = (kStemRow - 1) / 3;                       // 133 staged pixels per row
        constexpr int kQuads = (kPx + 1 + 3) / 4;                     // float4 loads per plane row (pixels -1 .. 4*kQuads-2)
        const bool vec_ok = (w & 3) == 0 && (reinterpret_cast<uintptr_t>(xv) & 15) == 0;
        for (int e = threadIdx.x; e < 7 * 3 * kQuads; e += 256) {
            const int rc = e / kQuads, qd = e - rc * kQuads;
            const int r = rc / 3, ch = rc - r * 3;
            const int iy = 2 * oy - 3 + r, ix0 = ix_start - 1 + 4 * qd;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (iy >= 0 && iy < h) {
                const float *rowp = xi + (static_cast<long long>(ch) * h + iy) * w;
                if (vec_ok && ix0 >= 0 && ix0 + 3 < w) {
                    const float4 f = __ldg(reinterpret_cast<const float4 *>(rowp + ix0));
                    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (ix0 + t >= 0 && ix0 + t < w) v[t] = __ldg(rowp + ix0 + t);
                }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int px = 4 * qd - 1 + t;
                if (px >= 0 && px < kPx) sin[r * kStemRow + px * 3 + ch] = __bfloat16_as_ushort(__float2bfloat16_rn(v[t]));
            }
        }
        if (threadIdx.x < 7) sin[threadIdx.x * kStemRow + kStemRow - 1] = 0;
    } else {
        const unsigned short *xi = static_cast<const unsigned short *>(xv) + static_cast<long long>(img) * h * w * 3;
        for (int e = threadIdx.x; e < 7 * kStemRow; e += 256) {
            const int r = e / kStemRow, c = e - r * kStemRow;
            const int iy = 2 * oy - 3 + r;
            const int px = c / 3;
            const int ix = ix_start + px;
            unsigned short v = 0;
            if (c < kStemRow - 1 && iy >= 0 && iy < h && ix >= 0 && ix < w) v = __ldg(xi + (static_cast<long long>(iy) * w + ix_start) * 3 + c);
            sin[e] = v;
        }
    }
    __syncthreads();
    // Patch row = 8 groups of 24 taps: group r < 7 holds filter row r as (s, c) = 21 real taps + 3 zeros, group 7 is zero.  A
    // 16-byte chunk is 8 consecutive staged elements of one filter row (4 aligned 32-bit shared loads: every index below is even).
    const int npx = min(kStemSeg, ow - ox0);
    __nv_bfloat16 *arow = a + ((static_cast<long long>(img) * oh + oy) * ow + ox0) * kStemK;
    for (int q = threadIdx.x; q < npx * (kStemK / 8); q += 256) {
        const int p = q / (kStemK / 8), c = q - p * (kStemK / 8);
        const int r = c / 3, part = c - r * 3;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (r < 7) {
            const unsigned *src = reinterpret_cast<const unsigned *>(sin + r * kStemRow + part * 8 + 6 * p);
            o.x = src[0]; o.y = src[1];
            if (part < 2) { o.z = src[2]; o.w = src[3]; }
            else o.z = src[2] & 0xffffu;                       // taps 16..20 of the row; 21..23 are padding
        }
        *reinterpret_cast<uint4 *>(arow + static_cast<long long>(q) * 8) = o;
    }
}

}  // namespace
}  // namespace regda

using namespace regda;

static int stem_im2col(const void *x, bool f32_nchw, void *a, int n, int h, int w, void *stream) {
    if (!x || !a || n < 1 || h < 1 || w < 1) return fail(REGDA_ERR_INVALID_ARG, "stem_im2col: bad arguments");
    if (reinterpret_cast<uintptr_t>(a) & 15) return fail(REGDA_ERR_INVALID_ARG, "stem_im2col: output must be 16-byte aligned");
    const int oh = (h - 1) / 2 + 1, ow = (w - 1) / 2 + 1;
    const int segs = (ow + kStemSeg - 1) / kStemSeg;
    const long long blocks = static_cast<long long>(n) * oh * segs;
    if (blocks > 0x7fffffffll) return fail(REGDA_ERR_UNSUPPORTED, "stem_im2col: too many output rows for one launch");
    if (f32_nchw)
        stem_im2col_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<__nv_bfloat16 *>(a), n, h, w, oh, ow, segs);
    else
        stem_im2col_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<__nv_bfloat16 *>(a), n, h, w, oh, ow, segs);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

// x bf16 [n][h][w][3] (channels-last image), a bf16 [n][oh][ow][192] with oh = (h-1)/2+1, ow = (w-1)/2+1
extern "C" int regda_stem_im2col_bf16(const void *x, void *a, int n, int h, int w, void *stream) {
    return stem_im2col(x, false, a, n, h, w, stream);
}

// x float32 [n][3][h][w] (the loader's NCHW image), rounded to bf16 on the way: same patch matrix as the bf16 form of bf16(x)
extern "C" int regda_stem_im2col_f32nchw(const float *x, void *a, int n, int h, int w, void *stream) {
    return stem_im2col(x, true, a, n, h, w, stream);
}
