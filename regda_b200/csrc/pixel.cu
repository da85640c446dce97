// Full-resolution per-pixel chain of the self-training step (HBM-bound, ~0 flop/byte):
//   Aligner.label_refine  (reference regda/gast/alignment.py:194-265, :283-298, :396-423)
//   pseudo_selection      (reference regda/gast/pseudo_generation.py:59-93)
// and the fused form refine -> select that never materialises the refined [b,c,H,W] tensor.
//
// Data layout: soft / pred / simi are [b][c][plane] float32 exactly as the reference's NCHW
// tensors; the feature map is consumed as channels-last rows [b*h*w][k].
#include <cuda_bf16.h>

#include "common.cuh"

namespace regda {
namespace {

constexpr int kPxThreads = 256;
constexpr float kEps = 1e-7f;   // Aligner.eps, alignment.py:43

// ----------------------------------------------------------------------------------------
// Pearson distance between feature rows and prototypes (alignment.py:396-423)
// ----------------------------------------------------------------------------------------
// centred prototypes pc[c][k] and their unbiased std: one block per prototype
__global__ void __launch_bounds__(256)
proto_stats_kernel(const float *__restrict__ proto, float *__restrict__ pc, float *__restrict__ pstd, int k) {
    __shared__ float red[8];
    __shared__ float bcast;
    const int c = blockIdx.x;
    const float *p = proto + static_cast<size_t>(c) * k;
    float s = 0.f;
    for (int i = threadIdx.x; i < k; i += blockDim.x) s += p[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        bcast = t / static_cast<float>(k);
    }
    __syncthreads();
    const float mean = bcast;
    float q = 0.f;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const float d = p[i] - mean;
        pc[static_cast<size_t>(c) * k + i] = d;
        q += d * d;
    }
    q = warp_sum(q);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        pstd[c] = sqrtf(t / static_cast<float>(k - 1));
    }
}

// feature rows are float32 (the reference's feat) or bf16 (the layout the model's InstanceNorm kernel writes: the trainer
// hands it over without the 200 MB float32 copy); elements 4i .. 4i+3 of a row:
__device__ __forceinline__ float4 load_row4(const float *x, int i) { return reinterpret_cast<const float4 *>(x)[i]; }
__device__ __forceinline__ float4 load_row4(const __nv_bfloat16 *x, int i) {
    const uint2 u = reinterpret_cast<const uint2 *>(x)[i];
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
}
__device__ __forceinline__ float load_row1(const float *x, int i) { return x[i]; }
__device__ __forceinline__ float load_row1(const __nv_bfloat16 *x, int i) { return __bfloat162float(x[i]); }

// one warp per feature row.  out_mode 0: dist[n][c];  1: 1/dist written as planes [b][c][hw]
template <int CMAX, typename T>
__global__ void __launch_bounds__(256)
pearson_kernel(const T *__restrict__ rows, const float *__restrict__ pc, const float *__restrict__ pstd,
               float *__restrict__ out, long long n, int c, int k, int hw, int out_mode, float kdiv) {
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const T *x = rows + row * k;
    const bool vec = (k % 4 == 0) && ((reinterpret_cast<uintptr_t>(rows) & 15) == 0);
    float s = 0.f;
    if (vec) {
        for (int i = lane; i < k / 4; i += 32) { const float4 v = load_row4(x, i); s += (v.x + v.y) + (v.z + v.w); }
    } else {
        for (int i = lane; i < k; i += 32) s += load_row1(x, i);
    }
    const float mean = warp_sum(s) / static_cast<float>(k);
    float q = 0.f, dot[CMAX];
#pragma unroll
    for (int j = 0; j < CMAX; ++j) dot[j] = 0.f;
    if (vec) {
        for (int i = lane; i < k / 4; i += 32) {
            float4 v = load_row4(x, i);
            v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
            q += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
            for (int j = 0; j < CMAX; ++j)
                if (j < c) {
                    const float4 p = reinterpret_cast<const float4 *>(pc + static_cast<size_t>(j) * k)[i];
                    dot[j] += v.x * p.x + v.y * p.y + v.z * p.z + v.w * p.w;
                }
        }
    } else {
        for (int i = lane; i < k; i += 32) {
            const float d = load_row1(x, i) - mean;
            q += d * d;
#pragma unroll
            for (int j = 0; j < CMAX; ++j)
                if (j < c) dot[j] += d * pc[static_cast<size_t>(j) * k + i];
        }
    }
    q = warp_sum(q);
#pragma unroll
    for (int j = 0; j < CMAX; ++j) dot[j] = warp_sum(dot[j]);
    const float sx = sqrtf(q / static_cast<float>(k - 1));            // unbiased std (:414)
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < CMAX; ++j)
            if (j < c) {
                const float cov = dot[j] / kdiv;                       // (:410) k - 1 + eps
                const float div = sx * pstd[j] + kEps;                 // (:418-419)
                const float dist = (-1.0f * cov / div + 1.0f) * 0.5f;  // (:419)
                if (out_mode == 0) out[row * c + j] = dist;
                else {
                    const long long img = row / hw, pos = row % hw;
                    out[(img * c + j) * hw + pos] = 1.0f / dist;       // (:216)
                }
            }
    }
}

// ----------------------------------------------------------------------------------------
// per-pixel refinement weights
// ----------------------------------------------------------------------------------------
struct RefineArgs {
    const float *simi;    // [b][c][h][w] = 1 / pearson_dist
    const float *pred1;   // [b][c][h][w]
    const float *pred2;   // may be null
    const float *soft;    // [b][c][H*W]
    int c, h, w, H, W;
    float inv_temp, sy, sx;   // 1 / temperature; align_corners=True scales (in-1)/(out-1)
};

struct Taps {
    int o00, o01, o10, o11;
    float wy0, wy1, wx0, wx1;
};

// upsample_bilinear2d, align_corners=True: src = scale*dst; i0 = (int)src; lambda1 = src - i0
__device__ __forceinline__ Taps make_taps(int Y, int X, const RefineArgs &a) {
    const float fy = a.sy * static_cast<float>(Y), fx = a.sx * static_cast<float>(X);
    const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    const int y1 = y0 + (y0 < a.h - 1 ? 1 : 0), x1 = x0 + (x0 < a.w - 1 ? 1 : 0);
    Taps t;
    t.wy1 = fy - static_cast<float>(y0); t.wy0 = 1.0f - t.wy1;
    t.wx1 = fx - static_cast<float>(x0); t.wx0 = 1.0f - t.wx1;
    t.o00 = y0 * a.w + x0; t.o01 = y0 * a.w + x1; t.o10 = y1 * a.w + x0; t.o11 = y1 * a.w + x1;
    return t;
}

__device__ __forceinline__ float interp(const float *__restrict__ plane, const Taps &t) {
    return t.wy0 * (t.wx0 * __ldg(plane + t.o00) + t.wx1 * __ldg(plane + t.o01)) +
           t.wy1 * (t.wx0 * __ldg(plane + t.o10) + t.wx1 * __ldg(plane + t.o11));
}

template <int CMAX>
__device__ __forceinline__ void softmax_inplace(float (&v)[CMAX], int c) {
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < CMAX; ++j) if (j < c) m = fmaxf(m, v[j]);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < CMAX; ++j) if (j < c) { v[j] = expf(v[j] - m); s += v[j]; }
    // one correctly rounded reciprocal + c multiplications (<= 1.5 ulp from the quotient; the refinement is compared at 1e-4)
    const float inv = __frcp_rn(s);
#pragma unroll
    for (int j = 0; j < CMAX; ++j) if (j < c) v[j] = v[j] * inv;
}

template <int CMAX>
__device__ __forceinline__ void peak_normalise(float (&v)[CMAX], int c) {
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < CMAX; ++j) if (j < c) m = fmaxf(m, v[j]);
    const float inv = __frcp_rn(m + kEps);
#pragma unroll
    for (int j = 0; j < CMAX; ++j) if (j < c) v[j] = v[j] * inv;
}

// refined[c] for pixel (img, pos) -- alignment.py:216-236, 263-264
template <int CMAX>
__device__ __forceinline__ void refine_pixel(const RefineArgs &a, int img, int pos, float (&out)[CMAX]) {
    const int Y = pos / a.W, X = pos - Y * a.W;
    const Taps t = make_taps(Y, X, a);
    const int lp = a.h * a.w;
    const size_t base = static_cast<size_t>(img) * a.c * lp;
    float wgt[CMAX], p1[CMAX];
#pragma unroll
    for (int j = 0; j < CMAX; ++j) if (j < a.c) wgt[j] = interp(a.simi + base + static_cast<size_t>(j) * lp, t);
    softmax_inplace<CMAX>(wgt, a.c);                 // _softmax_T(temp=1) (:220)
    peak_normalise<CMAX>(wgt, a.c);                  // (:221-222)
#pragma unroll
    for (int j = 0; j < CMAX; ++j) if (j < a.c) p1[j] = interp(a.pred1 + base + static_cast<size_t>(j) * lp, t) * a.inv_temp;
    softmax_inplace<CMAX>(p1, a.c);
    if (a.pred2 != nullptr) {
        float p2[CMAX];
#pragma unroll
        for (int j = 0; j < CMAX; ++j) if (j < a.c) p2[j] = interp(a.pred2 + base + static_cast<size_t>(j) * lp, t) * a.inv_temp;
        softmax_inplace<CMAX>(p2, a.c);
#pragma unroll
        for (int j = 0; j < CMAX; ++j) if (j < a.c) p1[j] = (p1[j] + p2[j]) * 0.5f;   // (:230-231)
    }
    peak_normalise<CMAX>(p1, a.c);                   // (:235)
    const size_t HW = static_cast<size_t>(a.H) * a.W;
    const float *s = a.soft + static_cast<size_t>(img) * a.c * HW + pos;
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < CMAX; ++j)
        if (j < a.c) {
            out[j] = (wgt[j] + p1[j]) * __ldg(s + static_cast<size_t>(j) * HW);      // (:263)
            sum += out[j];
        }
    const float inv = __frcp_rn(sum + kEps);         // _logits_norm (:295-297)
#pragma unroll
    for (int j = 0; j < CMAX; ++j) if (j < a.c) out[j] = out[j] * inv;
}

// block-wide per-class max -> one atomicMax per class per block (values are >= 0: the int
// ordering of the bit patterns equals the float ordering)
template <int CMAX>
__device__ __forceinline__ void block_class_max(float (&v)[CMAX], int c, unsigned *gmax) {
    __shared__ float red[CMAX][kPxThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < CMAX; ++j)
        if (j < c) {
            const float m = warp_max(v[j]);
            if (lane == 0) red[j][wid] = m;
        }
    __syncthreads();
    if (threadIdx.x < c) {
        float m = red[threadIdx.x][0];
        for (int i = 1; i < kPxThreads / 32; ++i) m = fmaxf(m, red[threadIdx.x][i]);
        atomicMax(gmax + threadIdx.x, __float_as_uint(m));
    }
}

// Grid-stride over the pixels of one image (blockIdx.y): a thread keeps the running per-class maximum of the pixels it visits, so
// the block-wide reduction and the c atomicMax per block happen once per block, not once per 256 pixels (8192 blocks x 6
// same-address atomics per image serialise in L2).
template <int CMAX, bool WRITE>
__global__ void __launch_bounds__(kPxThreads, CMAX <= 8 ? 3 : 1)
refine_max_kernel(const RefineArgs a, float *__restrict__ soft_out, unsigned *__restrict__ gmax) {
    const int img = blockIdx.y;
    const int HW = a.H * a.W;
    float m[CMAX];
#pragma unroll
    for (int j = 0; j < CMAX; ++j) m[j] = 0.f;
    for (int pos = blockIdx.x * kPxThreads + threadIdx.x; pos < HW; pos += gridDim.x * kPxThreads) {
        float r[CMAX];
#pragma unroll
        for (int j = 0; j < CMAX; ++j) r[j] = 0.f;
        refine_pixel<CMAX>(a, img, pos, r);
        if (WRITE) {
            float *o = soft_out + static_cast<size_t>(img) * a.c * HW + pos;
#pragma unroll
            for (int j = 0; j < CMAX; ++j) if (j < a.c) o[static_cast<size_t>(j) * HW] = r[j];
        }
#pragma unroll
        for (int j = 0; j < CMAX; ++j) m[j] = fmaxf(m[j], r[j]);
    }
    if (gmax != nullptr) block_class_max<CMAX>(m, a.c, gmax + img * a.c);
}

// pseudo_generation.py:76-88 for one pixel whose c probabilities are in v[]
template <int CMAX>
__device__ __forceinline__ long long select_pixel(const float (&v)[CMAX], int c, const unsigned *gmax, float top, float low, long long ignore_label) {
    int n_pass = 0, which = 0;
#pragma unroll
    for (int j = 0; j < CMAX; ++j)
        if (j < c) {
            const float thr = fmaxf(__fmul_rn(__uint_as_float(gmax[j]), top), low);   // (:77, :80-81)
            if (v[j] > thr) { if (n_pass == 0) which = j; ++n_pass; }                 // strict > (:83)
        }
    return n_pass == 1 ? static_cast<long long>(which) : ignore_label;                // (:85-88)
}

template <int CMAX>
__global__ void __launch_bounds__(kPxThreads)
refine_select_kernel(const RefineArgs a, const unsigned *__restrict__ gmax, long long *__restrict__ out,
                     float top, float low, long long ignore_label) {
    const int img = blockIdx.y;
    const int HW = a.H * a.W;
    const int pos = blockIdx.x * kPxThreads + threadIdx.x;
    if (pos >= HW) return;
    float r[CMAX];
    refine_pixel<CMAX>(a, img, pos, r);
    out[static_cast<size_t>(img) * HW + pos] = select_pixel<CMAX>(r, a.c, gmax + img * a.c, top, low, ignore_label);
}

// standalone pseudo_selection on a given soft tensor
template <int CMAX>
__global__ void __launch_bounds__(kPxThreads)
soft_max_kernel(const float *__restrict__ soft, int c, long long HW, unsigned *__restrict__ gmax, int32_t *flags) {
    const int img = blockIdx.y;
    const float *s = soft + static_cast<size_t>(img) * c * HW;
    float m[CMAX];
    bool bad = false;
#pragma unroll
    for (int j = 0; j < CMAX; ++j) m[j] = 0.f;
    for (long long pos = static_cast<long long>(blockIdx.x) * kPxThreads + threadIdx.x; pos < HW; pos += static_cast<long long>(gridDim.x) * kPxThreads) {
#pragma unroll
        for (int j = 0; j < CMAX; ++j)
            if (j < c) {
                const float v = __ldg(s + static_cast<size_t>(j) * HW + pos);
                bad |= !(v <= 1.0f) || !(v >= 0.0f);           // the reference asserts 0 <= mask <= 1 (:71)
                m[j] = fmaxf(m[j], v);
            }
    }
    if (bad) raise_flag(flags, REGDA_FLAG_PROB_RANGE);
    block_class_max<CMAX>(m, c, gmax + img * c);
}

template <int CMAX>
__global__ void __launch_bounds__(kPxThreads)
soft_select_kernel(const float *__restrict__ soft, int c, long long HW, const unsigned *__restrict__ gmax,
                   long long *__restrict__ out, float top, float low, long long ignore_label) {
    const int img = blockIdx.y;
    const float *s = soft + static_cast<size_t>(img) * c * HW;
    for (long long pos = static_cast<long long>(blockIdx.x) * kPxThreads + threadIdx.x; pos < HW; pos += static_cast<long long>(gridDim.x) * kPxThreads) {
        float v[CMAX];
#pragma unroll
        for (int j = 0; j < CMAX; ++j) v[j] = (j < c) ? __ldg(s + static_cast<size_t>(j) * HW + pos) : 0.f;
        out[static_cast<size_t>(img) * HW + pos] = select_pixel<CMAX>(v, c, gmax + img * c, top, low, ignore_label);
    }
}

struct RefineWs {
    float *pc, *pstd, *simi;
    unsigned *gmax;
    size_t bytes;
};

RefineWs carve_refine_ws(void *ws, int b, int c, int k, int h, int w) {
    RefineWs r;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += align_up(n, 256); return o; };
    const size_t o_pc = take(static_cast<size_t>(c) * k * 4), o_std = take(static_cast<size_t>(c) * 4);
    const size_t o_simi = take(static_cast<size_t>(b) * c * h * w * 4), o_max = take(static_cast<size_t>(b) * c * 4);
    char *p = static_cast<char *>(ws);
    r.pc = reinterpret_cast<float *>(p + o_pc); r.pstd = reinterpret_cast<float *>(p + o_std);
    r.simi = reinterpret_cast<float *>(p + o_simi); r.gmax = reinterpret_cast<unsigned *>(p + o_max);
    r.bytes = off;
    return r;
}

// rows: float32, or bf16 when rows_bf16
int launch_pearson(const void *rows, bool rows_bf16, const float *protos, float *pc, float *pstd, float *out, long long n, int c, int k,
                   int hw, int out_mode, cudaStream_t st) {
    proto_stats_kernel<<<c, 256, 0, st>>>(protos, pc, pstd, k);
    REGDA_LAUNCH_CHECK();
    const float kdiv = static_cast<float>(static_cast<double>(k - 1) + 1e-7);
    const unsigned blocks = static_cast<unsigned>((n + 7) / 8);
    const float *rf = static_cast<const float *>(rows);
    const __nv_bfloat16 *rb = static_cast<const __nv_bfloat16 *>(rows);
    if (rows_bf16) {
        if (c <= 8) pearson_kernel<8><<<blocks, 256, 0, st>>>(rb, pc, pstd, out, n, c, k, hw, out_mode, kdiv);
        else pearson_kernel<16><<<blocks, 256, 0, st>>>(rb, pc, pstd, out, n, c, k, hw, out_mode, kdiv);
    } else {
        if (c <= 8) pearson_kernel<8><<<blocks, 256, 0, st>>>(rf, pc, pstd, out, n, c, k, hw, out_mode, kdiv);
        else pearson_kernel<16><<<blocks, 256, 0, st>>>(rf, pc, pstd, out, n, c, k, hw, out_mode, kdiv);
    }
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

int check_refine_args(int b, int c, int k, int h, int w, int H, int W, double temp) {
    if (b < 0 || c < 1 || k < 2 || h < 1 || w < 1 || H < 1 || W < 1) return fail(REGDA_ERR_INVALID_ARG, "label_refine: bad shape");
    if (c > 16) return fail(REGDA_ERR_UNSUPPORTED, "label_refine: at most 16 classes");
    if (!(temp > 0)) return fail(REGDA_ERR_INVALID_ARG, "label_refine: temp must be > 0 (alignment.py:285)");
    if (static_cast<long long>(H) * W > 0x7fffffffll) return fail(REGDA_ERR_UNSUPPORTED, "label_refine: image too large");
    return REGDA_OK;
}

RefineArgs make_refine_args(const float *simi, const float *p1, const float *p2, const float *soft, int c, int h, int w, int H, int W, double temp) {
    RefineArgs a;
    a.simi = simi; a.pred1 = p1; a.pred2 = p2; a.soft = soft;
    a.c = c; a.h = h; a.w = w; a.H = H; a.W = W;
    a.inv_temp = static_cast<float>(1.0 / temp);
    a.sy = H > 1 ? static_cast<float>(h - 1) / static_cast<float>(H - 1) : 0.f;
    a.sx = W > 1 ? static_cast<float>(w - 1) / static_cast<float>(W - 1) : 0.f;
    return a;
}

}  // namespace
}  // namespace regda

using namespace regda;

extern "C" size_t regda_refine_workspace_bytes(int b, int c, int k, int h, int w) {
    if (b < 0 || c < 1 || k < 1 || h < 1 || w < 1) return 0;
    return carve_refine_ws(nullptr, b, c, k, h, w).bytes;
}

extern "C" size_t regda_pearson_workspace_bytes(int c, int k) {
    return align_up(static_cast<size_t>(c) * k * 4, 256) + align_up(static_cast<size_t>(c) * 4, 256);
}

extern "C" int regda_pearson_dist(const float *rows, const float *prototypes, float *dist, int64_t n, int c, int k,
                                  void *workspace, size_t workspace_bytes, void *stream) {
    if (n < 0 || c < 1 || k < 2) return fail(REGDA_ERR_INVALID_ARG, "pearson_dist: bad shape");
    if (c > 16) return fail(REGDA_ERR_UNSUPPORTED, "pearson_dist: at most 16 prototypes");
    if (n == 0) return REGDA_OK;
    if (!rows || !prototypes || !dist) return fail(REGDA_ERR_INVALID_ARG, "pearson_dist: null pointer");
    if (!workspace || workspace_bytes < regda_pearson_workspace_bytes(c, k)) return fail(REGDA_ERR_WORKSPACE, "pearson_dist: workspace too small");
    float *pc = static_cast<float *>(workspace);
    float *pstd = reinterpret_cast<float *>(static_cast<char *>(workspace) + align_up(static_cast<size_t>(c) * k * 4, 256));
    return launch_pearson(rows, false, prototypes, pc, pstd, dist, n, c, k, 1, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int regda_label_refine(const float *feat_nhwc, const float *prototypes, const float *pred1, const float *pred2,
                                  const float *soft_in, float *soft_out, int b, int c, int k, int h, int w, int H, int W,
                                  double temp, void *workspace, size_t workspace_bytes, void *stream) {
    int rc = check_refine_args(b, c, k, h, w, H, W, temp);
    if (rc) return rc;
    if (b == 0) return REGDA_OK;
    if (!feat_nhwc || !prototypes || !pred1 || !soft_in || !soft_out) return fail(REGDA_ERR_INVALID_ARG, "label_refine: null pointer");
    const RefineWs ws = carve_refine_ws(workspace, b, c, k, h, w);
    if (!workspace || workspace_bytes < ws.bytes) return fail(REGDA_ERR_WORKSPACE, "label_refine: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = launch_pearson(feat_nhwc, false, prototypes, ws.pc, ws.pstd, ws.simi, static_cast<long long>(b) * h * w, c, k, h * w, 1, st);
    if (rc) return rc;
    const RefineArgs a = make_refine_args(ws.simi, pred1, pred2, soft_in, c, h, w, H, W, temp);
    const dim3 grid((H * W + kPxThreads - 1) / kPxThreads, b);
    if (c <= 8) refine_max_kernel<8, true><<<grid, kPxThreads, 0, st>>>(a, soft_out, nullptr);
    else refine_max_kernel<16, true><<<grid, kPxThreads, 0, st>>>(a, soft_out, nullptr);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

static int refine_select_impl(const void *feat_nhwc, bool feat_bf16, const float *prototypes, const float *pred1, const float *pred2,
                              const float *soft_in, int64_t *hard_out, int b, int c, int k, int h, int w, int H, int W,
                              double temp, double cutoff_top, double cutoff_low, int64_t ignore_label,
                              void *workspace, size_t workspace_bytes, void *stream) {
    int rc = check_refine_args(b, c, k, h, w, H, W, temp);
    if (rc) return rc;
    if (b == 0) return REGDA_OK;
    if (!feat_nhwc || !prototypes || !pred1 || !soft_in || !hard_out) return fail(REGDA_ERR_INVALID_ARG, "refine_select: null pointer");
    const RefineWs ws = carve_refine_ws(workspace, b, c, k, h, w);
    if (!workspace || workspace_bytes < ws.bytes) return fail(REGDA_ERR_WORKSPACE, "refine_select: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = launch_pearson(feat_nhwc, feat_bf16, prototypes, ws.pc, ws.pstd, ws.simi, static_cast<long long>(b) * h * w, c, k, h * w, 1, st);
    if (rc) return rc;
    REGDA_CUDA_CHECK(cudaMemsetAsync(ws.gmax, 0, static_cast<size_t>(b) * c * 4, st));
    const RefineArgs a = make_refine_args(ws.simi, pred1, pred2, soft_in, c, h, w, H, W, temp);
    const dim3 grid((H * W + kPxThreads - 1) / kPxThreads, b);
    const float top = static_cast<float>(cutoff_top), low = static_cast<float>(cutoff_low);
    // max pass: ~8 blocks per SM over all images, each block walks its image with a grid stride
    const dim3 mgrid(std::min<int>(grid.x, std::max(1, 8 * sm_count() / b)), b);
    // The selection needs every image's per-class maximum of the REFINED probabilities first.  With room in the workspace for
    // them (regda_refine_select_workspace_bytes: b*c*H*W floats more) the max pass writes the refined probabilities and the
    // selection is a plain threshold pass over them (24 B/px written + read, L2-friendly) instead of a second evaluation of the
    // three soft-maxes and eighteen bilinear taps per pixel; without it the refinement is recomputed (same bits either way).
    const size_t cache_bytes = static_cast<size_t>(b) * c * H * W * sizeof(float);
    float *cache = workspace_bytes >= align_up(ws.bytes, 256) + cache_bytes ? reinterpret_cast<float *>(static_cast<char *>(workspace) + align_up(ws.bytes, 256)) : nullptr;
    const long long HW = static_cast<long long>(H) * W;
    long long *o = reinterpret_cast<long long *>(hard_out);
    if (c <= 8) {
        if (cache) {
            refine_max_kernel<8, true><<<mgrid, kPxThreads, 0, st>>>(a, cache, ws.gmax);
            REGDA_LAUNCH_CHECK();
            soft_select_kernel<8><<<mgrid, kPxThreads, 0, st>>>(cache, c, HW, ws.gmax, o, top, low, ignore_label);
        } else {
            refine_max_kernel<8, false><<<mgrid, kPxThreads, 0, st>>>(a, nullptr, ws.gmax);
            REGDA_LAUNCH_CHECK();
            refine_select_kernel<8><<<grid, kPxThreads, 0, st>>>(a, ws.gmax, o, top, low, ignore_label);
        }
    } else {
        if (cache) {
            refine_max_kernel<16, true><<<mgrid, kPxThreads, 0, st>>>(a, cache, ws.gmax);
            REGDA_LAUNCH_CHECK();
            soft_select_kernel<16><<<mgrid, kPxThreads, 0, st>>>(cache, c, HW, ws.gmax, o, top, low, ignore_label);
        } else {
            refine_max_kernel<16, false><<<mgrid, kPxThreads, 0, st>>>(a, nullptr, ws.gmax);
            REGDA_LAUNCH_CHECK();
            refine_select_kernel<16><<<grid, kPxThreads, 0, st>>>(a, ws.gmax, o, top, low, ignore_label);
        }
    }
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

// workspace of regda_refine_select[_bf16feat] including the refined-probability cache (see refine_select_impl)
extern "C" size_t regda_refine_select_workspace_bytes(int b, int c, int k, int h, int w, int H, int W) {
    if (b < 0 || c < 1 || k < 1 || h < 1 || w < 1 || H < 1 || W < 1) return 0;
    return align_up(carve_refine_ws(nullptr, b, c, k, h, w).bytes, 256) + static_cast<size_t>(b) * c * H * W * sizeof(float);
}

extern "C" int regda_refine_select(const float *feat_nhwc, const float *prototypes, const float *pred1, const float *pred2,
                                   const float *soft_in, int64_t *hard_out, int b, int c, int k, int h, int w, int H, int W,
                                   double temp, double cutoff_top, double cutoff_low, int64_t ignore_label,
                                   void *workspace, size_t workspace_bytes, void *stream) {
    return refine_select_impl(feat_nhwc, false, prototypes, pred1, pred2, soft_in, hard_out, b, c, k, h, w, H, W, temp, cutoff_top, cutoff_low,
                              ignore_label, workspace, workspace_bytes, stream);
}

// same with the feature rows in bf16 (what the model's InstanceNorm kernel writes; float32 arithmetic inside as before)
extern "C" int regda_refine_select_bf16feat(const void *feat_nhwc_bf16, const float *prototypes, const float *pred1, const float *pred2,
                                            const float *soft_in, int64_t *hard_out, int b, int c, int k, int h, int w, int H, int W,
                                            double temp, double cutoff_top, double cutoff_low, int64_t ignore_label,
                                            void *workspace, size_t workspace_bytes, void *stream) {
    return refine_select_impl(feat_nhwc_bf16, true, prototypes, pred1, pred2, soft_in, hard_out, b, c, k, h, w, H, W, temp, cutoff_top, cutoff_low,
                              ignore_label, workspace, workspace_bytes, stream);
}

extern "C" size_t regda_select_workspace_bytes(int b, int c) {
    if (b < 0 || c < 1) return 0;
    return align_up(static_cast<size_t>(b) * c * 4, 256);
}

extern "C" int regda_pseudo_select(const float *soft, int64_t *out, int b, int c, int64_t hw, double cutoff_top, double cutoff_low,
                                   int64_t ignore_label, int32_t *flags, void *workspace, size_t workspace_bytes, void *stream) {
    if (b < 0 || c < 1 || hw < 0) return fail(REGDA_ERR_INVALID_ARG, "pseudo_select: bad shape");
    if (c > 32) return fail(REGDA_ERR_UNSUPPORTED, "pseudo_select: at most 32 classes");
    if (b == 0 || hw == 0) return REGDA_OK;
    if (!soft || !out) return fail(REGDA_ERR_INVALID_ARG, "pseudo_select: null pointer");
    if (!workspace || workspace_bytes < regda_select_workspace_bytes(b, c)) return fail(REGDA_ERR_WORKSPACE, "pseudo_select: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned *gmax = static_cast<unsigned *>(workspace);
    REGDA_CUDA_CHECK(cudaMemsetAsync(gmax, 0, static_cast<size_t>(b) * c * 4, st));
    const int per_img = static_cast<int>(std::min<int64_t>((hw + kPxThreads - 1) / kPxThreads, std::max(1, 16 * sm_count() / b)));
    const dim3 grid(per_img, b);
    const float top = static_cast<float>(cutoff_top), low = static_cast<float>(cutoff_low);
    long long *o = reinterpret_cast<long long *>(out);
#define REGDA_SELECT(CM)                                                                         \
    soft_max_kernel<CM><<<grid, kPxThreads, 0, st>>>(soft, c, hw, gmax, flags);                  \
    REGDA_LAUNCH_CHECK();                                                                        \
    soft_select_kernel<CM><<<grid, kPxThreads, 0, st>>>(soft, c, hw, gmax, o, top, low, ignore_label);
    if (c <= 8) { REGDA_SELECT(8) } else if (c <= 16) { REGDA_SELECT(16) } else { REGDA_SELECT(32) }
#undef REGDA_SELECT
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}
