// Prototype side of the step (reference regda/gast/alignment.py):
//   DownscaleLabel.forward            :466-481   (integer majority vote per scale x scale block)
//   _compute_local_prototypes sums    :300-327   (6-row segmented sum instead of the [n,c,k] temp)
//   _ema / init_avg                   :435-438, :121-122
#include <cuda_bf16.h>

#include "common.cuh"

namespace regda {
namespace {

// one warp per output cell; classes 0..n_classes (the last plane is "ignored")
__global__ void __launch_bounds__(256)
downscale_kernel(const long long *__restrict__ label, long long *__restrict__ out, int b, int H, int W, int scale,
                 int n_classes, long long ignore_label, float min_ratio, int32_t *flags) {
    const int th = H / scale, tw = W / scale;
    const long long cell = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (cell >= static_cast<long long>(b) * th * tw) return;
    const int lane = threadIdx.x & 31;
    const int img = static_cast<int>(cell / (th * tw));
    const int rem = static_cast<int>(cell - static_cast<long long>(img) * th * tw);
    const int cy = rem / tw, cx = rem - cy * tw;
    const long long *base = label + (static_cast<size_t>(img) * H + static_cast<size_t>(cy) * scale) * W + static_cast<size_t>(cx) * scale;
    const int npx = scale * scale;
    int best = -1, arg = 0;
    bool bad = false;
    // count class by class with ballots: lanes keep their pixels' codes in a small loop
    for (int k = 0; k <= n_classes; ++k) {
        int cnt = 0;
        for (int i0 = 0; i0 < npx; i0 += 32) {
            const int i = i0 + lane;
            bool hit = false;
            if (i < npx) {
                const int dy = i / scale, dx = i - dy * scale;
                long long l = base[static_cast<size_t>(dy) * W + dx];
                if (l == ignore_label) l = n_classes;                       // (:474)
                if (l < 0 || l > n_classes) bad = true;                     // one_hot would raise
                hit = l == k;
            }
            cnt += __popc(__ballot_sync(0xffffffffu, hit));
        }
        if (cnt > best) { best = cnt; arg = k; }                            // first maximum wins (:478)
    }
    if (bad) raise_flag(flags, REGDA_FLAG_LABEL_RANGE);
    if (lane == 0) {
        const float ratio = __fdiv_rn(static_cast<float>(best), static_cast<float>(npx));   // avg_pool2d (:477)
        long long v = arg;
        if (arg == n_classes) v = ignore_label;                             // (:479)
        if (ratio < min_ratio) v = ignore_label;                            // (:480)
        out[cell] = v;
    }
}

// Segmented sum: rows [n][k] float32 by label in [0,c) -> partial[chunk][c][k] (+ counts).
// Block = (chunk of rows) x (slab of 4*blockDim columns); per-class float4 accumulators in
// registers; label is uniform per row so the class test is warp-uniform.
__device__ __forceinline__ float4 load_feat4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 load_feat4(const __nv_bfloat16 *p) {
    const uint2 u = *reinterpret_cast<const uint2 *>(p);
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
}

template <int CMAX, typename T>
__global__ void __launch_bounds__(256)
class_sums_partial_kernel(const T *__restrict__ feat, const long long *__restrict__ lab, float *__restrict__ partial,
                          float *__restrict__ pcount, long long n, int c, int k, int rows_per_chunk, long long ignore_label) {
    const int chunk = blockIdx.y;
    const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const long long r0 = static_cast<long long>(chunk) * rows_per_chunk;
    const long long r1 = min(n, r0 + rows_per_chunk);
    float4 acc[CMAX];
    float cnt[CMAX];
#pragma unroll
    for (int j = 0; j < CMAX; ++j) { acc[j] = make_float4(0.f, 0.f, 0.f, 0.f); cnt[j] = 0.f; }
    if (col < k) {
        // 8 rows per trip: the labels, then the feature loads of the rows that count, are all in flight before the first add (one
        // row per trip was a chain of 64 dependent load latencies); rows are added in row order, as before
        constexpr int kU = 8;
        for (long long rb = r0; rb < r1; rb += kU) {
            int cls[kU];
            float4 v[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const long long l = rb + u < r1 ? lab[rb + u] : ignore_label;
                cls[u] = (l == ignore_label || l < 0 || l >= c) ? -1 : static_cast<int>(l);      // (:442-452) ignored rows add nothing
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (cls[u] >= 0) v[u] = load_feat4(feat + (rb + u) * k + col);
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                // (multiply-by-mask keeps the per-class accumulators in registers: the compare-and-add form is turned into a
                // dynamically indexed local-memory array; x + 0 * v is exact, so the sums are the same bits)
#pragma unroll
                for (int j = 0; j < CMAX; ++j) {
                    const float m = j == cls[u] ? 1.f : 0.f;
                    acc[j].x = fmaf(m, v[u].x, acc[j].x); acc[j].y = fmaf(m, v[u].y, acc[j].y);
                    acc[j].z = fmaf(m, v[u].z, acc[j].z); acc[j].w = fmaf(m, v[u].w, acc[j].w);
                    cnt[j] += m;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < CMAX; ++j)
            if (j < c) *reinterpret_cast<float4 *>(partial + (static_cast<size_t>(chunk) * c + j) * k + col) = acc[j];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < CMAX; ++j) if (j < c) pcount[chunk * c + j] = cnt[j];
    }
}

// fixed-order reduction over chunks (deterministic), optional accumulate into the running sums
__global__ void __launch_bounds__(256)
class_sums_reduce_kernel(const float *__restrict__ partial, const float *__restrict__ pcount, float *__restrict__ sums,
                         float *__restrict__ counts, int nchunks, int c, int k, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c * k) {
        float s = 0.f;
        for (int ch = 0; ch < nchunks; ++ch) s += partial[static_cast<size_t>(ch) * c * k + i];
        sums[i] = accumulate ? sums[i] + s : s;
    }
    if (i < c) {
        float s = 0.f;
        for (int ch = 0; ch < nchunks; ++ch) s += pcount[ch * c + i];
        counts[i] = accumulate ? counts[i] + s : s;
    }
}

// local = sums/(n+eps); where(n<1, proto, local); proto = (1-decay)*local + decay*proto  (:319-325, :435-438)
__global__ void __launch_bounds__(256)
proto_ema_kernel(float *__restrict__ proto, const float *__restrict__ sums, const float *__restrict__ counts,
                 int c, int k, float one_minus_decay, float decay, int init_avg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c * k) return;
    const float n = counts[i / k];
    const float local = sums[i] / (n + 1e-7f);
    if (init_avg) { proto[i] = local; return; }                              // init_avg (:121-122)
    const float old = proto[i];
    const float cur = n < 1.0f ? old : local;
    proto[i] = one_minus_decay * cur + decay * old;
}

constexpr int kRowsPerChunk = 64;

}  // namespace
}  // namespace regda

using namespace regda;

extern "C" int regda_downscale_label(const int64_t *label, int64_t *out, int b, int H, int W, int scale,
                                     int n_classes, int64_t ignore_label, double min_ratio,
                                     int32_t *flags, void *stream) {
    if (b < 0 || H < 0 || W < 0 || scale < 1 || n_classes < 1) return fail(REGDA_ERR_INVALID_ARG, "downscale_label: bad shape");
    const long long cells = static_cast<long long>(b) * (H / scale) * (W / scale);
    if (cells == 0) return REGDA_OK;
    if (!label || !out) return fail(REGDA_ERR_INVALID_ARG, "downscale_label: null pointer");
    downscale_kernel<<<static_cast<unsigned>((cells + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long *>(label), reinterpret_cast<long long *>(out), b, H, W, scale, n_classes,
        ignore_label, static_cast<float>(min_ratio), flags);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

extern "C" size_t regda_class_sums_workspace_bytes(int64_t n, int c, int k) {
    if (n < 0 || c < 1 || k < 1) return 0;
    const size_t nchunks = static_cast<size_t>((n + kRowsPerChunk - 1) / kRowsPerChunk);
    return align_up(nchunks * c * k * 4, 256) + align_up(nchunks * c * 4, 256);
}

static int class_sums_impl(const void *feat_nhwc, bool feat_bf16, const int64_t *label_ds, float *sums, float *counts,
                           int64_t n, int c, int k, int64_t ignore_label, int accumulate,
                           void *workspace, size_t workspace_bytes, void *stream) {
    if (n < 0 || c < 1 || k < 1) return fail(REGDA_ERR_INVALID_ARG, "class_sums: bad shape");
    if (c > 16) return fail(REGDA_ERR_UNSUPPORTED, "class_sums: at most 16 classes");
    if (k % 4 != 0 || (reinterpret_cast<uintptr_t>(feat_nhwc) & 15) || (feat_bf16 && k % 8 != 0)) return fail(REGDA_ERR_UNSUPPORTED, "class_sums: k must be a multiple of 4 (bf16 rows: 8) and rows 16-byte aligned");
    if (!sums || !counts) return fail(REGDA_ERR_INVALID_ARG, "class_sums: null output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) {
        if (!accumulate) {
            REGDA_CUDA_CHECK(cudaMemsetAsync(sums, 0, static_cast<size_t>(c) * k * 4, st));
            REGDA_CUDA_CHECK(cudaMemsetAsync(counts, 0, static_cast<size_t>(c) * 4, st));
        }
        return REGDA_OK;
    }
    if (!feat_nhwc || !label_ds) return fail(REGDA_ERR_INVALID_ARG, "class_sums: null input");
    const size_t need = regda_class_sums_workspace_bytes(n, c, k);
    if (!workspace || workspace_bytes < need) return fail(REGDA_ERR_WORKSPACE, "class_sums: workspace too small");
    const int nchunks = static_cast<int>((n + kRowsPerChunk - 1) / kRowsPerChunk);
    float *partial = static_cast<float *>(workspace);
    float *pcount = reinterpret_cast<float *>(static_cast<char *>(workspace) + align_up(static_cast<size_t>(nchunks) * c * k * 4, 256));
    const int threads = 128;
    const dim3 grid((k / 4 + threads - 1) / threads, nchunks);
    const long long *lab = reinterpret_cast<const long long *>(label_ds);
    const float *ff = static_cast<const float *>(feat_nhwc);
    const __nv_bfloat16 *fb = static_cast<const __nv_bfloat16 *>(feat_nhwc);
    if (feat_bf16) {
        if (c <= 8) class_sums_partial_kernel<8><<<grid, threads, 0, st>>>(fb, lab, partial, pcount, n, c, k, kRowsPerChunk, ignore_label);
        else class_sums_partial_kernel<16><<<grid, threads, 0, st>>>(fb, lab, partial, pcount, n, c, k, kRowsPerChunk, ignore_label);
    } else {
        if (c <= 8) class_sums_partial_kernel<8><<<grid, threads, 0, st>>>(ff, lab, partial, pcount, n, c, k, kRowsPerChunk, ignore_label);
        else class_sums_partial_kernel<16><<<grid, threads, 0, st>>>(ff, lab, partial, pcount, n, c, k, kRowsPerChunk, ignore_label);
    }
    REGDA_LAUNCH_CHECK();
    class_sums_reduce_kernel<<<(c * k + 255) / 256, 256, 0, st>>>(partial, pcount, sums, counts, nchunks, c, k, accumulate);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

extern "C" int regda_class_sums(const float *feat_nhwc, const int64_t *label_ds, float *sums, float *counts,
                                int64_t n, int c, int k, int64_t ignore_label, int accumulate,
                                void *workspace, size_t workspace_bytes, void *stream) {
    return class_sums_impl(feat_nhwc, false, label_ds, sums, counts, n, c, k, ignore_label, accumulate, workspace, workspace_bytes, stream);
}

// same with bf16 feature rows (float32 sums as before)
extern "C" int regda_class_sums_bf16feat(const void *feat_nhwc_bf16, const int64_t *label_ds, float *sums, float *counts,
                                         int64_t n, int c, int k, int64_t ignore_label, int accumulate,
                                         void *workspace, size_t workspace_bytes, void *stream) {
    return class_sums_impl(feat_nhwc_bf16, true, label_ds, sums, counts, n, c, k, ignore_label, accumulate, workspace, workspace_bytes, stream);
}

extern "C" int regda_prototype_ema(float *prototypes, const float *sums, const float *counts, int c, int k, double decay, void *stream) {
    if (c < 1 || k < 1 || !prototypes || !sums || !counts) return fail(REGDA_ERR_INVALID_ARG, "prototype_ema: bad argument");
    if (!(decay > 0.0 && decay < 1.0)) return fail(REGDA_ERR_INVALID_ARG, "prototype_ema: decay must be in (0,1) (alignment.py:312)");
    proto_ema_kernel<<<(c * k + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        prototypes, sums, counts, c, k, static_cast<float>(1.0 - decay), static_cast<float>(decay), 0);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

extern "C" int regda_prototype_init_avg(float *prototypes, const float *sums, const float *counts, int c, int k, void *stream) {
    if (c < 1 || k < 1 || !prototypes || !sums || !counts) return fail(REGDA_ERR_INVALID_ARG, "prototype_init_avg: bad argument");
    proto_ema_kernel<<<(c * k + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(prototypes, sums, counts, c, k, 0.f, 0.f, 1);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}
