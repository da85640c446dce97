// PrototypeContrastiveLoss (reference regda/loss.py:18-47), the alignment loss of the stage-2 step
// (tools/train_align_reg.py:186-189), forward and backward:
//   rows f with label != ignore:  fh = f / max(|f|, 1e-12),  ph_c = p_c / max(|p_c|, 1e-12),
//   z_c = fh . ph_c / T,  loss = mean over valid rows of CE(z, label)
// The reference materialises the masked [N,K] feature copy, its normalised copy and the [N,C] logits; here one warp
// streams a row once (HBM-bound: 8 KB per row of 2048 float32), keeps C running dot products against the normalised
// prototypes staged in shared memory, and leaves (cos_c, 1/|f|, label) per row in a small workspace; the backward
// kernel re-reads the row once and writes dL/df = (sum_c w_c ph_c - fh (fh . sum_c w_c ph_c)) / |f| with
// w_c = (softmax_c - [c = label]) / (T * n_valid) * upstream.  The loss is a fixed-order sum (no float atomics).
#include <algorithm>

#include "common.cuh"

namespace regda {
namespace {

constexpr int kPclThreads = 256;
constexpr int kPclMaxC = 8;

struct PclArgs {
    const float *feat;        // [n][k]
    const long long *label;   // [n]
    const float *proto;       // [c][k]
    float *rowinfo;           // [n][kPclMaxC + 2]: cos_c (c < C), inv_norm, label (as float, -1 = ignored)
    float *partial;           // [blocks][2]: loss sum, valid count
    int32_t *flags;
    long long n;
    int k, c;
    long long ignore_label;
    float inv_temp;
};

__device__ __forceinline__ void stage_protos(const PclArgs &a, float *sp) {
    // normalised prototypes into shared memory: warp w handles classes w, w + 8, ...
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int cls = warp; cls < a.c; cls += kPclThreads / 32) {
        const float *p = a.proto + static_cast<size_t>(cls) * a.k;
        float s = 0.f;
        for (int i = lane; i < a.k; i += 32) s = fmaf(p[i], p[i], s);
        s = warp_sum(s);
        const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
        for (int i = lane; i < a.k; i += 32) sp[cls * a.k + i] = p[i] * inv;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kPclThreads)
pcl_fwd_kernel(const PclArgs a) {
    extern __shared__ float sp[];                 // [c][k]
    stage_protos(a, sp);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long wid = static_cast<long long>(blockIdx.x) * (kPclThreads / 32) + warp;
    const long long nw = static_cast<long long>(gridDim.x) * (kPclThreads / 32);
    float loss_sum = 0.f, cnt = 0.f;
    bool bad = false;
    for (long long r = wid; r < a.n; r += nw) {
        const long long l = a.label[r];
        float *info = a.rowinfo + r * (kPclMaxC + 2);
        if (l == a.ignore_label) {
            if (lane == 0) info[kPclMaxC + 1] = -1.f;
            continue;
        }
        if (static_cast<unsigned long long>(l) >= static_cast<unsigned long long>(a.c)) {      // nn.CrossEntropyLoss raises
            bad = true;
            if (lane == 0) info[kPclMaxC + 1] = -1.f;
            continue;
        }
        const float4 *f4 = reinterpret_cast<const float4 *>(a.feat + r * a.k);
        float ss = 0.f, dot[kPclMaxC];
#pragma unroll
        for (int j = 0; j < kPclMaxC; ++j) dot[j] = 0.f;
        for (int i = lane; i < a.k / 4; i += 32) {
            const float4 v = __ldg(f4 + i);
            ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss))));
#pragma unroll
            for (int j = 0; j < kPclMaxC; ++j)
                if (j < a.c) {
                    const float4 p = *reinterpret_cast<const float4 *>(sp + j * a.k + 4 * i);
                    dot[j] = fmaf(v.x, p.x, fmaf(v.y, p.y, fmaf(v.z, p.z, fmaf(v.w, p.w, dot[j]))));
                }
        }
        ss = warp_sum(ss);
#pragma unroll
        for (int j = 0; j < kPclMaxC; ++j)
            if (j < a.c) dot[j] = warp_sum(dot[j]);
        const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
        float m = -INFINITY, z[kPclMaxC];
#pragma unroll
        for (int j = 0; j < kPclMaxC; ++j)
            if (j < a.c) { z[j] = dot[j] * inv * a.inv_temp; m = fmaxf(m, z[j]); }
        float s = 0.f, zl = 0.f;
#pragma unroll
        for (int j = 0; j < kPclMaxC; ++j)
            if (j < a.c) { s += expf(z[j] - m); if (j == static_cast<int>(l)) zl = z[j]; }
        loss_sum += logf(s) + m - zl;
        cnt += 1.f;
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < kPclMaxC; ++j)
                if (j < a.c) info[j] = dot[j] * inv;          // cos_c
            info[kPclMaxC] = inv;
            info[kPclMaxC + 1] = static_cast<float>(l);
        }
    }
    if (bad) raise_flag(a.flags, REGDA_FLAG_LABEL_RANGE);
    // every lane of a warp carries the same loss_sum / cnt: one value per warp, fixed-order block sum
    __shared__ float red[2][kPclThreads / 32];
    if (lane == 0) { red[0][warp] = loss_sum; red[1][warp] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f, n = 0.f;
        for (int i = 0; i < kPclThreads / 32; ++i) { t += red[0][i]; n += red[1][i]; }
        a.partial[2 * blockIdx.x] = t;
        a.partial[2 * blockIdx.x + 1] = n;
    }
}

// loss = sum / count (0/0 = NaN, as nn.CrossEntropyLoss gives for an empty selection); stats[0] = loss, stats[1] = count
__global__ void __launch_bounds__(256)
pcl_finalize_kernel(const float *__restrict__ partial, int blocks, float *__restrict__ stats) {
    __shared__ double rs[256], rn[256];
    double s = 0.0, n = 0.0;
    for (int i = threadIdx.x; i < blocks; i += 256) { s += static_cast<double>(partial[2 * i]); n += static_cast<double>(partial[2 * i + 1]); }
    rs[threadIdx.x] = s; rn[threadIdx.x] = n;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { rs[threadIdx.x] += rs[threadIdx.x + o]; rn[threadIdx.x] += rn[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { stats[0] = static_cast<float>(rs[0] / rn[0]); stats[1] = static_cast<float>(rn[0]); }
}

__global__ void __launch_bounds__(kPclThreads)
pcl_bwd_kernel(const PclArgs a, const float *__restrict__ stats, const float *__restrict__ upstream, float *__restrict__ dfeat) {
    extern __shared__ float sp[];
    stage_protos(a, sp);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long wid = static_cast<long long>(blockIdx.x) * (kPclThreads / 32) + warp;
    const long long nw = static_cast<long long>(gridDim.x) * (kPclThreads / 32);
    const float scale = (upstream ? upstream[0] : 1.0f) * a.inv_temp / stats[1];
    for (long long r = wid; r < a.n; r += nw) {
        const float *info = a.rowinfo + r * (kPclMaxC + 2);
        float4 *d4 = reinterpret_cast<float4 *>(dfeat + r * a.k);
        const float lf = info[kPclMaxC + 1];
        if (lf < 0.f) {
            for (int i = lane; i < a.k / 4; i += 32) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        const int l = static_cast<int>(lf);
        const float inv = info[kPclMaxC];
        float w[kPclMaxC], m = -INFINITY, cosv[kPclMaxC];
#pragma unroll
        for (int j = 0; j < kPclMaxC; ++j)
            if (j < a.c) { cosv[j] = info[j]; m = fmaxf(m, cosv[j] * a.inv_temp); }
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kPclMaxC; ++j)
            if (j < a.c) { w[j] = expf(cosv[j] * a.inv_temp - m); s += w[j]; }
        float proj = 0.f;                          // fh . sum_c w_c ph_c = sum_c w_c cos_c
#pragma unroll
        for (int j = 0; j < kPclMaxC; ++j)
            if (j < a.c) { w[j] = (w[j] / s - (j == l ? 1.f : 0.f)) * scale; proj = fmaf(w[j], cosv[j], proj); }
        const float4 *f4 = reinterpret_cast<const float4 *>(a.feat + r * a.k);
        const float fcoef = proj * inv * inv;      // f * inv^2 * proj = fh * proj / |f|
        for (int i = lane; i < a.k / 4; i += 32) {
            const float4 v = __ldg(f4 + i);
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < kPclMaxC; ++j)
                if (j < a.c) {
                    const float4 p = *reinterpret_cast<const float4 *>(sp + j * a.k + 4 * i);
                    g.x = fmaf(w[j], p.x, g.x); g.y = fmaf(w[j], p.y, g.y); g.z = fmaf(w[j], p.z, g.z); g.w = fmaf(w[j], p.w, g.w);
                }
            d4[i] = make_float4(g.x * inv - v.x * fcoef, g.y * inv - v.y * fcoef, g.z * inv - v.z * fcoef, g.w * inv - v.w * fcoef);
        }
    }
}

int pcl_blocks(long long n) { return static_cast<int>(std::min<long long>((n + 7) / 8, 4ll * sm_count())); }

}  // namespace
}  // namespace regda

using namespace regda;

extern "C" size_t regda_pcl_workspace_bytes(int64_t n) {
    if (n < 0) return 0;
    return align_up(static_cast<size_t>(n) * (kPclMaxC + 2) * 4, 256) + align_up(static_cast<size_t>(pcl_blocks(n)) * 8 + 8, 256);
}

static int pcl_args(PclArgs &a, const float *feat, const int64_t *label, const float *proto, int64_t n, int k, int c, int64_t ignore_label,
                    double temperature, int32_t *flags, void *workspace, size_t workspace_bytes) {
    if (n < 0 || k < 4 || k % 4 != 0 || c < 1) return fail(REGDA_ERR_INVALID_ARG, "pcl: bad shape");
    if (c > kPclMaxC || static_cast<size_t>(c) * k * 4 > 200 * 1024) return fail(REGDA_ERR_UNSUPPORTED, "pcl: at most 8 classes / 200 KB of prototypes");
    if (!(temperature > 0)) return fail(REGDA_ERR_INVALID_ARG, "pcl: temperature must be positive");
    if (n > 0 && (!feat || !label || !proto)) return fail(REGDA_ERR_INVALID_ARG, "pcl: null pointer");
    if (!workspace || workspace_bytes < regda_pcl_workspace_bytes(n)) return fail(REGDA_ERR_WORKSPACE, "pcl: workspace too small");
    a.feat = feat; a.label = reinterpret_cast<const long long *>(label); a.proto = proto;
    a.rowinfo = static_cast<float *>(workspace);
    a.partial = reinterpret_cast<float *>(static_cast<char *>(workspace) + align_up(static_cast<size_t>(n) * (kPclMaxC + 2) * 4, 256));
    a.flags = flags; a.n = n; a.k = k; a.c = c; a.ignore_label = ignore_label; a.inv_temp = static_cast<float>(1.0 / temperature);
    return REGDA_OK;
}

// stats float32 [2] receives (loss, number of valid rows); keep `workspace` and `stats` for regda_pcl_backward
extern "C" int regda_pcl_forward(const float *feat, const int64_t *label, const float *proto, float *stats, int64_t n, int k, int c,
                                 int64_t ignore_label, double temperature, int32_t *flags, void *workspace, size_t workspace_bytes,
                                 void *stream) {
    PclArgs a;
    if (int rc = pcl_args(a, feat, label, proto, n, k, c, ignore_label, temperature, flags, workspace, workspace_bytes)) return rc;
    if (!stats) return fail(REGDA_ERR_INVALID_ARG, "pcl_forward: null stats");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = std::max(1, pcl_blocks(n));
    const size_t smem = static_cast<size_t>(c) * k * 4;
    if (smem > 32 * 1024) REGDA_CUDA_CHECK(cudaFuncSetAttribute(pcl_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    pcl_fwd_kernel<<<blocks, kPclThreads, smem, st>>>(a);
    REGDA_LAUNCH_CHECK();
    pcl_finalize_kernel<<<1, 256, 0, st>>>(a.partial, blocks, stats);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

// dfeat float32 [n][k] = upstream[0] * dloss/dfeat (upstream: device scalar, NULL = 1); rows that were ignored get zeros
extern "C" int regda_pcl_backward(const float *feat, const int64_t *label, const float *proto, const float *stats, const float *upstream,
                                  float *dfeat, int64_t n, int k, int c, int64_t ignore_label, double temperature, void *workspace,
                                  size_t workspace_bytes, void *stream) {
    PclArgs a;
    if (int rc = pcl_args(a, feat, label, proto, n, k, c, ignore_label, temperature, nullptr, workspace, workspace_bytes)) return rc;
    if (!stats || !dfeat) return fail(REGDA_ERR_INVALID_ARG, "pcl_backward: null pointer");
    if (n == 0) return REGDA_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = static_cast<size_t>(c) * k * 4;
    if (smem > 32 * 1024) REGDA_CUDA_CHECK(cudaFuncSetAttribute(pcl_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    pcl_bwd_kernel<<<pcl_blocks(n), kPclThreads, smem, st>>>(a, stats, upstream, dfeat);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}
