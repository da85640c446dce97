// loss_calc + CrossEntropy (reference regda/utils/tools.py:240-252, regda/gast/balance.py:88-101):
// bilinear (align_corners=True) upsampling of one head's low-resolution logits, per-pixel
// cross entropy with ignore_index, MEAN OVER ALL PIXELS (ignored ones add 0), and -- in the same
// pass -- the gradient w.r.t. the low-resolution logits.  The [b,c,H,W] upsampled logits, their
// softmax and their gradient are never materialised (the reference writes each of them to HBM).
//
// One block per (image, strip of ceil(H/h) full-resolution rows): such a strip touches at most
// three low-resolution rows, which are staged in shared memory.  A thread walks one column
// of the strip, accumulating the row-direction part of the transposed interpolation in
// registers and parking it in shared memory ([3][c][W]); the column-direction part is then a
// GATHER: one thread per staged low-resolution element sums the ~2*W/w columns whose
// interpolation footprint covers it, in a fixed order (no shared-memory float atomics: those
// compile to compare-and-swap loops that serialise 16-way on the ~W/w columns sharing a
// target), then one global atomicAdd per touched low-resolution element.
#include <algorithm>

#include "common.cuh"

namespace regda {
namespace {

constexpr int kCeThreads = 256;

struct CeArgs {
    const float *pred;       // [b][c][h][w]
    const long long *label;  // [b][H][W]
    float *dpred;            // [b][c][h][w] or null
    float *partial;          // [blocks] loss partial sums
    const float *class_weight;  // [c] or null (ClassBalance, balance.py:30-33)
    int32_t *flags;
    int c, h, w, H, W, rows_per_block;
    long long ignore_label;
    float sy, sx, gscale;    // gscale = grad_scale / (b*H*W)
};

template <int CMAX>
__global__ void __launch_bounds__(kCeThreads, CMAX <= 8 ? 3 : 1)
ce_bilinear_kernel(const CeArgs a) {
    extern __shared__ float sm[];
    const int img = blockIdx.y;
    const int Y0 = blockIdx.x * a.rows_per_block;
    const int Y1 = min(a.H, Y0 + a.rows_per_block);
    const int ylo = static_cast<int>(a.sy * static_cast<float>(Y0));
    const int nrow = min(3, a.h - ylo);
    const int plane = a.h * a.w;
    float *lo = sm;                        // [3][c][w] staged logits
    float *gcol = sm + 3 * a.c * a.w;      // [3][c][W] per-column gradient (row direction already applied)
    const float *pimg = a.pred + static_cast<size_t>(img) * a.c * plane;
    for (int i = threadIdx.x; i < 3 * a.c * a.w; i += kCeThreads) {
        const int r = i / (a.c * a.w), rem = i - r * a.c * a.w, j = rem / a.w, x = rem - j * a.w;
        lo[i] = r < nrow ? pimg[static_cast<size_t>(j) * plane + (ylo + r) * a.w + x] : 0.f;
    }
    __syncthreads();

    float loss_sum = 0.f;
    bool bad = false;
    const long long *lab = a.label + static_cast<size_t>(img) * a.H * a.W;
    for (int X = threadIdx.x; X < a.W; X += kCeThreads) {
        const float fx = a.sx * static_cast<float>(X);
        const int x0 = static_cast<int>(fx);
        const int x1 = x0 + (x0 < a.w - 1 ? 1 : 0);
        const float wx1 = fx - static_cast<float>(x0), wx0 = 1.0f - wx1;
        // the column direction of the interpolation does not depend on Y: interpolate the three staged rows once per column
        float g[3][CMAX], hx[3][CMAX];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int j = 0; j < CMAX; ++j) {
                g[r][j] = 0.f;
                hx[r][j] = 0.f;
                if (j < a.c) {
                    const float *p = lo + (r * a.c + j) * a.w;
                    hx[r][j] = wx0 * p[x0] + wx1 * p[x1];
                }
            }
        for (int Y = Y0; Y < Y1; ++Y) {
            const long long l = lab[static_cast<size_t>(Y) * a.W + X];
            if (l == a.ignore_label) continue;
            if (static_cast<unsigned long long>(l) >= static_cast<unsigned long long>(a.c)) { bad = true; continue; }
            const float fy = a.sy * static_cast<float>(Y);
            const int y0 = static_cast<int>(fy);
            const int y1 = y0 + (y0 < a.h - 1 ? 1 : 0);
            const float wy1 = fy - static_cast<float>(y0), wy0 = 1.0f - wy1;
            const int r0 = y0 - ylo, r1 = y1 - ylo;       // block-uniform (0..2): the row selects below do not diverge
            float z[CMAX];
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < CMAX; ++j)
                if (j < a.c) {
                    const float h0 = r0 == 0 ? hx[0][j] : (r0 == 1 ? hx[1][j] : hx[2][j]);
                    const float h1 = r1 == 0 ? hx[0][j] : (r1 == 1 ? hx[1][j] : hx[2][j]);
                    z[j] = wy0 * h0 + wy1 * h1;
                    m = fmaxf(m, z[j]);
                }
            float s = 0.f, zl = 0.f;
#pragma unroll
            for (int j = 0; j < CMAX; ++j)
                if (j < a.c) {
                    if (j == static_cast<int>(l)) zl = z[j];
                    z[j] = expf(z[j] - m);
                    s += z[j];
                }
            const float wgt = a.class_weight ? a.class_weight[l] : 1.0f;
            loss_sum += wgt * (logf(s) + m - zl);            // -log_softmax[label]
            const float inv = wgt / s;
            const float w0 = (r0 == 0 ? wy0 : 0.f) + (r1 == 0 ? wy1 : 0.f);
            const float w1 = (r0 == 1 ? wy0 : 0.f) + (r1 == 1 ? wy1 : 0.f);
            const float w2 = (r0 == 2 ? wy0 : 0.f) + (r1 == 2 ? wy1 : 0.f);
#pragma unroll
            for (int j = 0; j < CMAX; ++j)
                if (j < a.c) {
                    const float gj = z[j] * inv - (j == static_cast<int>(l) ? wgt : 0.f);
                    g[0][j] += w0 * gj;
                    g[1][j] += w1 * gj;
                    g[2][j] += w2 * gj;
                }
        }
        if (a.dpred != nullptr) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int j = 0; j < CMAX; ++j)
                    if (j < a.c) gcol[(r * a.c + j) * a.W + X] = g[r][j];
        }
    }
    if (bad) raise_flag(a.flags, REGDA_FLAG_LABEL_RANGE);

    __shared__ float red[kCeThreads / 32];
    loss_sum = warp_sum(loss_sum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = loss_sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < kCeThreads / 32; ++i) t += red[i];
        a.partial[blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
    if (a.dpred != nullptr) {
        // (the __syncthreads above also ordered the gcol stores)  Column X feeds low-resolution column x0(X) with weight
        // wx0 and x0(X)+1 with wx1, x0 / wx1 computed exactly as in the forward part; the columns that can reach x lie
        // within (x-1)/sx .. (x+1)/sx, scanned with one column of slack on each side.
        float *dimg = a.dpred + static_cast<size_t>(img) * a.c * plane;
        const float inv_sx = a.sx > 0.f ? 1.0f / a.sx : 0.f;
        for (int i = threadIdx.x; i < nrow * a.c * a.w; i += kCeThreads) {
            const int r = i / (a.c * a.w), rem = i - r * a.c * a.w, j = rem / a.w, x = rem - j * a.w;
            int Xlo = 0, Xhi = a.W - 1;
            if (a.sx > 0.f) {
                Xlo = max(0, static_cast<int>(static_cast<float>(x - 1) * inv_sx) - 1);
                Xhi = min(a.W - 1, static_cast<int>(static_cast<float>(x + 1) * inv_sx) + 2);
            }
            const float *gr = gcol + (r * a.c + j) * a.W;
            float v = 0.f;
            for (int X = Xlo; X <= Xhi; ++X) {
                const float fx = a.sx * static_cast<float>(X);
                const int x0 = static_cast<int>(fx);
                const float wx1 = fx - static_cast<float>(x0);
                const bool inner = x0 < a.w - 1;             // at the last column x1 == x0: both weights land on x0
                const float wgt = (x0 == x ? (inner ? 1.0f - wx1 : 1.0f) : 0.f) + ((x0 + 1 == x && inner) ? wx1 : 0.f);
                v = fmaf(wgt, gr[X], v);
            }
            if (v != 0.f) atomicAdd(dimg + static_cast<size_t>(j) * plane + (ylo + r) * a.w + x, v * a.gscale);
        }
    }
}

// fixed-order final sum: loss = sum(partials) / n_pixels  (torch.mean over ALL pixels, balance.py:101)
__global__ void __launch_bounds__(256)
ce_finalize_kernel(const float *__restrict__ partial, int n, float inv_npx, float *__restrict__ loss) {
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += static_cast<double>(partial[i]);
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = static_cast<float>(red[0]) * inv_npx;
}

// ClassBalance._local_freq / _one_hot counts (balance.py:43-66)
__global__ void __launch_bounds__(256)
class_count_kernel(const long long *__restrict__ label, long long n, int c, long long ignore_label,
                   unsigned long long *__restrict__ out, int32_t *flags) {
    extern __shared__ unsigned scnt[];   // [c+1]: classes, then n_valid
    for (int i = threadIdx.x; i <= c; i += blockDim.x) scnt[i] = 0u;
    __syncthreads();
    bool bad = false;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long l = label[i];
        const bool valid = l != ignore_label;
        const unsigned vm = __ballot_sync(__activemask(), valid);
        (void)vm;
        if (valid) {
            atomicAdd(&scnt[c], 1u);
            if (static_cast<unsigned long long>(l) < static_cast<unsigned long long>(c)) atomicAdd(&scnt[l], 1u);
            else if (l != c) bad = true;
        }
    }
    if (bad) raise_flag(flags, REGDA_FLAG_LABEL_RANGE);
    __syncthreads();
    for (int i = threadIdx.x; i <= c; i += blockDim.x)
        if (scnt[i]) atomicAdd(out + i, static_cast<unsigned long long>(scnt[i]));
}

}  // namespace
}  // namespace regda

using namespace regda;

// Strip height: any strip of at most ceil(H/h) full-resolution rows touches <= 3 low-resolution rows.  A quarter of that
// (4 rows at 512 / 32) gives 4x the blocks: 1024 blocks for 8 images instead of 256, which is what fills 148 SMs.
static int ce_rows_per_block(int h, int H) { return std::max(1, ((H + h - 1) / h + 3) / 4); }

extern "C" size_t regda_ce_workspace_bytes(int b, int h, int H) {
    if (b < 0 || h < 1 || H < 1) return 0;
    const int rpb = ce_rows_per_block(h, H);
    return align_up(static_cast<size_t>(b) * ((H + rpb - 1) / rpb) * 4, 256);
}

extern "C" int regda_ce_bilinear(const float *pred, const int64_t *label, float *loss, float *dpred,
                                 int b, int c, int h, int w, int H, int W, int64_t ignore_label,
                                 double grad_scale, const float *class_weight,
                                 int32_t *flags, void *workspace, size_t workspace_bytes, void *stream) {
    if (b < 0 || c < 1 || h < 1 || w < 1 || H < 1 || W < 1) return fail(REGDA_ERR_INVALID_ARG, "ce_bilinear: bad shape");
    if (c > 16) return fail(REGDA_ERR_UNSUPPORTED, "ce_bilinear: at most 16 classes");
    if (!loss) return fail(REGDA_ERR_INVALID_ARG, "ce_bilinear: null loss pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (b == 0) { REGDA_CUDA_CHECK(cudaMemsetAsync(loss, 0, 4, st)); return REGDA_OK; }
    if (!pred || !label) return fail(REGDA_ERR_INVALID_ARG, "ce_bilinear: null pointer");
    const size_t need = regda_ce_workspace_bytes(b, h, H);
    if (!workspace || workspace_bytes < need) return fail(REGDA_ERR_WORKSPACE, "ce_bilinear: workspace too small");
    const size_t smem = static_cast<size_t>(3) * c * (w + W) * 4;      // staged logits [3][c][w] + per-column gradients [3][c][W]
    if (smem > 200 * 1024) return fail(REGDA_ERR_UNSUPPORTED, "ce_bilinear: low-resolution row too wide for shared memory");
    CeArgs a;
    a.pred = pred; a.label = reinterpret_cast<const long long *>(label); a.dpred = dpred;
    a.partial = static_cast<float *>(workspace); a.class_weight = class_weight; a.flags = flags;
    a.c = c; a.h = h; a.w = w; a.H = H; a.W = W; a.rows_per_block = ce_rows_per_block(h, H);
    a.ignore_label = ignore_label;
    a.sy = H > 1 ? static_cast<float>(h - 1) / static_cast<float>(H - 1) : 0.f;
    a.sx = W > 1 ? static_cast<float>(w - 1) / static_cast<float>(W - 1) : 0.f;
    const double npx = static_cast<double>(b) * H * W;
    a.gscale = static_cast<float>(grad_scale / npx);
    if (dpred) REGDA_CUDA_CHECK(cudaMemsetAsync(dpred, 0, static_cast<size_t>(b) * c * h * w * 4, st));
    const dim3 grid((H + a.rows_per_block - 1) / a.rows_per_block, b);
    // (register arrays are sized by the template argument: the 6- and 7-class configurations of the recipes get their own)
#define REGDA_CE_LAUNCH(CM)                                                                                                              \
    do {                                                                                                                                 \
        if (smem > 48 * 1024)                                                                                                            \
            REGDA_CUDA_CHECK(cudaFuncSetAttribute(ce_bilinear_kernel<CM>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
        ce_bilinear_kernel<CM><<<grid, kCeThreads, smem, st>>>(a);                                                                       \
    } while (0)
    if (c == 6) REGDA_CE_LAUNCH(6);
    else if (c == 7) REGDA_CE_LAUNCH(7);
    else if (c <= 8) REGDA_CE_LAUNCH(8);
    else REGDA_CE_LAUNCH(16);
#undef REGDA_CE_LAUNCH
    REGDA_LAUNCH_CHECK();
    ce_finalize_kernel<<<1, 256, 0, st>>>(a.partial, static_cast<int>(grid.x * grid.y), static_cast<float>(1.0 / npx), loss);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

extern "C" int regda_class_count(const int64_t *label, int64_t n, int c, int64_t ignore_label,
                                 int64_t *counts_out, int32_t *flags, void *stream) {
    if (n < 0 || c < 1 || c > 4096) return fail(REGDA_ERR_INVALID_ARG, "class_count: bad shape");
    if (!counts_out) return fail(REGDA_ERR_INVALID_ARG, "class_count: null output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    REGDA_CUDA_CHECK(cudaMemsetAsync(counts_out, 0, static_cast<size_t>(c + 1) * 8, st));
    if (n == 0) return REGDA_OK;
    if (!label) return fail(REGDA_ERR_INVALID_ARG, "class_count: null input");
    const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 4ll * sm_count()));
    class_count_kernel<<<blocks, 256, static_cast<size_t>(c + 1) * 4, st>>>(reinterpret_cast<const long long *>(label), n, c, ignore_label,
                                                                           reinterpret_cast<unsigned long long *>(counts_out), flags);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}
