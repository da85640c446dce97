// Weight gradient of a convolution on the tcgen05 tensor cores (sm_100a):
//
//   dW[co][r][s][ci] += sum_{n,oh,ow} dY[n,oh,ow,co] * X[n, oh*stride + r*dil - pad, ow*stride + s*dil - pad, ci]
//
// replaces cuDNN's convolution_backward (weight) behind nn.Conv2d in the reference's ResNet / PPM heads
// (regda/_resnets.py:92-112, regda/models/Encoder.py:33-40).
//
// GEMM view: M = Cout, N = Cin, K = pixels, one GEMM per filter tap.  Pixels are the reduction dimension and the
// channels are contiguous in NHWC memory, so BOTH operands are MN-major and are consumed straight from the
// activation tensors: a K-block is a BH x BW patch of 64 output pixels of one image; the A tile is two 4-D TMA
// boxes {64 co, BW, BH, 1} of dY, the B tile BLOCK_N/64 boxes {64 ci, BW, BH, 1} of X at the tap-shifted (and,
// for stride 2, element-strided) coordinates -- halo / padding / patch overhang are TMA zero fill.  Each box
// lands as 64 K-rows of 128 swizzled bytes; UMMA descriptors: LBO = 8 KB between 64-wide M/N blocks, SBO = 1 KB
// between 8-row K groups, +2 KB start address per K=16 step.
//
// Work item = (128-cout tile, BLOCK_N-cin tile, tap, K split); fp32 accumulators in TMEM; the epilogue adds the tile
// into the fp32 OHWI gradient with TMA reduce-stores (split-K partials and the two forward passes of a training step
// accumulate in place).  Warp 0 = TMA, warp 1 = MMA issuer, warps 2-9 = epilogue.  Feature maps smaller than 64 pixels
// (pyramid-pooling branches) put several images into one K-block: box {64 ch, BW, BH, BN}, BN*BH*BW = 64.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace regda {
namespace {

using namespace tc;

constexpr int kWgM = 128;            // cout tile
constexpr int kWgK = 64;             // pixels per stage

struct WgradGeom {
    int n, h, w, cin, cout, oh, ow;
    int r, s, pad, dil, stride;
    int bh, bw, bn, tiles_h, tiles_w; // K-block = bn images x (bh x bw) OUTPUT pixels, bn*bh*bw == 64
    int ksteps;                      // ceil(n / bn) * tiles_h * tiles_w
    int splits, steps_per_split;
    int n_tiles;                     // ceil(cin / BLOCK_N)
    int wct;                         // channels per tap of the gradient tensor (>= cin: gradient of the first cin channels of a wider weight)
};

// ---------------------------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM loops over the work items (cout tile, cin tile, tap, K split), ordered so that
// CTAs running side by side work on the SAME pixel range and tap (the dY / X patches are shared through L2); two
// accumulators in tensor memory let the fp32 reduction of item i into the gradient overlap the MMAs of item i+1;
// BLOCK_N = 256 cin per tile halves the L2->SM bytes per flop of the 128 x 128 tile.
// ---------------------------------------------------------------------------------------------------------
template <int BLOCK_N, int STAGES>
struct WgPersistSmem {
    static constexpr int kABytes = kWgM * kWgK * 2;
    static constexpr int kBBytes = BLOCK_N * kWgK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStoreOffset = STAGES * kStageBytes;     // 4 KB per epilogue warp: one 32 cout x 32 cin fp32 box
    static constexpr int kBarOffset = kStoreOffset + 8 * 4096;
    static constexpr int kTotal = kBarOffset + (2 * STAGES + 4) * 8 + 8;
};

struct WgItem { int m_blk, n_blk, tap, k_lo, k_hi; };

__device__ __forceinline__ WgItem wg_decode(int item, const WgradGeom &g, int m_tiles) {
    WgItem it;
    it.n_blk = item % g.n_tiles; item /= g.n_tiles;
    it.m_blk = item % m_tiles; item /= m_tiles;
    const int taps = g.r * g.s;
    it.tap = item % taps;
    const int split = item / taps;
    it.k_lo = split * g.steps_per_split;
    it.k_hi = min(g.ksteps, it.k_lo + g.steps_per_split);
    return it;
}

constexpr int kWgPersistThreads = 320;      // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quadrant, half the columns each)

// Epilogue: each warp stages 32 cout rows x 32 cin columns of fp32 in 128B-swizzled shared memory and adds the box to the
// gradient with ONE TMA reduce-store (cp.reduce.async.bulk.tensor .add: whole 128-byte lines, reduced in L2) instead of
// 16-byte-per-lane red.global instructions that touch 32 different lines each.
template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kWgPersistThreads, 1)
conv_wgrad_persistent_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                             const __grid_constant__ CUtensorMap tmap_dw,
                             float *__restrict__ dw, const WgradGeom g, const int m_tiles, const int num_items) {
    using L = WgPersistSmem<BLOCK_N, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + L::kBarOffset);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tfull_bar = empty_bar + STAGES;
    uint64_t *tempty_bar = tfull_bar + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    pdl_trigger();
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_dy);
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_dw);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar + i, 1); mbar_init(tempty_bar + i, 8); }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 2 * BLOCK_N);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const int per_img = g.tiles_h * g.tiles_w;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const WgItem it = wg_decode(item, g, m_tiles);
                const int fr = it.tap / g.s, fs = it.tap - fr * g.s;
                for (int kb = it.k_lo; kb < it.k_hi; ++kb) {
                    const int iblk = kb / per_img;
                    const int img = iblk * g.bn;                       // first image of the K-block (small maps: bn images per block)
                    const int t = kb - iblk * per_img;
                    const int th = t / g.tiles_w, tw = t - th * g.tiles_w;
                    const int oh0 = th * g.bh, ow0 = tw * g.bw;
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    uint8_t *sa = smem + stage * L::kStageBytes;
                    uint8_t *sb = sa + L::kABytes;
                    mbar_arrive_expect_tx(full_bar + stage, L::kStageBytes);
#pragma unroll
                    for (int i = 0; i < kWgM / 64; ++i)
                        tma_load_4d(sa + i * 8192, &tmap_dy, full_bar + stage, it.m_blk * kWgM + i * 64, ow0, oh0, img);
                    const int ix = ow0 * g.stride + fs * g.dil - g.pad, iy = oh0 * g.stride + fr * g.dil - g.pad;
#pragma unroll
                    for (int i = 0; i < BLOCK_N / 64; ++i)
                        tma_load_4d(sb + i * 8192, &tmap_x, full_bar + stage, it.n_blk * BLOCK_N + i * 64, ix, iy, img);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(kWgM, BLOCK_N, 1, 1);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const WgItem it = wg_decode(item, g, m_tiles);
                const int num_k = it.k_hi - it.k_lo;
                mbar_wait(tempty_bar + acc, acc_phase ^ 1);
                tc_fence_after_sync();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(full_bar + stage, phase);
                    tc_fence_after_sync();
                    const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
                    const uint32_t sb = sa + L::kABytes;
                    const uint64_t adesc = make_smem_desc(sa, 8192, 1024);
                    const uint64_t bdesc = make_smem_desc(sb, 8192, 1024);
#pragma unroll
                    for (int k = 0; k < kWgK / 16; ++k)
                        umma_bf16(tmem_d, adesc + static_cast<uint64_t>(128 * k), bdesc + static_cast<uint64_t>(128 * k), idesc,
                                  (kb | k) != 0 ? 1u : 0u);
                    umma_commit(empty_bar + stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull_bar + acc);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        const int q = warp & 3;
        const int col_lo = ((warp - 2) >> 2) * (BLOCK_N / 2), col_hi = col_lo + BLOCK_N / 2;
        uint8_t *stage_p = smem + L::kStoreOffset + (warp - 2) * 4096;
        const uint32_t my_row_s = smem_u32(stage_p) + static_cast<uint32_t>(lane) * 128u;
        const uint32_t sw = static_cast<uint32_t>(lane & 7);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const WgItem it = wg_decode(item, g, m_tiles);
            const int row0 = it.m_blk * kWgM + q * 32;                                   // first cout row of this warp's box
            const int colbase = it.tap * g.wct + it.n_blk * BLOCK_N;                     // column in the [cout][taps*wct] matrix
            const bool live = it.k_hi > it.k_lo && row0 < g.cout;
            mbar_wait(tfull_bar + acc, acc_phase);
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
#pragma unroll 1
            for (int c = col_lo; c < col_hi; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + static_cast<uint32_t>(c), v);
                tmem_ld_wait();
                if (c + 32 == col_hi) {
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar + acc);      // accumulator read completely: MMAs of item i+2 may start
                }
                if (live) {
                    if (lane == 0) bulk_wait_read0();                  // previous reduce-store has read the staging buffer
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st_shared_v4(my_row_s + ((static_cast<uint32_t>(j) ^ sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_reduce_add_2d(&tmap_dw, stage_p, colbase + c, row0);
                        bulk_commit();
                    }
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) bulk_wait_read0();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 2 * BLOCK_N);
    }
}

template <int BLOCK_N, int STAGES>
int launch_wgrad_persistent(const CUtensorMap &tdy, const CUtensorMap &tx, float *dw, const WgradGeom &g, cudaStream_t st) {
    using L = WgPersistSmem<BLOCK_N, STAGES>;
    auto kern = conv_wgrad_persistent_kernel<BLOCK_N, STAGES>;
    const int smem = L::kTotal + 1024;
    static_assert(L::kTotal + 1024 <= 232448, "persistent wgrad kernel: shared memory over the 227 KB limit");
    CUtensorMap tdw;
    if (!encode_f32_2d_sw128(&tdw, dw, g.cout, static_cast<long long>(g.r) * g.s * g.wct, 32)) return REGDA_ERR_CUDA;
    REGDA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int m_tiles = (g.cout + kWgM - 1) / kWgM;
    const int num_items = g.n_tiles * m_tiles * g.r * g.s * g.splits;
    const int grid = std::min(num_items, sm_count());
    REGDA_CUDA_CHECK(launch_pdl(kern, dim3(grid), dim3(kWgPersistThreads), smem, st, tdy, tx, tdw, dw, g, m_tiles, num_items));
    return REGDA_OK;
}

}  // namespace
}  // namespace regda

using namespace regda;

extern "C" int regda_conv_wgrad_supported(int n, int h, int w, int cin, int cout, int r, int s, int stride, int pad, int dil) {
    if (n < 1 || h < 1 || w < 1 || stride < 1 || stride > 2 || r < 1 || s < 1 || r * s > 49 || dil < 1 || pad < 0) return 0;
    if (cin % 64 != 0 || cout % 64 != 0) return 0;
    if (h + 2 * pad < dil * (r - 1) + 1 || w + 2 * pad < dil * (s - 1) + 1) return 0;
    return 1;
}

static int wgrad_impl(const void *dy, const void *x, float *dw, int wct, int n, int h, int w, int cin, int cout,
                      int r, int s, int stride, int pad, int dil, void *stream) {
    if (!regda_conv_wgrad_supported(n, h, w, cin, cout, r, s, stride, pad, dil))
        return fail(REGDA_ERR_UNSUPPORTED, "conv_wgrad: shape not covered by the tcgen05 kernel");
    if (!dy || !x || !dw) return fail(REGDA_ERR_INVALID_ARG, "conv_wgrad: null pointer");
    if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dw)) & 15)
        return fail(REGDA_ERR_INVALID_ARG, "conv_wgrad: tensors must be 16-byte aligned");
    WgradGeom g;
    g.n = n; g.h = h; g.w = w; g.cin = cin; g.cout = cout; g.r = r; g.s = s; g.pad = pad; g.dil = dil; g.stride = stride;
    g.wct = wct;
    g.oh = (h + 2 * pad - dil * (r - 1) - 1) / stride + 1;
    g.ow = (w + 2 * pad - dil * (s - 1) - 1) / stride + 1;
    int bw = 1;
    while (bw * 2 <= g.ow && bw * 2 <= kWgK) bw *= 2;
    int bh = kWgK / bw, bn = 1;
    if (static_cast<long long>(g.oh) * g.ow < kWgK) {
        // small map (pyramid-pooling branches): a K-block of 64 pixels spans several images
        bh = 1;
        while (bh * 2 <= g.oh && bw * bh * 2 <= kWgK) bh *= 2;
        bn = kWgK / (bw * bh);
    }
    g.bw = bw; g.bh = bh; g.bn = bn;
    g.tiles_w = (g.ow + g.bw - 1) / g.bw;
    g.tiles_h = (g.oh + g.bh - 1) / g.bh;
    g.ksteps = ((n + bn - 1) / bn) * g.tiles_h * g.tiles_w;
    const int block_n = cin % 256 == 0 ? 256 : (cin % 128 == 0 ? 128 : 64);
    g.n_tiles = cin / block_n;
    const int tiles = g.n_tiles * ((cout + kWgM - 1) / kWgM) * r * s;
    // split K so that every SM gets work (~2 items per SM), >= 8 K-steps per item
    int splits = (2 * sm_count() + tiles - 1) / tiles;
    splits = std::max(1, std::min(splits, g.ksteps / 8));
    g.steps_per_split = (g.ksteps + splits - 1) / splits;
    g.splits = (g.ksteps + g.steps_per_split - 1) / g.steps_per_split;
    ensure_context(dy);
    CUtensorMap tdy, tx;
    if (!encode_nhwc(&tdy, dy, n, g.oh, g.ow, cout, g.bw, g.bh, 1, g.bn)) return REGDA_ERR_CUDA;   /* text set by encode_bf16_sw128 (dy) */
    if (!encode_nhwc(&tx, x, n, h, w, cin, g.bw, g.bh, stride, g.bn)) return REGDA_ERR_CUDA;   /* text set by encode_bf16_sw128 (x) */
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (block_n == 256) return launch_wgrad_persistent<256, 4>(tdy, tx, dw, g, st);
    if (block_n == 128) return launch_wgrad_persistent<128, 6>(tdy, tx, dw, g, st);
    return launch_wgrad_persistent<64, 8>(tdy, tx, dw, g, st);
}

// dy bf16 [n][oh][ow][cout], x bf16 [n][h][w][cin], dw fp32 [cout][r][s][cin] (ACCUMULATED into)
extern "C" int regda_conv_wgrad_bf16(const void *dy, const void *x, float *dw, int n, int h, int w, int cin, int cout,
                                     int r, int s, int stride, int pad, int dil, void *stream) {
    return wgrad_impl(dy, x, dw, cin, n, h, w, cin, cout, r, s, stride, pad, dil, stream);
}

// gradient of the FIRST cin channels of a weight with wct >= cin channels per tap: dw fp32 [cout][r][s][wct], columns [0, cin) of every tap
extern "C" int regda_conv_wgrad_wslice_bf16(const void *dy, const void *x, float *dw, int wct, int n, int h, int w, int cin, int cout,
                                            int r, int s, int stride, int pad, int dil, void *stream) {
    if (wct < cin || wct % 4) return fail(REGDA_ERR_INVALID_ARG, "conv_wgrad_wslice: weight channel count must be >= cin and a multiple of 4");
    return wgrad_impl(dy, x, dw, wct, n, h, w, cin, cout, r, s, stride, pad, dil, stream);
}
