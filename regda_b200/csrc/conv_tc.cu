// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05) for sm_100a.
//
//   Y[n,oh,ow,co] = sum_{r,s,ci} X[n, oh + r*dil - pad, ow + s*dil - pad, ci] * W[co,r,s,ci]      (stride 1)
//
// replaces the cuDNN convolutions behind nn.Conv2d in the reference's ResNet / PPM heads
// (regda/_resnets.py:92-112, regda/models/Encoder.py:33-40).  NHWC bf16 activations, OHWI bf16
// weights, fp32 accumulation in tensor memory.
//
// GEMM view: M = pixels, N = Cout, K = R*S*Cin.  One CTA computes a 128-pixel x BLOCK_N tile:
//   * the 128 pixels are a BH x BW spatial patch of ONE image, so the A operand of filter tap
//     (r,s) and channel chunk c0 is ONE 4-D TMA box {64 ch, BW, BH, 1} of X at coordinates
//     (c0, ow0 + s*dil - pad, oh0 + r*dil - pad, n): the halo / zero padding is TMA's
//     out-of-bounds zero fill -- no im2col buffer, no index arithmetic on the SM;
//   * the B operand is a 2-D box {64 k, BLOCK_N} of the OHWI weight matrix [Cout][R*S*Cin];
//   * both land in 128B-swizzled K-major shared-memory tiles consumed directly by tcgen05.mma
//     (UMMA 128 x BLOCK_N x 16), accumulating in TMEM;
//   * feature maps smaller than 128 pixels (the 1x1 .. 6x6 pyramid-pooling branches, the deep layers of small
//     tiles) put SEVERAL images into one M tile: the box becomes {64 ch, BW, BH, BN} with BN*BH*BW = 128 -- still
//     one TMA instruction, every image's halo still zero-filled;
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 2-9 = epilogue
//     (tcgen05.ld -> bf16 -> swizzled smem -> TMA store), STAGES-deep mbarrier ring between producer and MMA.
// The same kernel computes the data gradient of a stride-1 convolution when it is given dY and reads the forward
// weights MN-major in place (tap order flipped).  A stride-2 data gradient is the stride-1 one on the zero-inserted dY
// (regda_zero_insert2_bf16).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace regda {
namespace {

using namespace tc;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;          // bf16: 128 bytes = one swizzle span
constexpr int kUmmaK = 16;
constexpr int kPersistThreads = 320; // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two warps per TMEM lane quadrant, each
                                     // draining half of the accumulator's columns)

struct ConvGeom {
    int n, h, w, cin, cout;          // input image, reduction channels, output channels
    int oh, ow;                      // output image
    int r, s, pad, dil, stride;
    int flip;                        // dgrad: weight tap = taps-1-tap
    int bh, bw, bn;                  // M tile = bn images x (bh x bw) output pixels, bn*bh*bw == 128
    int tiles_h, tiles_w, tiles_img; // patches per image, image blocks (ceil(n / bn))
    int kc;                          // cin / 64
    int wct;                         // channels per tap of the WEIGHT tensor (>= the channels this GEMM touches: a convolution over
                                     // the first channels of a wider OHWI weight reads / differentiates it in place)
    int contig;                      // tile walk: 0 = strided, channel block fastest; 1 = contiguous ranges, channel block slowest
    int w_early;                     // the weights were written long before the preceding kernel of the stream: their first pipeline
                                     // stages may be requested BEFORE the programmatic-dependent-launch wait (regda_conv_hint_static_weights)
};

struct TileCoord { int n_blk, tw, th, img0; };

// tile t -> (channel block, M tile).  Default: channel block fastest -- the CTAs that run side by side share an activation patch
// in L2 and, walking the list with a stride of gridDim.x, keep their channel block whenever it divides the grid size.  g.contig
// (channel-block count does not divide the grid): M tile fastest, channel block slowest, each CTA a contiguous range.
__device__ __forceinline__ TileCoord decode_tile(int t, int n_tiles_n, const ConvGeom &g) {
    TileCoord c;
    int m;
    if (g.contig) {
        const int m_tiles = g.tiles_img * g.tiles_h * g.tiles_w;
        c.n_blk = t / m_tiles;
        m = t - c.n_blk * m_tiles;
    } else {
        c.n_blk = t % n_tiles_n;
        m = t / n_tiles_n;
    }
    c.tw = m % g.tiles_w; m /= g.tiles_w;
    c.th = m % g.tiles_h;
    c.img0 = (m / g.tiles_h) * g.bn;
    return c;
}

// ---------------------------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM walks the tile list (decode_tile: N tile fastest, so CTAs running side by side share the
// activation patch in L2); the TMA / MMA / epilogue warps each loop over the tiles on their own, coupled only by
// mbarriers: the shared-memory ring keeps streaming across tile boundaries and TWO accumulators in tensor memory
// (2 x BLOCK_N columns) let the epilogue of tile i drain while the MMAs of tile i+1 run.  Barrier / TMEM /
// tensor-map set-up is paid once per SM instead of once per tile, which is what the many small-K convolutions
// of layer3 (K = 256 .. 2304, 4-36 k-blocks per tile) are bound by.  BLOCK_N = 256 halves the L2->SM bytes per
// flop of the 128 x 128 tile.
// ---------------------------------------------------------------------------------------------------------
template <int BLOCK_N, int STAGES, bool YBUF = false>
struct PersistSmem {
    static constexpr int kABytes = kBlockM * kBlockK * 2;
    static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kEpiWarps = BLOCK_N >= 128 ? 8 : 4; // TMA-store epilogue: warps that drain the accumulator
    static constexpr int kStoreOffset = STAGES * kStageBytes; // 4 KB per epilogue warp: one 32-row x 64-channel bf16 box
    static constexpr int kYOffset = kStoreOffset + 8 * 4096;  // BNRED: 4 KB per epilogue warp for the BatchNorm input box
    static constexpr int kBarOffset = kYOffset + (YBUF ? 8 * 4096 : 0);
    static constexpr int kNumBars = 2 * STAGES + 4 + 8;      // full, empty, tmem_full[2], tmem_empty[2], ybar[8]
    static constexpr int kTotal = kBarOffset + kNumBars * 8 + 8;
};

// STATS: the epilogue also accumulates the per-channel sum and sum of squares of the (bf16-rounded) outputs --
// the train-mode BatchNorm statistics of the layer that follows -- into stats[group][2][cout] with one fp32
// atomic per channel per warp per tile (warp-transposing butterfly: 31 shuffles per quantity per 32 channels),
// which removes BatchNorm's own pass over the activation.  group = image / imgs_per_group.
//
// Epilogue (bf16 output): each warp stages its 32 pixels x 64 channels in 128B-swizzled shared memory and writes them with
// ONE TMA store (full 128-byte lines; the box is clipped at the image border) instead of 16-byte-per-lane scattered stores
// (32 distinct lines per store instruction); the BatchNorm statistics are column sums read back from the staged tile
// (32 conflict-free LDS per lane for 2 channels).
//
// OUT_F32: the accumulator is written as float32 (each lane stores whole 128-byte lines of its pixel row; an optional
// float32 addend is added first).  This is the epilogue of the float32 PARITY mode (ops/tc.py: every float32 operand is
// split into bf16 hi + lo parts and the product hi*hi + hi*lo + lo*hi is accumulated in one pass over a 3x longer K) and of
// any caller that wants the raw fp32 accumulators.
//
// BNRED (data-gradient launches whose output is the gradient of a BatchNorm+ReLU output): the epilogue also performs the
// first half of that BatchNorm's backward.  It masks the gradient with the ReLU bit mask the forward wrote (the stored
// tensor is dz = dout * [out > 0], which is also the residual branch's gradient), TMA-loads the matching box of the
// BatchNorm INPUT y, and accumulates sum(dz) and sum(dz * y) per channel and statistics group into `stats` -- the two
// reductions of bn_bwd_reduce_kernel, which then does not run at all (norm.cu, regda_bn_backward_bf16 dz_ready = 1).
// REGDA_CONV_TRACE (scripts/conv_trace.cu only; never defined in the library build): per-CTA phase time stamps of a launch
#ifdef REGDA_CONV_TRACE
constexpr int kTraceSlots = 32, kTraceCtas = 160, kTraceLaunches = 64;
__device__ unsigned long long g_conv_trace[kTraceLaunches * kTraceCtas * kTraceSlots * 2];
#define CONV_TRACE(slot)                                                                                          \
    do {                                                                                                          \
        unsigned long long gt_;                                                                                   \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                                                   \
        unsigned long long *tp_ = g_conv_trace + ((static_cast<size_t>(trace_id) * kTraceCtas + blockIdx.x) * kTraceSlots + (slot)) * 2; \
        tp_[0] = gt_;                                                                                             \
        tp_[1] = static_cast<unsigned long long>(clock64());                                                     \
    } while (0)
#else
#define CONV_TRACE(slot)
#endif

template <int BLOCK_N, int STAGES, bool B_MN, bool STATS, bool OUT_F32, bool BNRED>
__global__ void __launch_bounds__(kPersistThreads, 1)
conv_persistent_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                       const __grid_constant__ CUtensorMap tmap_y, const __grid_constant__ CUtensorMap tmap_bn,
                       __nv_bfloat16 *__restrict__ y, const ConvGeom g, const int n_tiles_n, const int num_tiles,
#ifdef REGDA_CONV_TRACE
                       float *__restrict__ stats, const int imgs_per_group_, const __nv_bfloat16 *__restrict__ addend,
#else
                       float *__restrict__ stats, const int imgs_per_group, const __nv_bfloat16 *__restrict__ addend,
#endif
                       const unsigned char *__restrict__ relu_mask) {
    static_assert(!BNRED || (!OUT_F32 && !STATS), "BNRED rides on the TMA-store epilogue and shares the statistics registers");
    static_assert(!OUT_F32 || !STATS, "the float32-output epilogue carries no BatchNorm statistics");
    using L = PersistSmem<BLOCK_N, STAGES, BNRED>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + L::kBarOffset);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tfull_bar = empty_bar + STAGES;                // [2] accumulator ready for the epilogue
    uint64_t *tempty_bar = tfull_bar + 2;                    // [2] accumulator drained (4 epilogue warps arrive)
    uint64_t *ybar = tempty_bar + 2;                         // [8] BNRED: per-epilogue-warp "BatchNorm input box landed"
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(ybar + 8);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_k = g.r * g.s * g.kc;
    // this CTA's tiles: every gridDim.x-th tile, or (g.contig) a contiguous range of the list (sizes differ by at most one) -- either
    // way the channel block and the BatchNorm statistics group change at most once each on the way (the fused statistics are
    // flushed when they do; a channel block that alternated from tile to tile would flush at every tile)
    const int tile_lo = g.contig ? static_cast<int>(static_cast<long long>(blockIdx.x) * num_tiles / gridDim.x) : static_cast<int>(blockIdx.x);
    const int tile_hi = g.contig ? static_cast<int>(static_cast<long long>(blockIdx.x + 1) * num_tiles / gridDim.x) : num_tiles;
    const int tile_step = g.contig ? 1 : static_cast<int>(gridDim.x);
#ifdef REGDA_CONV_TRACE
    const int trace_id = (imgs_per_group_ >> 16) % kTraceLaunches;
    const int imgs_per_group = imgs_per_group_ & 0xffff;
    if (threadIdx.x == 0) CONV_TRACE(0);
#endif

    pdl_trigger();          // the next kernel's prologue may overlap this kernel (it blocks in its own pdl_wait)
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_w);
        if (!OUT_F32) prefetch_tmap(&tmap_y);
        if (BNRED) prefetch_tmap(&tmap_bn);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar + i, 1); mbar_init(tempty_bar + i, OUT_F32 ? 8 : L::kEpiWarps); }
        for (int i = 0; i < 8; ++i) mbar_init(ybar + i, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 2 * BLOCK_N);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) CONV_TRACE(1);
    // B operand (weights) of k-block (tap, c0) of channel block n_blk into pipeline stage buffer sb
    auto load_b = [&](uint8_t *sb, uint64_t *bar, int n_blk, int tap, int c0) {
        if (B_MN) {
            const int wtap = g.flip ? g.r * g.s - 1 - tap : tap;
#pragma unroll
            for (int i = 0; i < BLOCK_N / 64; ++i) tma_load_3d(sb + i * 8192, &tmap_w, bar, n_blk * BLOCK_N + i * 64, wtap, c0);
        } else {
            tma_load_2d(sb, &tmap_w, bar, tap * g.wct + c0, n_blk * BLOCK_N);
        }
    };
    // Static weights (g.w_early): the weight halves of this CTA's first pipeline stages do not depend on the kernel this launch
    // waits for -- request them now, so that their DRAM latency overlaps the wait; the activation halves follow after it.
    const int pre = g.w_early ? min(STAGES, num_k) : 0;
    if (pre > 0 && warp == 0 && elect_one()) {
        const int n_blk = decode_tile(tile_lo, n_tiles_n, g).n_blk;
        int tap = 0, c0 = 0;
        for (int kb = 0; kb < pre; ++kb) {
            mbar_arrive_expect_tx(full_bar + kb, L::kStageBytes);
            load_b(smem + kb * L::kStageBytes + L::kABytes, full_bar + kb, n_blk, tap, c0);
            c0 += kBlockK;
            if (c0 == g.cin) { c0 = 0; ++tap; }
        }
    }
    pdl_wait();             // everything above touched only shared / tensor memory and static weights; fresh global memory from here on
    if (threadIdx.x == 0) CONV_TRACE(2);

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            CONV_TRACE(12);
            int stage = 0;
            uint32_t phase = 0;
            int early = pre;                 // k-blocks whose barrier is armed and whose weight half is already on its way
            for (int t = tile_lo; t < tile_hi; t += tile_step) {
                const TileCoord tc_ = decode_tile(t, n_tiles_n, g);
                const int n_blk = tc_.n_blk, img = tc_.img0;
                const int ix0 = tc_.tw * g.bw * g.stride - g.pad, iy0 = tc_.th * g.bh * g.stride - g.pad;
                int tap = 0, fr = 0, fs = 0, c0 = 0;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    uint8_t *sa = smem + stage * L::kStageBytes;
                    const bool armed = early > 0;       // barrier armed and weight half requested before the PDL wait
                    if (armed) --early;
                    else mbar_arrive_expect_tx(full_bar + stage, L::kStageBytes);
                    tma_load_4d(sa, &tmap_x, full_bar + stage, c0, ix0 + fs * g.dil, iy0 + fr * g.dil, img);
                    if (!armed) load_b(sa + L::kABytes, full_bar + stage, n_blk, tap, c0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    c0 += kBlockK;
                    if (c0 == g.cin) { c0 = 0; ++tap; if (++fs == g.s) { fs = 0; ++fr; } }
                }
            }
            CONV_TRACE(13);
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, 0, B_MN ? 1 : 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = tile_lo; t < tile_hi; t += tile_step) {
                mbar_wait(tempty_bar + acc, acc_phase ^ 1);           // epilogue has drained this accumulator
                tc_fence_after_sync();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(full_bar + stage, phase);
                    tc_fence_after_sync();
                    if (kb == 0 && t == tile_lo) CONV_TRACE(3);
                    const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
                    const uint32_t sb = sa + L::kABytes;
                    const uint64_t adesc = make_smem_desc(sa, 0, 1024);
                    const uint64_t bdesc = B_MN ? make_smem_desc(sb, 8192, 1024) : make_smem_desc(sb, 0, 1024);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_bf16(tmem_d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>((B_MN ? 128 : 2) * k), idesc,
                                  (kb | k) != 0 ? 1u : 0u);
                    umma_commit(empty_bar + stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(tfull_bar + acc);
                CONV_TRACE(4);          // (overwritten per tile: the last tile's commit stays)
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (!OUT_F32) {
        // ===== epilogue: TMEM -> registers -> bf16 -> swizzled shared memory -> TMA store (NHWC) =====
        constexpr int kCols = BLOCK_N >= 128 ? BLOCK_N / 2 : BLOCK_N;   // columns per warp: 128 / 64 / 64
        const int wq = warp - 2;
        const int q = warp & 3;                                        // TMEM lane quadrant of this warp
        if (wq < L::kEpiWarps) {
            const int col_lo = (BLOCK_N >= 128 ? (wq >> 2) : 0) * kCols;
            const int row = q * 32 + lane;
            // row -> (image, y, x) inside the tile: x fastest, then y, then the image of a multi-image tile
            const int pw = row % g.bw, ph = (row / g.bw) % g.bh, pn = row / (g.bw * g.bh);
            // the warp's 32 rows are one box {64 ch, sbw px, sbh rows, sbn images} (launch_persistent_impl): its origin in the tile
            const int bpw0 = (q * 32) % g.bw, bph0 = ((q * 32) / g.bw) % g.bh, bpn0 = (q * 32) / (g.bw * g.bh);
            const uint32_t stage_s = smem_u32(smem + L::kStoreOffset + wq * 4096);
            const uint32_t my_row_s = stage_s + static_cast<uint32_t>(lane) * 128u;
            const uint32_t sw = static_cast<uint32_t>(lane & 7);
            uint8_t *ybuf = smem + L::kYOffset + wq * 4096;          // BNRED only
            const uint32_t ybuf_s = smem_u32(ybuf);
            uint32_t yphase = 0;
            // BatchNorm statistics stay in registers across the tiles of this CTA for as long as (channel block, statistics
            // group) does not change -- each changes at most once over the CTA's tiles.  Lane L accumulates the 8
            // channels of 16-byte chunk (L & 7) of every 64-channel chunk over rows (L >> 3) + 4 i of its warp's 32, as packed
            // float32 pairs (FADD2 / FFMA2).  The four row groups, the four warps that share a column range and finally the
            // CTAs are folded at flush time only: one fp32 reduction per channel and quantity per CTA (per-tile or per-warp
            // atomics on the same few hundred addresses from every CTA serialise in L2 and hold back the grid's completion).
            constexpr int kChunks = kCols / 64;
            constexpr int kEpiThreads = L::kEpiWarps * 32;
            const uint32_t lch = static_cast<uint32_t>(lane & 7), lr = static_cast<uint32_t>(lane >> 3);
            uint64_t sacc[kChunks][8];              // [0..3]: first quantity of the chunk's channel pairs, [4..7]: second quantity
#pragma unroll
            for (int i = 0; i < kChunks; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) sacc[i][j] = 0ull;
            int stat_key = -1;
            auto flush_stats = [&](int key, bool final) {
                float *base = stats + static_cast<size_t>(key & 0xffff) * 2 * g.cout + static_cast<size_t>(key >> 16) * BLOCK_N;
                // scratch = the warp's own staging buffer as float [2 quantities][kCols]: its last TMA store must have read it
                if (lane == 0) bulk_wait_read0();
                __syncwarp();
#pragma unroll
                for (int i = 0; i < kChunks; ++i) {
                    float2 keep_s = make_float2(0.f, 0.f), keep_q = make_float2(0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float2 v = f32x2_unpack(sacc[i][j]);
                        v.x += __shfl_xor_sync(0xffffffffu, v.x, 8);  v.y += __shfl_xor_sync(0xffffffffu, v.y, 8);
                        v.x += __shfl_xor_sync(0xffffffffu, v.x, 16); v.y += __shfl_xor_sync(0xffffffffu, v.y, 16);
                        // all four row groups hold the totals now; lane L keeps channel pair (L >> 3) of its 16-byte chunk
                        if (static_cast<uint32_t>(j) == lr) keep_s = v;
                        if (static_cast<uint32_t>(j) == 4 + lr) keep_q = v;
                        sacc[i][j] = 0ull;
                    }
                    const uint32_t ch = static_cast<uint32_t>(i) * 64u + lch * 8u + 2u * lr;
                    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(stage_s + ch * 4u), "f"(keep_s.x), "f"(keep_s.y) : "memory");
                    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(stage_s + (static_cast<uint32_t>(kCols) + ch) * 4u), "f"(keep_q.x), "f"(keep_q.y) : "memory");
                }
                // fold the four warps that share a column range: one (coalesced) reduction per channel and quantity per CTA
                named_bar_sync(1, kEpiThreads);
                const uint32_t store_s = smem_u32(smem + L::kStoreOffset);
                for (int item = wq * 32 + lane; item < 2 * BLOCK_N; item += kEpiThreads) {
                    const int qty = item / BLOCK_N, col = item - qty * BLOCK_N;
                    const int hh = col / kCols, cc = col - hh * kCols;
                    const uint32_t a = store_s + static_cast<uint32_t>(hh * 4) * 4096u + static_cast<uint32_t>(qty * kCols + cc) * 4u;
                    const float v = (ld_shared_f32(a) + ld_shared_f32(a + 4096u)) + (ld_shared_f32(a + 8192u) + ld_shared_f32(a + 12288u));
                    atomicAdd(base + static_cast<size_t>(qty) * g.cout + col, v);
                }
                if (!final) named_bar_sync(1, kEpiThreads);      // the scratch is staging memory again
            };
            // bf16 pack of one 32-column accumulator segment of this lane's row (zeros where `keep` has no bit)
            auto pack32 = [](const uint32_t (&v)[32], uint32_t keep, uint32_t (&pk)[16]) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    __nv_bfloat162 b = __floats2bfloat162_rn(((keep >> j) & 1u) ? __uint_as_float(v[j]) : 0.f,
                                                             ((keep >> (j + 1)) & 1u) ? __uint_as_float(v[j + 1]) : 0.f);
                    pk[j >> 1] = *reinterpret_cast<uint32_t *>(&b);
                }
            };
            auto add_bf16x8 = [](uint32_t *v, const uint4 &a) {
                const uint32_t w4[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&w4[e]));
                    v[2 * e] = __float_as_uint(__uint_as_float(v[2 * e]) + f.x);
                    v[2 * e + 1] = __float_as_uint(__uint_as_float(v[2 * e + 1]) + f.y);
                }
            };
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = tile_lo; t < tile_hi; t += tile_step) {
                const TileCoord tc_ = decode_tile(t, n_tiles_n, g);
                const int n_blk = tc_.n_blk, tw = tc_.tw, th = tc_.th;
                const int img = tc_.img0 + pn, bimg = tc_.img0 + bpn0;
                const int oh = th * g.bh + ph, ow = tw * g.bw + pw;
                const bool valid = oh < g.oh && ow < g.ow && img < g.n;
                if (STATS || BNRED) {
                    // (the host only selects STATS / BNRED when all images of a tile belong to one statistics group)
                    const int key = (n_blk << 16) | (tc_.img0 / imgs_per_group);
                    if (key != stat_key) {
                        if (stat_key >= 0) flush_stats(stat_key, false);
                        stat_key = key;
                    }
                }
                const size_t pix_off = ((static_cast<size_t>(img) * g.oh + oh) * g.ow + ow) * g.cout + static_cast<size_t>(n_blk) * BLOCK_N;
                const bool has_add = addend != nullptr && valid;
                const uint4 *ap = reinterpret_cast<const uint4 *>(addend + pix_off + col_lo);
                uint4 cur[4], nxt[4];                        // the addend of one 32-channel segment of this lane's row, and the next one
                if (has_add) {
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) cur[q4] = __ldg(ap + q4);
                }
                uint32_t mbits[kCols / 32];                  // BNRED: ReLU mask of this lane's pixel, one bit per channel
                if (BNRED) {
                    // the BatchNorm input box of the first 64-channel chunk (the buffer is free: the previous tile's column
                    // sums ended with a proxy fence + __syncwarp) and the mask words do not depend on the accumulator: request them first
                    if (lane == 0) {
                        mbar_arrive_expect_tx(ybar + wq, 4096);
                        tma_load_4d(ybuf, &tmap_bn, ybar + wq, n_blk * BLOCK_N + col_lo, tw * g.bw + bpw0, th * g.bh + bph0, bimg);
                    }
                    // this lane's kCols mask bits are 16 (8) consecutive, equally aligned bytes: ONE load (four scalar ones at a lane stride
                    // of a pixel row quadruple the sector requests of the load pipe)
                    const unsigned char *mp = relu_mask + ((pix_off + col_lo) >> 3);
                    if (kCols == 128) {
                        const uint4 m4 = valid ? __ldg(reinterpret_cast<const uint4 *>(mp)) : make_uint4(0u, 0u, 0u, 0u);
                        mbits[0] = m4.x; mbits[1] = m4.y; mbits[kCols / 32 - 2] = m4.z; mbits[kCols / 32 - 1] = m4.w;
                    } else {
                        const uint2 m2 = valid ? __ldg(reinterpret_cast<const uint2 *>(mp)) : make_uint2(0u, 0u);
                        mbits[0] = m2.x; mbits[kCols / 32 - 1] = m2.y;
                    }
                }
                mbar_wait(tfull_bar + acc, acc_phase);
                tc_fence_after_sync();
                if (wq == 0 && lane == 0) { if (t == tile_lo) CONV_TRACE(5); CONV_TRACE(6); }
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N + col_lo);
#pragma unroll
                for (int c64 = 0; c64 < kCols; c64 += 64) {
#pragma unroll
                    for (int hseg = 0; hseg < 2; ++hseg) {
                        const int c = c64 + hseg * 32;               // column of this segment inside the warp's range
                        uint32_t v[32];
                        tmem_ld_32x32(taddr + static_cast<uint32_t>(c), v);
                        if (has_add && c + 32 < kCols) {
#pragma unroll
                            for (int q4 = 0; q4 < 4; ++q4) nxt[q4] = __ldg(ap + (c + 32) / 8 + q4);
                        }
                        tmem_ld_wait();
                        if (c + 32 == kCols) {
                            // the accumulator has been read completely: hand it back to the MMA warp before the stores
                            tc_fence_before_sync();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(tempty_bar + acc);
                        }
                        if (has_add) {
#pragma unroll
                            for (int q4 = 0; q4 < 4; ++q4) {
                                add_bf16x8(v + q4 * 8, cur[q4]);
                                cur[q4] = nxt[q4];
                            }
                        }
                        // rows outside the image are stored as zeros (the TMA store clips them; the statistics must not see them)
                        uint32_t pk[16];
                        const uint32_t keep = BNRED ? mbits[c / 32] : (valid ? 0xffffffffu : 0u);
                        if (!BNRED && keep == 0xffffffffu) {       // (a ReLU mask is hardly ever all ones)
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
                                pk[j >> 1] = *reinterpret_cast<uint32_t *>(&b);
                            }
                        } else {
                            pack32(v, keep, pk);
                        }
                        if (hseg == 0) {
                            // the previous TMA store of this warp must have finished reading the staging buffer
                            if (lane == 0) bulk_wait_read0();
                            __syncwarp();
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j)      // 128B swizzle: 16-byte chunk ^ (row & 7)
                            st_shared_v4(my_row_s + ((static_cast<uint32_t>(hseg * 4 + j) ^ sw) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_4d(&tmap_y, smem + L::kStoreOffset + wq * 4096, n_blk * BLOCK_N + col_lo + c64, tw * g.bw + bpw0, th * g.bh + bph0, bimg);
                        bulk_commit();
                    }
                    if (STATS) {
                        // column sums of the staged tile (bf16-rounded values, i.e. what BatchNorm will read back): 8 x LDS.128 per lane
#pragma unroll
                        for (int i8 = 0; i8 < 8; ++i8) {
                            const uint32_t r = lr + 4u * static_cast<uint32_t>(i8);
                            const uint4 wv = ld_shared_v4(stage_s + r * 128u + ((lch ^ (r & 7u)) << 4));
                            const uint32_t w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const uint64_t f = bf16x2_to_f32x2(w4[e]);
                                f32x2_add(sacc[c64 / 64][e], f);
                                f32x2_fma(sacc[c64 / 64][4 + e], f, f);
                            }
                        }
                    }
                    if (BNRED) {
                        // sum(dz) and sum(dz * y): dz from the staged tile, y from the TMA-loaded BatchNorm input box (same layout)
                        mbar_wait(ybar + wq, yphase);
                        yphase ^= 1;
#pragma unroll
                        for (int i8 = 0; i8 < 8; ++i8) {
                            const uint32_t r = lr + 4u * static_cast<uint32_t>(i8);
                            const uint32_t off = r * 128u + ((lch ^ (r & 7u)) << 4);
                            const uint4 dv = ld_shared_v4(stage_s + off), yv = ld_shared_v4(ybuf_s + off);
                            const uint32_t d4[4] = {dv.x, dv.y, dv.z, dv.w}, y4[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const uint64_t d = bf16x2_to_f32x2(d4[e]);
                                f32x2_add(sacc[c64 / 64][e], d);
                                f32x2_fma(sacc[c64 / 64][4 + e], d, bf16x2_to_f32x2(y4[e]));
                            }
                        }
                        // Every lane must HOLD its y values before the buffer is handed back to the TMA unit (async proxy): the shared
                        // loads above are only ISSUED at this point and ptxas schedules the products that consume them after the next
                        // box has been requested.  Observed with the addend path keeping the load/store pipe busy and the next box
                        // coming from L2: the box overtook the last loads of the previous one.  The proxy fence (MEMBAR + FENCE.VIEW.ASYNC)
                        // completes this thread's outstanding shared-memory reads and orders them before the async-proxy write.
                        fence_proxy_async();
                        __syncwarp();                                  // every lane is done with ybuf
                        if (c64 + 64 < kCols && lane == 0) {
                            mbar_arrive_expect_tx(ybar + wq, 4096);
                            tma_load_4d(ybuf, &tmap_bn, ybar + wq, n_blk * BLOCK_N + col_lo + c64 + 64, tw * g.bw + bpw0, th * g.bh + bph0, bimg);
                        }
                    }
                    if (wq == 0 && lane == 0) CONV_TRACE(16 + (c64 / 64) * 8 + 7);
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (wq == 0 && lane == 0) CONV_TRACE(7);
            if ((STATS || BNRED) && stat_key >= 0) flush_stats(stat_key, true);
            if (wq == 0 && lane == 0) CONV_TRACE(8);
            if (lane == 0) bulk_wait_read0();          // shared memory must outlive the last TMA store's reads
            if (wq == 0 && lane == 0) CONV_TRACE(9);
        }
    } else {
        // ===== epilogue (OUT_F32): TMEM -> registers (+ float32 addend) -> float32 global (NHWC) =====
        float *yf = reinterpret_cast<float *>(y);
        const float *addf = reinterpret_cast<const float *>(addend);
        const int q = warp & 3;                                        // TMEM lane quadrant of this warp
        const int col_lo = ((warp - 2) >> 2) * (BLOCK_N / 2), col_hi = col_lo + BLOCK_N / 2;    // this warp's half of the columns
        const int row = q * 32 + lane;
        const int pw = row % g.bw, ph = (row / g.bw) % g.bh, pn = row / (g.bw * g.bh);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = tile_lo; t < tile_hi; t += tile_step) {
            const TileCoord tc_ = decode_tile(t, n_tiles_n, g);
            const int img = tc_.img0 + pn;
            const int oh = tc_.th * g.bh + ph, ow = tc_.tw * g.bw + pw;
            const bool valid = oh < g.oh && ow < g.ow && img < g.n;
            const size_t pix_off = ((static_cast<size_t>(img) * g.oh + oh) * g.ow + ow) * g.cout + static_cast<size_t>(tc_.n_blk) * BLOCK_N;
            const bool has_add = addf != nullptr && valid;
            mbar_wait(tfull_bar + acc, acc_phase);
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
#pragma unroll 1
            for (int c = col_lo; c < col_hi; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + static_cast<uint32_t>(c), v);
                float4 ad[8];
                if (has_add) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) ad[j] = __ldg(reinterpret_cast<const float4 *>(addf + pix_off + c) + j);
                }
                tmem_ld_wait();
                if (valid) {
                    float4 *dst = reinterpret_cast<float4 *>(yf + pix_off + c);        // one whole 128-byte line per lane
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                               __uint_as_float(v[4 * j + 3]));
                        if (has_add) { o.x += ad[j].x; o.y += ad[j].y; o.z += ad[j].z; o.w += ad[j].w; }
                        dst[j] = o;
                    }
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar + acc);            // this warp's part of the accumulator is free
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) CONV_TRACE(10);
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 2 * BLOCK_N);
        if (lane == 0) CONV_TRACE(11);
    }
}

int make_tmap_x(CUtensorMap *m, const void *x, const ConvGeom &g) {
    if (!encode_nhwc(m, x, g.n, g.h, g.w, g.cin, g.bw, g.bh, g.stride, g.bn))
        return REGDA_ERR_CUDA;   /* text set by encode_bf16_sw128 (activations) */
    return REGDA_OK;
}

// fprop weights: [cout][ktot] row-major, box {64 k, block_n}
int make_tmap_w(CUtensorMap *m, const void *w, int cout, int ktot, int block_n) {
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(ktot), static_cast<cuuint64_t>(cout)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ktot) * 2};
    const cuuint32_t box[2] = {kBlockK, static_cast<cuuint32_t>(block_n)};
    const cuuint32_t estr[2] = {1, 1};
    if (!encode_bf16_sw128(m, w, 2, dims, strides, box, estr)) return REGDA_ERR_CUDA;   /* text set by encode_bf16_sw128 (weights) */
    return REGDA_OK;
}

// dgrad weights in place: memory [red][taps][wct >= out] (the forward conv's OHWI), box {64 out, 1 tap, 64 red}
int make_tmap_w_mn(CUtensorMap *m, const void *w, int red, int taps, int out, int wct) {
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(out), static_cast<cuuint64_t>(taps), static_cast<cuuint64_t>(red)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(wct) * 2, static_cast<cuuint64_t>(taps) * wct * 2};
    const cuuint32_t box[3] = {64, 1, 64};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (!encode_bf16_sw128(m, w, 3, dims, strides, box, estr)) return REGDA_ERR_CUDA;   /* text set by encode_bf16_sw128 (weights, MN) */
    return REGDA_OK;
}

struct BnRed {                       // BNRED launch: the BatchNorm whose output gradient this data-gradient launch produces
    const void *bn_y = nullptr;      // its input y, bf16 [n][h][w][c] (same shape as the gradient)
    const unsigned char *mask = nullptr;   // its ReLU mask, one bit per element
};

// `y` / `addend` are bf16, or float32 when OUT_F32
// regda_conv_hint_static_weights(): consumed by the next convolution this thread launches
thread_local int t_static_weights = 0;

template <int BLOCK_N, int STAGES, bool B_MN, bool STATS, bool OUT_F32, bool BNRED = false>
int launch_persistent_impl(const CUtensorMap &tx, const CUtensorMap &tw, void *y, const ConvGeom &g_in, cudaStream_t st,
                           float *stats, int imgs_per_group, const void *addend, const BnRed &br = BnRed()) {
    ConvGeom g = g_in;
    g.w_early = (t_static_weights && pdl_level() >= 1) ? 1 : 0;
    t_static_weights = 0;
    {
        const int nn = g.cout / BLOCK_N, nt = nn * g.tiles_img * g.tiles_h * g.tiles_w;
        static const int force = getenv("REGDA_CONV_CONTIG") ? atoi(getenv("REGDA_CONV_CONTIG")) : -1;      // sweeps / debugging only
        g.contig = force >= 0 ? (force != 0) : ((std::min(nt, sm_count()) % nn) != 0 ? 1 : 0);
    }
    using L = PersistSmem<BLOCK_N, STAGES, BNRED>;
    auto kern = conv_persistent_kernel<BLOCK_N, STAGES, B_MN, STATS, OUT_F32, BNRED>;
    const int smem = L::kTotal + 1024;
    static_assert(L::kTotal + 1024 <= 232448, "persistent conv kernel: shared memory over the 227 KB limit");
    REGDA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int n_tiles_n = g.cout / BLOCK_N;
    const int num_tiles = n_tiles_n * g.tiles_img * g.tiles_h * g.tiles_w;
    const int grid = std::min(num_tiles, sm_count());
    CUtensorMap ty = tx, tbn = tx;          // unused by the float32 epilogue / without BNRED
    if (!OUT_F32) {
        // output map [n][oh][ow][cout]; box = one epilogue warp's 32 tile rows x 64 channels: {sbw px, sbh rows, sbn images}
        const int sbw = std::min(g.bw, 32), sbh = std::min(g.bh, 32 / sbw), sbn = 32 / (sbw * sbh);
        if (!encode_nhwc(&ty, y, g.n, g.oh, g.ow, g.cout, sbw, sbh, 1, sbn)) return REGDA_ERR_CUDA;
        if (BNRED && !encode_nhwc(&tbn, br.bn_y, g.n, g.oh, g.ow, g.cout, sbw, sbh, 1, sbn)) return REGDA_ERR_CUDA;
    }
    REGDA_CUDA_CHECK(launch_pdl(kern, dim3(grid), dim3(kPersistThreads), smem, st, tx, tw, ty, tbn, static_cast<__nv_bfloat16 *>(y), g, n_tiles_n,
                                num_tiles, stats, imgs_per_group, static_cast<const __nv_bfloat16 *>(addend), br.mask));
    return REGDA_OK;
}

// 256-wide tiles are taken when they give at least half an SM-count of tiles (otherwise 128-wide tiles fill the machine better);
// regda_conv_tune() overrides the threshold (tile-policy sweeps: scripts/bench_conv.py --min-tiles-256)
int g_min_tiles_256 = 0;
int min_tiles_256() { return g_min_tiles_256 > 0 ? g_min_tiles_256 : sm_count() / 2; }

int pick_block_n(const ConvGeom &g) {
    const long long m_tiles = static_cast<long long>(g.tiles_img) * g.tiles_h * g.tiles_w;
    if (g.cout % 256 == 0 && m_tiles * (g.cout / 256) >= min_tiles_256()) return 256;
    if (g.cout % 128 == 0) return 128;
    return 64;
}

// data gradient + the reductions of the BatchNorm backward that consumes it (BNRED); `red` [groups][2][g.cout]
int launch_dgrad_bnred(const void *act, const void *wgt, __nv_bfloat16 *out, const ConvGeom &g, int taps, cudaStream_t st,
                       float *red, int imgs_per_group, const __nv_bfloat16 *addend, const BnRed &br) {
    CUtensorMap tx, tw;
    int rc = make_tmap_x(&tx, act, g);
    if (rc) return rc;
    const int block_n = pick_block_n(g);
    rc = make_tmap_w_mn(&tw, wgt, g.cin, taps, g.cout, g.wct);
    if (rc) return rc;
    // one pipeline stage fewer than the plain kernel at 256 / 128: the 32 KB of BatchNorm-input boxes take their place
    if (block_n == 256) return launch_persistent_impl<256, 3, true, false, false, true>(tx, tw, out, g, st, red, imgs_per_group, addend, br);
    if (block_n == 128) return launch_persistent_impl<128, 5, true, false, false, true>(tx, tw, out, g, st, red, imgs_per_group, addend, br);
    return launch_persistent_impl<64, 6, true, false, false, true>(tx, tw, out, g, st, red, imgs_per_group, addend, br);
}

template <int BLOCK_N, int STAGES, bool B_MN>
int launch_persistent(const CUtensorMap &tx, const CUtensorMap &tw, void *y, const ConvGeom &g, cudaStream_t st, float *stats,
                      int imgs_per_group, const void *addend, bool out_f32) {
    if (out_f32) return launch_persistent_impl<BLOCK_N, STAGES, B_MN, false, true>(tx, tw, y, g, st, nullptr, 1, addend);
    if (!B_MN && stats != nullptr) return launch_persistent_impl<BLOCK_N, STAGES, false, true, false>(tx, tw, y, g, st, stats, imgs_per_group, addend);
    return launch_persistent_impl<BLOCK_N, STAGES, B_MN, false, false>(tx, tw, y, g, st, nullptr, 1, addend);
}

template <bool B_MN>
int launch_conv(const void *act, const void *wgt, void *out, const ConvGeom &g, int taps, cudaStream_t st, float *stats = nullptr,
                int imgs_per_group = 1, const void *addend = nullptr, bool out_f32 = false) {
    // g.cin = reduction channels, g.cout = output channels of THIS GEMM (already swapped for dgrad)
    CUtensorMap tx, tw;
    int rc = make_tmap_x(&tx, act, g);
    if (rc) return rc;
    const int block_n = pick_block_n(g);
    rc = B_MN ? make_tmap_w_mn(&tw, wgt, g.cin, taps, g.cout, g.wct) : make_tmap_w(&tw, wgt, g.cout, taps * g.wct, block_n);
    if (rc) return rc;
    if (block_n == 256) return launch_persistent<256, 4, B_MN>(tx, tw, out, g, st, stats, imgs_per_group, addend, out_f32);
    if (block_n == 128) return launch_persistent<128, 6, B_MN>(tx, tw, out, g, st, stats, imgs_per_group, addend, out_f32);
    return launch_persistent<64, 8, B_MN>(tx, tw, out, g, st, stats, imgs_per_group, addend, out_f32);
}

int geom_init(ConvGeom &g, int n, int h, int w, int cin, int cout, int r, int s, int stride, int pad, int dil) {
    g.n = n; g.h = h; g.w = w; g.cin = cin; g.cout = cout; g.r = r; g.s = s; g.pad = pad; g.dil = dil; g.stride = stride; g.flip = 0;
    g.oh = (h + 2 * pad - dil * (r - 1) - 1) / stride + 1;
    g.ow = (w + 2 * pad - dil * (s - 1) - 1) / stride + 1;
    int bw = 1;
    while (bw * 2 <= g.ow && bw * 2 <= kBlockM) bw *= 2;         // largest power of two <= min(ow, 128)
    int bh = kBlockM / bw, bn = 1;
    if (static_cast<long long>(g.oh) * g.ow < kBlockM) {
        // small map: the patch does not reach 128 pixels inside one image -- fill the M tile with several images
        bh = 1;
        while (bh * 2 <= g.oh && bw * bh * 2 <= kBlockM) bh *= 2;
        bn = kBlockM / (bw * bh);
    }
    g.bw = bw; g.bh = bh; g.bn = bn;
    g.tiles_w = (g.ow + g.bw - 1) / g.bw;
    g.tiles_h = (g.oh + g.bh - 1) / g.bh;
    g.tiles_img = (n + bn - 1) / bn;
    g.kc = cin / kBlockK;
    g.wct = cin;
    g.w_early = 0;
    g.contig = 0;
    return REGDA_OK;
}

bool shape_ok(int n, int h, int w, int cin, int cout, int r, int s, int stride, int pad, int dil) {
    if (n < 1 || h < 1 || w < 1 || stride < 1 || stride > 2 || r < 1 || s < 1 || r * s > 49 || dil < 1 || pad < 0) return false;
    if (cin % 64 != 0 || cout % 64 != 0 || cin < 64 || cout < 64) return false;
    if (h + 2 * pad < dil * (r - 1) + 1 || w + 2 * pad < dil * (s - 1) + 1) return false;
    return true;
}

// all images of one M tile must fall into one BatchNorm statistics group (fused statistics / fused backward reductions)
bool groups_ok(const ConvGeom &g, int groups) {
    if (groups < 1 || g.n % groups != 0) return false;
    return g.bn == 1 || groups == 1 || (g.n / groups) % g.bn == 0;
}

bool aligned16(const void *a, const void *b, const void *c) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}

// geometry of the data gradient as a stride-1 convolution over dy [n][oh][ow][cout] producing [n][h][w][cin]
bool dgrad_geom(ConvGeom &g, int n, int h, int w, int cin, int cout, int r, int s, int pad, int dil) {
    const int oh = h + 2 * pad - dil * (r - 1), ow = w + 2 * pad - dil * (s - 1);
    geom_init(g, n, oh, ow, cout, cin, r, s, 1, dil * (r - 1) - pad, dil);
    g.flip = 1;
    g.wct = cin;                     // MN-major weight rows are the forward convolution's input channels
    return g.oh == h && g.ow == w;
}

}  // namespace
}  // namespace regda

using namespace regda;

// Hint for the NEXT forward / data-gradient convolution launched by the calling thread: its weight tensor was last written long
// before the kernel that precedes the launch in its stream (true for the bf16 weight copies of a training step, which the optimizer
// kernel writes once per step; NOT true for a weight-like operand produced by the kernel just before).  The kernel then requests the
// weight halves of its first pipeline stages before its programmatic-dependent-launch wait.  Consumed by that launch.
extern "C" int regda_conv_hint_static_weights(void) {
    t_static_weights = 1;
    return REGDA_OK;
}

// tile-policy knob for sweeps: minimum number of 128 x 256 tiles for the 256-wide tile to be chosen (0 = default, SM count / 2)
extern "C" int regda_conv_tune(int min_tiles_256_value) {
    g_min_tiles_256 = min_tiles_256_value > 0 ? min_tiles_256_value : 0;
    return REGDA_OK;
}

// 1 if (shape, alignment) is covered by the tcgen05 kernel: channel counts multiples of 64, stride 1 or 2, any map size
extern "C" int regda_conv_fprop_supported(int n, int h, int w, int cin, int cout, int r, int s, int stride, int pad, int dil) {
    return shape_ok(n, h, w, cin, cout, r, s, stride, pad, dil) ? 1 : 0;
}

// 1 if the forward kernel can also produce the BatchNorm statistics of its output for `groups` statistics groups
extern "C" int regda_conv_fprop_stats_supported(int n, int h, int w, int cin, int cout, int r, int s, int stride, int pad, int dil, int groups) {
    if (!shape_ok(n, h, w, cin, cout, r, s, stride, pad, dil)) return 0;
    ConvGeom g;
    geom_init(g, n, h, w, cin, cout, r, s, stride, pad, dil);
    return groups_ok(g, groups) ? 1 : 0;
}

extern "C" int regda_conv_fprop_bf16(const void *x, const void *wgt, void *y, int n, int h, int w, int cin, int cout,
                                     int r, int s, int stride, int pad, int dil, void *stream) {
    if (!regda_conv_fprop_supported(n, h, w, cin, cout, r, s, stride, pad, dil))
        return fail(REGDA_ERR_UNSUPPORTED, "conv_fprop: shape not covered by the tcgen05 kernel");
    if (!x || !wgt || !y) return fail(REGDA_ERR_INVALID_ARG, "conv_fprop: null pointer");
    if (!aligned16(x, wgt, y)) return fail(REGDA_ERR_INVALID_ARG, "conv_fprop: tensors must be 16-byte aligned");
    ConvGeom g;
    geom_init(g, n, h, w, cin, cout, r, s, stride, pad, dil);
    ensure_context(x);
    return launch_conv<false>(x, wgt, y, g, r * s, static_cast<cudaStream_t>(stream));
}

// Same convolution with the raw float32 accumulators as output: y float32 [n][oh][ow][cout].
extern "C" int regda_conv_fprop_bf16_f32out(const void *x, const void *wgt, float *y, int n, int h, int w, int cin, int cout,
                                            int r, int s, int stride, int pad, int dil, void *stream) {
    if (!regda_conv_fprop_supported(n, h, w, cin, cout, r, s, stride, pad, dil))
        return fail(REGDA_ERR_UNSUPPORTED, "conv_fprop_f32out: shape not covered by the tcgen05 kernel");
    if (!x || !wgt || !y) return fail(REGDA_ERR_INVALID_ARG, "conv_fprop_f32out: null pointer");
    if (!aligned16(x, wgt, y)) return fail(REGDA_ERR_INVALID_ARG, "conv_fprop_f32out: tensors must be 16-byte aligned");
    ConvGeom g;
    geom_init(g, n, h, w, cin, cout, r, s, stride, pad, dil);
    ensure_context(x);
    return launch_conv<false>(x, wgt, y, g, r * s, static_cast<cudaStream_t>(stream), nullptr, 1, nullptr, true);
}

// Forward convolution that also produces the train-mode BatchNorm statistics of its output:
// bn_stats float32 [groups][2][cout] = per-group per-channel (sum, sum of squares) over the group's n/groups images,
// written (zeroed here first).  Feed it to regda_bn_forward_bf16(..., have_stats = 1).
extern "C" int regda_conv_fprop_stats_bf16(const void *x, const void *wgt, void *y, int n, int h, int w, int cin, int cout,
                                           int r, int s, int stride, int pad, int dil, float *bn_stats, int groups, int stats_zeroed,
                                           void *stream) {
    if (!regda_conv_fprop_stats_supported(n, h, w, cin, cout, r, s, stride, pad, dil, groups))
        return fail(REGDA_ERR_UNSUPPORTED, "conv_fprop_stats: shape / statistics groups not covered by the tcgen05 kernel");
    if (!x || !wgt || !y || !bn_stats) return fail(REGDA_ERR_INVALID_ARG, "conv_fprop_stats: null pointer");
    if (!aligned16(x, wgt, y)) return fail(REGDA_ERR_INVALID_ARG, "conv_fprop: tensors must be 16-byte aligned");
    ConvGeom g;
    geom_init(g, n, h, w, cin, cout, r, s, stride, pad, dil);
    ensure_context(x);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!stats_zeroed) REGDA_CUDA_CHECK(cudaMemsetAsync(bn_stats, 0, static_cast<size_t>(groups) * 2 * cout * sizeof(float), st));
    return launch_conv<false>(x, wgt, y, g, r * s, st, bn_stats, n / groups);
}

// Forward convolution whose epilogue adds `addend` (bf16, the output's shape) before rounding / taking the statistics;
// bn_stats == NULL: no statistics.  (The folded PPM fuse convolution hands the pyramid branches' contribution in this way.)
extern "C" int regda_conv_fprop_addend_bf16(const void *x, const void *wgt, int wct, void *y, int n, int h, int w, int cin, int cout,
                                            int r, int s, int stride, int pad, int dil, const void *addend, float *bn_stats, int groups,
                                            int stats_zeroed, void *stream) {
    if (wct < cin || wct % 8) return fail(REGDA_ERR_INVALID_ARG, "conv_fprop_addend: weight channel count must be >= cin and a multiple of 8");
    if (!regda_conv_fprop_supported(n, h, w, cin, cout, r, s, stride, pad, dil))
        return fail(REGDA_ERR_UNSUPPORTED, "conv_fprop_addend: shape not covered by the tcgen05 kernel");
    if (bn_stats && !regda_conv_fprop_stats_supported(n, h, w, cin, cout, r, s, stride, pad, dil, groups))
        return fail(REGDA_ERR_UNSUPPORTED, "conv_fprop_addend: statistics groups not covered by the tcgen05 kernel");
    if (!x || !wgt || !y) return fail(REGDA_ERR_INVALID_ARG, "conv_fprop_addend: null pointer");
    if (!aligned16(x, wgt, y) || (addend && (reinterpret_cast<uintptr_t>(addend) & 15)))
        return fail(REGDA_ERR_INVALID_ARG, "conv_fprop_addend: tensors must be 16-byte aligned");
    ConvGeom g;
    geom_init(g, n, h, w, cin, cout, r, s, stride, pad, dil);
    g.wct = wct;
    ensure_context(x);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (bn_stats && !stats_zeroed) REGDA_CUDA_CHECK(cudaMemsetAsync(bn_stats, 0, static_cast<size_t>(groups) * 2 * cout * sizeof(float), st));
    return launch_conv<false>(x, wgt, y, g, r * s, st, bn_stats, bn_stats ? n / groups : 1, addend);
}

// Data gradient of a STRIDE-1 convolution, reading the forward weights IN PLACE:
//   dx[n][h][w][cin] = sum_{r,s,co} dy[n][y + pad - r*dil][x + pad - s*dil][co] * wgt[co][r][s][cin]
// dy bf16 [n][oh][ow][cout], wgt bf16 [cout][r][s][cin] (OHWI), dx bf16 [n][h][w][cin].
// (A stride-2 convolution's data gradient is this call on the zero-inserted dy: regda_zero_insert2_bf16.)
extern "C" int regda_conv_dgrad_supported(int n, int h, int w, int cin, int cout, int r, int s, int stride, int pad, int dil) {
    if (stride != 1 || r != s) return 0;
    const int pad2 = dil * (r - 1) - pad;
    if (pad2 < 0) return 0;
    const int oh = h + 2 * pad - dil * (r - 1), ow = w + 2 * pad - dil * (s - 1);
    if (oh < 1 || ow < 1) return 0;
    return regda_conv_fprop_supported(n, oh, ow, cout, cin, r, s, 1, pad2, dil);
}

extern "C" int regda_conv_dgrad_bnred_supported(int n, int h, int w, int cin, int cout, int r, int s, int stride, int pad, int dil, int groups) {
    if (!regda_conv_dgrad_supported(n, h, w, cin, cout, r, s, stride, pad, dil)) return 0;
    ConvGeom g;
    if (!dgrad_geom(g, n, h, w, cin, cout, r, s, pad, dil)) return 0;
    return groups_ok(g, groups) ? 1 : 0;
}

extern "C" int regda_conv_dgrad_bf16(const void *dy, const void *wgt, void *dx, int n, int h, int w, int cin, int cout,
                                     int r, int s, int stride, int pad, int dil, const void *addend, void *stream) {
    if (!regda_conv_dgrad_supported(n, h, w, cin, cout, r, s, stride, pad, dil))
        return fail(REGDA_ERR_UNSUPPORTED, "conv_dgrad: shape not covered by the tcgen05 kernel");
    if (!dy || !wgt || !dx) return fail(REGDA_ERR_INVALID_ARG, "conv_dgrad: null pointer");
    if (!aligned16(dy, wgt, dx)) return fail(REGDA_ERR_INVALID_ARG, "conv_dgrad: tensors must be 16-byte aligned");
    ConvGeom g;
    if (!dgrad_geom(g, n, h, w, cin, cout, r, s, pad, dil)) return fail(REGDA_ERR_INVALID_ARG, "conv_dgrad: inconsistent geometry");
    ensure_context(dy);
    if (addend != nullptr && (reinterpret_cast<uintptr_t>(addend) & 15)) return fail(REGDA_ERR_INVALID_ARG, "conv_dgrad: addend must be 16-byte aligned");
    return launch_conv<true>(dy, wgt, dx, g, r * s, static_cast<cudaStream_t>(stream), nullptr, 1, addend);
}

// Data gradient with respect to the FIRST cin input channels of a convolution whose weight has wct >= cin channels per tap
// (wgt bf16 [cout][r][s][wct], read in place): dx bf16 [n][h][w][cin].
extern "C" int regda_conv_dgrad_wslice_bf16(const void *dy, const void *wgt, int wct, void *dx, int n, int h, int w, int cin, int cout,
                                            int r, int s, int stride, int pad, int dil, const void *addend, void *stream) {
    if (!regda_conv_dgrad_supported(n, h, w, cin, cout, r, s, stride, pad, dil))
        return fail(REGDA_ERR_UNSUPPORTED, "conv_dgrad_wslice: shape not covered by the tcgen05 kernel");
    if (!dy || !wgt || !dx || wct < cin || wct % 8) return fail(REGDA_ERR_INVALID_ARG, "conv_dgrad_wslice: bad arguments");
    if (!aligned16(dy, wgt, dx) || (addend && (reinterpret_cast<uintptr_t>(addend) & 15)))
        return fail(REGDA_ERR_INVALID_ARG, "conv_dgrad_wslice: tensors must be 16-byte aligned");
    ConvGeom g;
    if (!dgrad_geom(g, n, h, w, cin, cout, r, s, pad, dil)) return fail(REGDA_ERR_INVALID_ARG, "conv_dgrad_wslice: inconsistent geometry");
    g.wct = wct;
    ensure_context(dy);
    return launch_conv<true>(dy, wgt, dx, g, r * s, static_cast<cudaStream_t>(stream), nullptr, 1, addend);
}

// float32 output (and float32 addend): the raw accumulators of the same data gradient
extern "C" int regda_conv_dgrad_bf16_f32out(const void *dy, const void *wgt, float *dx, int n, int h, int w, int cin, int cout,
                                            int r, int s, int stride, int pad, int dil, const float *addend, void *stream) {
    if (!regda_conv_dgrad_supported(n, h, w, cin, cout, r, s, stride, pad, dil))
        return fail(REGDA_ERR_UNSUPPORTED, "conv_dgrad_f32out: shape not covered by the tcgen05 kernel");
    if (!dy || !wgt || !dx) return fail(REGDA_ERR_INVALID_ARG, "conv_dgrad_f32out: null pointer");
    if (!aligned16(dy, wgt, dx)) return fail(REGDA_ERR_INVALID_ARG, "conv_dgrad_f32out: tensors must be 16-byte aligned");
    ConvGeom g;
    if (!dgrad_geom(g, n, h, w, cin, cout, r, s, pad, dil)) return fail(REGDA_ERR_INVALID_ARG, "conv_dgrad_f32out: inconsistent geometry");
    ensure_context(dy);
    if (addend != nullptr && (reinterpret_cast<uintptr_t>(addend) & 15)) return fail(REGDA_ERR_INVALID_ARG, "conv_dgrad_f32out: addend must be 16-byte aligned");
    return launch_conv<true>(dy, wgt, dx, g, r * s, static_cast<cudaStream_t>(stream), nullptr, 1, addend, true);
}

// Data gradient fused with the first half of the backward of the BatchNorm(+ReLU) whose OUTPUT is this convolution's input:
// dx receives dz = (dgrad + addend) masked by relu_mask (bit e of the mask = [BatchNorm output element e > 0], written by
// regda_bn_forward_bf16), and red float32 [groups][2][cin] ACCUMULATES sum(dz) and sum(dz * bn_y) per statistics group and
// channel (bn_y = that BatchNorm's input, bf16 [n][h][w][cin]).  regda_bn_backward_bf16(..., dz_ready = 1) then only applies.
extern "C" int regda_conv_dgrad_bnred_bf16(const void *dy, const void *wgt, void *dx, int n, int h, int w, int cin, int cout,
                                           int r, int s, int stride, int pad, int dil, const void *addend, const void *bn_y,
                                           const void *relu_mask, float *red, int groups, void *stream) {
    if (!regda_conv_dgrad_bnred_supported(n, h, w, cin, cout, r, s, stride, pad, dil, groups))
        return fail(REGDA_ERR_UNSUPPORTED, "conv_dgrad_bnred: shape / statistics groups not covered by the tcgen05 kernel");
    if (!dy || !wgt || !dx || !bn_y || !relu_mask || !red) return fail(REGDA_ERR_INVALID_ARG, "conv_dgrad_bnred: null pointer");
    if (!aligned16(dy, wgt, dx) || !aligned16(bn_y, relu_mask, nullptr))
        return fail(REGDA_ERR_INVALID_ARG, "conv_dgrad_bnred: tensors must be 16-byte aligned");
    if (addend != nullptr && (reinterpret_cast<uintptr_t>(addend) & 15)) return fail(REGDA_ERR_INVALID_ARG, "conv_dgrad_bnred: addend must be 16-byte aligned");
    ConvGeom g;
    if (!dgrad_geom(g, n, h, w, cin, cout, r, s, pad, dil)) return fail(REGDA_ERR_INVALID_ARG, "conv_dgrad_bnred: inconsistent geometry");
    ensure_context(dy);
    BnRed br;
    br.bn_y = bn_y;
    br.mask = static_cast<const unsigned char *>(relu_mask);
    return launch_dgrad_bnred(dy, wgt, static_cast<__nv_bfloat16 *>(dx), g, r * s, static_cast<cudaStream_t>(stream), red, n / groups,
                              static_cast<const __nv_bfloat16 *>(addend), br);
}
