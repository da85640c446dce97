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
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 2-5 = epilogue
//     (tcgen05.ld -> bf16 -> global), STAGES-deep mbarrier ring between producer and MMA.
// The same kernel computes the data gradient of a stride-1 convolution when it is given dY and
// the flipped / transposed weights (host side prepares them).
#include "common.cuh"
#include "tc_common.cuh"

namespace regda {
namespace {

using namespace tc;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;          // bf16: 128 bytes = one swizzle span
constexpr int kUmmaK = 16;
constexpr int kThreads = 192;        // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue

struct ConvGeom {
    int n, h, w, cin, cout;          // input image, channels
    int oh, ow;                      // output image
    int r, s, pad, dil;
    int bh, bw;                      // spatial patch of one M tile (bh*bw == 128)
    int tiles_h, tiles_w;            // patches per image
    int kc;                          // cin / 64
};

template <int BLOCK_N, int STAGES>
struct SmemLayout {
    static constexpr int kABytes = kBlockM * kBlockK * 2;
    static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOffset = STAGES * kStageBytes;
    static constexpr int kTotal = kBarOffset + (2 * STAGES + 1) * 8 + 8;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kThreads, 2)
conv_fprop_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                  __nv_bfloat16 *__restrict__ y, const ConvGeom g) {
    using L = SmemLayout<BLOCK_N, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled tiles must start on a 1024-byte boundary of the shared window
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + L::kBarOffset);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *accum_bar = empty_bar + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // tile coordinates: blockIdx.x = N tile (fastest, so CTAs sharing an A tile run together), blockIdx.y = M tile
    const int n_blk = blockIdx.x;
    int m_blk = blockIdx.y;
    const int tw = m_blk % g.tiles_w; m_blk /= g.tiles_w;
    const int th = m_blk % g.tiles_h;
    const int img = m_blk / g.tiles_h;
    const int oh0 = th * g.bh, ow0 = tw * g.bw;
    const int num_k = g.r * g.s * g.kc;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_w);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 2) tmem_alloc(tmem_slot, BLOCK_N);          // power of two >= 32 columns
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < num_k; ++kb) {
                const int tap = kb / g.kc, c0 = (kb - tap * g.kc) * kBlockK;
                const int fr = tap / g.s, fs = tap - fr * g.s;
                mbar_wait(empty_bar + stage, phase ^ 1);
                uint8_t *sa = smem + stage * L::kStageBytes;
                uint8_t *sb = sa + L::kABytes;
                mbar_arrive_expect_tx(full_bar + stage, L::kStageBytes);
                tma_load_4d(sa, &tmap_x, full_bar + stage, c0, ow0 + fs * g.dil - g.pad, oh0 + fr * g.dil - g.pad, img);
                tma_load_2d(sb, &tmap_w, full_bar + stage, tap * g.cin + c0, n_blk * BLOCK_N);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < num_k; ++kb) {
                mbar_wait(full_bar + stage, phase);
                tc_fence_after_sync();
                const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
                const uint32_t sb = sa + L::kABytes;
                const uint64_t adesc = make_smem_desc(sa, 0, 1024);
                const uint64_t bdesc = make_smem_desc(sb, 0, 1024);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                    // advance the start address by k * 16 elements * 2 B = 32 B (>> 4 = 2) inside the swizzle span
                    umma_bf16(tmem_base, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                              (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(empty_bar + stage);          // frees the smem stage when these MMAs retire
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(accum_bar);                       // accumulator complete
        }
    } else {
        // ===== epilogue: TMEM -> registers -> bf16 -> global (NHWC) =====
        const int q = warp & 3;                           // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;                    // pixel index inside the patch
        const int ph = row / g.bw, pw = row - ph * g.bw;
        const int oh = oh0 + ph, ow = ow0 + pw;
        const bool valid = oh < g.oh && ow < g.ow;
        __nv_bfloat16 *dst = y + ((static_cast<size_t>(img) * g.oh + oh) * g.ow + ow) * g.cout + static_cast<size_t>(n_blk) * BLOCK_N;
        mbar_wait(accum_bar, 0);
        tc_fence_after_sync();
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c), v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    uint4 pk;
                    __nv_bfloat162 b0 = __floats2bfloat162_rn(__uint_as_float(v[j + 0]), __uint_as_float(v[j + 1]));
                    __nv_bfloat162 b1 = __floats2bfloat162_rn(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                    __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
                    __nv_bfloat162 b3 = __floats2bfloat162_rn(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
                    pk.x = *reinterpret_cast<uint32_t *>(&b0); pk.y = *reinterpret_cast<uint32_t *>(&b1);
                    pk.z = *reinterpret_cast<uint32_t *>(&b2); pk.w = *reinterpret_cast<uint32_t *>(&b3);
                    *reinterpret_cast<uint4 *>(dst + c + j) = pk;
                }
            }
        }
        tc_fence_before_sync();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, BLOCK_N);
    }
}

int make_tmap_x(CUtensorMap *m, const void *x, const ConvGeom &g) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(REGDA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(g.cin), static_cast<cuuint64_t>(g.w), static_cast<cuuint64_t>(g.h), static_cast<cuuint64_t>(g.n)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(g.cin) * 2, static_cast<cuuint64_t>(g.w) * g.cin * 2,
                                   static_cast<cuuint64_t>(g.h) * g.w * g.cin * 2};
    const cuuint32_t box[4] = {kBlockK, static_cast<cuuint32_t>(g.bw), static_cast<cuuint32_t>(g.bh), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(x), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail(REGDA_ERR_CUDA, "cuTensorMapEncodeTiled(activations) failed with %d", static_cast<int>(rc));
    return REGDA_OK;
}

int make_tmap_w(CUtensorMap *m, const void *w, int cout, int ktot, int block_n) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(REGDA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(ktot), static_cast<cuuint64_t>(cout)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ktot) * 2};
    const cuuint32_t box[2] = {kBlockK, static_cast<cuuint32_t>(block_n)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(w), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail(REGDA_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed with %d", static_cast<int>(rc));
    return REGDA_OK;
}

template <int BLOCK_N, int STAGES>
int launch_fprop(const CUtensorMap &tx, const CUtensorMap &tw, __nv_bfloat16 *y, const ConvGeom &g, cudaStream_t st) {
    using L = SmemLayout<BLOCK_N, STAGES>;
    auto kern = conv_fprop_kernel<BLOCK_N, STAGES>;
    const int smem = L::kTotal + 1024;     // slack for the 1024-byte alignment of the dynamic segment
    REGDA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const dim3 grid(g.cout / BLOCK_N, g.n * g.tiles_h * g.tiles_w);
    kern<<<grid, kThreads, smem, st>>>(tx, tw, y, g);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}

}  // namespace
}  // namespace regda

using namespace regda;

// 1 if (shape, alignment) is covered by the tcgen05 kernel
extern "C" int regda_conv_fprop_supported(int n, int h, int w, int cin, int cout, int r, int s, int stride, int pad, int dil) {
    if (n < 1 || h < 1 || w < 1 || stride != 1 || r < 1 || s < 1 || r * s > 49 || dil < 1 || pad < 0) return 0;
    if (cin % 64 != 0 || cout % 64 != 0) return 0;
    const int oh = h + 2 * pad - dil * (r - 1), ow = w + 2 * pad - dil * (s - 1);
    if (oh < 1 || ow < 1) return 0;
    if (static_cast<long long>(oh) * ow < 128) return 0;          // tiny maps (PPM branches) stay on the library path
    return 1;
}

extern "C" int regda_conv_fprop_bf16(const void *x, const void *wgt, void *y, int n, int h, int w, int cin, int cout,
                                     int r, int s, int stride, int pad, int dil, void *stream) {
    if (!regda_conv_fprop_supported(n, h, w, cin, cout, r, s, stride, pad, dil))
        return fail(REGDA_ERR_UNSUPPORTED, "conv_fprop: shape not covered by the tcgen05 kernel");
    if (!x || !wgt || !y) return fail(REGDA_ERR_INVALID_ARG, "conv_fprop: null pointer");
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wgt) | reinterpret_cast<uintptr_t>(y)) & 15)
        return fail(REGDA_ERR_INVALID_ARG, "conv_fprop: tensors must be 16-byte aligned");
    ConvGeom g;
    g.n = n; g.h = h; g.w = w; g.cin = cin; g.cout = cout; g.r = r; g.s = s; g.pad = pad; g.dil = dil;
    g.oh = h + 2 * pad - dil * (r - 1);
    g.ow = w + 2 * pad - dil * (s - 1);
    int bw = 1;
    while (bw * 2 <= g.ow && bw * 2 <= 128) bw *= 2;             // largest power of two <= min(ow, 128)
    g.bw = bw; g.bh = kBlockM / bw;
    g.tiles_w = (g.ow + g.bw - 1) / g.bw;
    g.tiles_h = (g.oh + g.bh - 1) / g.bh;
    g.kc = cin / kBlockK;
    CUtensorMap tx, tw;
    int rc = make_tmap_x(&tx, x, g);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    __nv_bfloat16 *yy = static_cast<__nv_bfloat16 *>(y);
    if (cout % 128 == 0) {
        rc = make_tmap_w(&tw, wgt, cout, r * s * cin, 128);
        if (rc) return rc;
        return launch_fprop<128, 3>(tx, tw, yy, g, st);
    }
    rc = make_tmap_w(&tw, wgt, cout, r * s * cin, 64);
    if (rc) return rc;
    return launch_fprop<64, 4>(tx, tw, yy, g, st);
}
