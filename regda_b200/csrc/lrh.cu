// Local Region Homogenizing (LRH) for sm_100a.
//
// Replaces Homogenizer.forward (reference regda/utils/local_region_homog.py:125-152):
// per-image per-region class histogram of the hard pseudo labels, float32 majority-ratio test,
// gather of the winning class back to every pixel of the region.
//
// HBM-bound integer work: 24 algorithmic bytes per pixel (labels 8 + regions 8 + out 8).
//
// Cluster path (the fast one): one thread-block cluster per image.  Each CTA streams 1/CL of
// the image ONCE from HBM with 256-bit loads, keeps a compressed copy (u8 label code + u16
// region id = 3 B/px) in shared memory, and histograms into its own shared-memory bins
// (two u16 counters per 32-bit word, per-thread run aggregation over 8 consecutive pixels
// so that one shared atomic usually covers 8 pixels).  Bins are merged across the cluster
// through distributed shared memory, the owner CTA of each region applies the exact float32
// test and broadcasts the winner byte to all CTAs, and the output is produced from the
// shared-memory copy -- DRAM traffic is exactly the algorithmic 24 B/px.
//
// Generic path: global-memory bins (warp-aggregated atomics), winners, apply; any region
// bound / image size / alignment, 40 B/px worst case.
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace regda {
namespace {

constexpr int kGroupPx = 8;            // consecutive pixels per thread per iteration
constexpr unsigned kWinNone = 0xFFu;   // region did not pass (or winner == ignore value)
constexpr unsigned kCodeIgnore = 0xFFu;
constexpr unsigned kCodeClassNum = 0xFEu;  // label == class_num: counts nowhere, passes through
constexpr int kMaxFastClasses = 16;

struct LrhArgs {
    const long long *labels;
    const long long *regions;
    long long *out;
    int b;
    int hw;
    int class_num;
    long long ignore_label;
    float percent;
    int region_bound;
    int32_t *flags;
};

__device__ __forceinline__ unsigned encode_label(long long l, const LrhArgs &a, bool &bad) {
    if (l == a.ignore_label) return kCodeIgnore;  // local_region_homog.py:118 (remap happens first)
    if (static_cast<unsigned long long>(l) < static_cast<unsigned long long>(a.class_num)) return static_cast<unsigned>(l);
    if (l == a.class_num) return kCodeClassNum;
    bad = true;                                    // one_hot would raise (:121)
    return kCodeIgnore;
}

__device__ __forceinline__ long long decode_label(unsigned code, const LrhArgs &a) {
    return code == kCodeIgnore ? a.ignore_label : (code == kCodeClassNum ? static_cast<long long>(a.class_num) : static_cast<long long>(code));
}

__device__ __forceinline__ unsigned checked_region(long long r, const LrhArgs &a, bool &bad) {
    if (static_cast<unsigned long long>(r) >= static_cast<unsigned long long>(a.region_bound)) {
        bad = true;                                // scatter index out of range (:140)
        return 0u;
    }
    return static_cast<unsigned>(r);
}

// local_region_homog.py:141-144.  cnt[c] are the per-class pixel counts of one region.
// ratio = float32(max) / (float32(valid) + 1e-5f), rejected iff ratio < float32(percent);
// first maximum wins ties.  Explicit _rn intrinsics: no contraction, IEEE division.
template <int NC>
__device__ __forceinline__ unsigned lrh_winner(const unsigned (&cnt)[NC], int class_num, float percent, long long ignore_label) {
    unsigned valid = 0, best = 0;
    int arg = 0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        if (c < class_num) {
            const unsigned v = cnt[c];
            valid += v;
            if (v > best) { best = v; arg = c; }
        }
    }
    const float ratio = __fdiv_rn(__uint2float_rn(best), __fadd_rn(__uint2float_rn(valid), 1e-5f));
    if (ratio < percent) return kWinNone;
    if (static_cast<long long>(arg) == ignore_label) return kWinNone;  // where(out == ignore, labels, out) (:151)
    return static_cast<unsigned>(arg);
}

// ----------------------------------------------------------------------------------------
// cluster path
// ----------------------------------------------------------------------------------------
template <int CW>
__device__ __forceinline__ void flush_run(unsigned *bins, unsigned region, unsigned long long acc) {
    if (region == 0u || acc == 0ull) return;       // region 0 is never homogenised (:149)
    unsigned *base = bins + region * CW;
#pragma unroll
    for (int w = 0; w < CW; ++w) {
        const unsigned byte = static_cast<unsigned>(acc >> (8 * w)) & 0xFFu;
        if (byte) atomicAdd(base + w, (byte & 0xFu) | ((byte >> 4) << 16));
    }
}

template <int CW>
__global__ void __launch_bounds__(1024, 1)
lrh_cluster_kernel(const LrhArgs a, const int px_cta, const int nclusters) {
    extern __shared__ __align__(16) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned CL = cluster.num_blocks();
    const unsigned rank = cluster.block_rank();
    const int cluster_id = blockIdx.x / CL;
    const int tid = threadIdx.x, nthreads = blockDim.x;

    const int nwords = a.region_bound * CW;
    const int rpc = ((a.region_bound + static_cast<int>(CL) - 1) / static_cast<int>(CL) + 15) & ~15;  // regions owned per CTA
    unsigned *bins = reinterpret_cast<unsigned *>(smem);
    unsigned char *win = smem + ((static_cast<size_t>(nwords) * 4 + 15) & ~static_cast<size_t>(15));
    const size_t tile_off = ((static_cast<size_t>(nwords) * 4 + 15) & ~static_cast<size_t>(15)) + static_cast<size_t>(rpc) * CL;
    uint4 *reg16 = reinterpret_cast<uint4 *>(smem + tile_off);                         // 8 x u16 per group
    uint2 *lab8 = reinterpret_cast<uint2 *>(smem + tile_off + static_cast<size_t>(px_cta) * 2);  // 8 x u8 per group

    const int start = min(a.hw, static_cast<int>(rank) * px_cta);
    const int end = min(a.hw, start + px_cta);
    const int ngroups = (end - start) / kGroupPx;
    const int r_lo = min(a.region_bound, static_cast<int>(rank) * rpc);
    const int r_hi = min(a.region_bound, r_lo + rpc);
    bool bad_label = false, bad_region = false;

    for (int img = cluster_id; img < a.b; img += nclusters) {
        for (int i = tid; i < nwords; i += nthreads) bins[i] = 0u;
        __syncthreads();

        // ---- pass 1: HBM -> compressed smem tile + smem histogram -------------------------
        const long long *lab = a.labels + static_cast<size_t>(img) * a.hw + start;
        const long long *reg = a.regions + static_cast<size_t>(img) * a.hw + start;
        for (int g = tid; g < ngroups; g += nthreads) {
            const i64x4 l0 = ldg256_stream(lab + g * kGroupPx);
            const i64x4 l1 = ldg256_stream(lab + g * kGroupPx + 4);
            const i64x4 r0 = ldg256_stream(reg + g * kGroupPx);
            const i64x4 r1 = ldg256_stream(reg + g * kGroupPx + 4);
            unsigned rr[kGroupPx], cc[kGroupPx];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                rr[k] = checked_region(r0.v[k], a, bad_region);
                rr[k + 4] = checked_region(r1.v[k], a, bad_region);
                cc[k] = encode_label(l0.v[k], a, bad_label);
                cc[k + 4] = encode_label(l1.v[k], a, bad_label);
            }
            reg16[g] = make_uint4(rr[0] | (rr[1] << 16), rr[2] | (rr[3] << 16), rr[4] | (rr[5] << 16), rr[6] | (rr[7] << 16));
            lab8[g] = make_uint2(cc[0] | (cc[1] << 8) | (cc[2] << 16) | (cc[3] << 24), cc[4] | (cc[5] << 8) | (cc[6] << 16) | (cc[7] << 24));
            unsigned cur = rr[0];
            unsigned long long acc = 0ull;   // 16 nibbles: per-class count (<= 8) of the current run
#pragma unroll
            for (int k = 0; k < kGroupPx; ++k) {
                if (rr[k] != cur) {
                    flush_run<CW>(bins, cur, acc);
                    cur = rr[k];
                    acc = 0ull;
                }
                if (cc[k] < static_cast<unsigned>(kMaxFastClasses)) acc += 1ull << (4 * cc[k]);
            }
            flush_run<CW>(bins, cur, acc);
        }
        cluster.sync();

        // ---- merge: the owner CTA of a region sums the CL partial bins over DSMEM ----------
        // (winner bytes are written to the owner's OWN table; everyone pulls the slices after
        // the sync with 16-byte DSMEM loads -- remote byte stores were the slow part.)
        for (int r = r_lo + tid; r < r_hi; r += nthreads) {
            unsigned tot[2 * CW];
#pragma unroll
            for (int c = 0; c < 2 * CW; ++c) tot[c] = 0u;
            for (unsigned j = 0; j < CL; ++j) {
                const unsigned *rb = cluster.map_shared_rank(bins, j) + r * CW;
#pragma unroll
                for (int w = 0; w < CW; ++w) {
                    const unsigned v = rb[w];
                    tot[2 * w] += v & 0xFFFFu;
                    tot[2 * w + 1] += v >> 16;
                }
            }
            win[r] = static_cast<unsigned char>(lrh_winner<2 * CW>(tot, a.class_num, a.percent, a.ignore_label));
        }
        cluster.sync();
        {
            // rpc is a multiple of 16 (see plan), so every owner slice is uint4-aligned
            const int nvec = rpc / 16;
            for (int i = tid; i < nvec * static_cast<int>(CL); i += nthreads) {
                const unsigned j = static_cast<unsigned>(i / nvec);
                if (j == rank) continue;
                const int off = static_cast<int>(j) * rpc + (i - static_cast<int>(j) * nvec) * 16;
                if (off >= a.region_bound) continue;
                const uint4 v = *reinterpret_cast<const uint4 *>(cluster.map_shared_rank(win, j) + off);
                *reinterpret_cast<uint4 *>(win + off) = v;
            }
        }
        __syncthreads();

        // ---- pass 2: smem tile + winners -> HBM ------------------------------------------
        long long *out = a.out + static_cast<size_t>(img) * a.hw + start;
        for (int g = tid; g < ngroups; g += nthreads) {
            const uint4 rv = reg16[g];
            const uint2 lv = lab8[g];
            const unsigned rw[4] = {rv.x, rv.y, rv.z, rv.w};
            i64x4 o0, o1;
#pragma unroll
            for (int k = 0; k < kGroupPx; ++k) {
                const unsigned r = (rw[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
                const unsigned code = ((k < 4 ? lv.x : lv.y) >> ((k & 3) * 8)) & 0xFFu;
                const unsigned wv = win[r];
                const long long res = (r == 0u || wv == kWinNone) ? decode_label(code, a) : static_cast<long long>(wv);
                if (k < 4) o0.v[k] = res; else o1.v[k - 4] = res;
            }
            stg256_stream(out + g * kGroupPx, o0);
            stg256_stream(out + g * kGroupPx + 4, o1);
        }
        // Remote CTAs may still be pulling this CTA's winner slice: the owner rewrites it only
        // after the NEXT image's first cluster.sync, which every puller has passed by then.
        // bins are re-zeroed at the top of the loop, but remote readers of the bins all
        // finished before the second cluster.sync above.
    }
    if (bad_label) raise_flag(a.flags, REGDA_FLAG_LABEL_RANGE);
    if (bad_region) raise_flag(a.flags, REGDA_FLAG_REGION_RANGE);
}

// ----------------------------------------------------------------------------------------
// cluster path, fast variant: ignore_label == -1 (the value the training tools pass) and
// class_num <= 14.  Same structure as lrh_cluster_kernel, ~4x fewer instructions per pixel:
//   * labels are encoded as code = label + 1 (0 = ignored, 1..class_num = classes,
//     class_num + 1 = "label == class_num"), validated in bulk (OR of the high words, max of
//     the low words) instead of per element;
//   * the loads of the next group are issued before the current group is histogrammed;
//   * winners are stored as code (0 = keep the input label) with win[0] == 0, so region 0 and
//     "did not pass" need no test of their own in pass 2.
// ----------------------------------------------------------------------------------------
struct alignas(32) u32x8 { unsigned v[8]; };
__device__ __forceinline__ u32x8 ldg256_u32(const void *p) {
    u32x8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg256_u32(void *p, const u32x8 &r) {
    asm volatile("st.global.L1::no_allocate.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
                 :: "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]), "l"(p) : "memory");
}

// nibble accumulator -> the packed bin words of one region.  acc nibble 0 counts ignored pixels
// (dropped), nibble k+1 counts class k (<= 8 per group).  32-bit form: classes 0..5 (+ the
// "label == class_num" nibble 7, which is never a class and is masked out by the caller's CW).
__device__ __forceinline__ void red_shared_add(unsigned saddr, unsigned v) {
    asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(saddr), "r"(v) : "memory");
}
// `bins_s` is the 32-bit shared-window address of the bin table (no generic-address arithmetic per atomic)
template <int CW>
__device__ __forceinline__ void flush_acc(unsigned bins_s, unsigned region, unsigned acc) {
    static_assert(CW <= 3, "32-bit nibble accumulator holds 6 classes");
    const unsigned t = acc >> 4;
    const unsigned ev = t & 0x0F0F0F0Fu;           // classes 0,2,4 in bytes 0,1,2
    const unsigned od = (t >> 4) & 0x0F0F0F0Fu;    // classes 1,3,5 in bytes 0,1,2 ; byte 3 == 0
    const unsigned base = bins_s + region * (CW * 4);
#pragma unroll
    for (int w = 0; w < CW; ++w) {
        const unsigned v = __byte_perm(ev, od, (7u << 12) | ((4u + w) << 8) | (7u << 4) | static_cast<unsigned>(w));
        if (v) red_shared_add(base + 4 * w, v);
    }
}
template <int CW>
__device__ __forceinline__ void flush_acc(unsigned bins_s, unsigned region, unsigned long long acc) {
    const unsigned base = bins_s + region * (CW * 4);
#pragma unroll
    for (int w = 0; w < CW; ++w) {
        const unsigned lo = static_cast<unsigned>(acc >> (8 * w + 4)) & 0xFu;
        const unsigned hi = static_cast<unsigned>(acc >> (8 * w + 8)) & 0xFu;
        const unsigned v = lo | (hi << 16);
        if (v) red_shared_add(base + 4 * w, v);
    }
}

template <int CW, typename AccT, int HIST>
__global__ void __launch_bounds__(512, 2)
lrh_cluster_fast(const LrhArgs a, const int px_cta, const int nclusters, const int prefetch_iters) {
    extern __shared__ __align__(16) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned CL = cluster.num_blocks();
    const unsigned rank = cluster.block_rank();
    const int cluster_id = blockIdx.x / CL;
    const int tid = threadIdx.x, nthreads = blockDim.x;

    const int nwords = a.region_bound * CW;
    const int rpc = ((a.region_bound + static_cast<int>(CL) - 1) / static_cast<int>(CL) + 15) & ~15;
    unsigned *bins = reinterpret_cast<unsigned *>(smem);
    const size_t win_off = (static_cast<size_t>(nwords) * 4 + 15) & ~static_cast<size_t>(15);
    unsigned char *win = smem + win_off;                                       // winner code of EVERY region (rpc * CL bytes)
    unsigned *acc = reinterpret_cast<unsigned *>(smem + win_off + static_cast<size_t>(rpc) * CL);   // cluster-wide counts of OWN regions
    const size_t tile_off = win_off + static_cast<size_t>(rpc) * CL + static_cast<size_t>(rpc) * (2 * CW) * 4;
    uint4 *reg16 = reinterpret_cast<uint4 *>(smem + tile_off);
    uint2 *lab8 = reinterpret_cast<uint2 *>(smem + tile_off + static_cast<size_t>(px_cta) * 2);

    const int start = min(a.hw, static_cast<int>(rank) * px_cta);
    const int end = min(a.hw, start + px_cta);
    const int ngroups = (end - start) / kGroupPx;
    const int r_lo = min(a.region_bound, static_cast<int>(rank) * rpc);
    const int r_hi = min(a.region_bound, r_lo + rpc);
    for (int i = tid; i < rpc * 2 * CW; i += nthreads) acc[i] = 0u;          // re-zeroed by the owner after every use
    for (int i = tid; i < nwords; i += nthreads) bins[i] = 0u;                // re-zeroed by the push phase
    cluster.sync();                                                            // nobody pushes into an un-zeroed table
    const unsigned max_code = static_cast<unsigned>(a.class_num) + 1u;
    const unsigned bound = static_cast<unsigned>(a.region_bound);
    const unsigned bins_s = static_cast<unsigned>(__cvta_generic_to_shared(bins));
    bool bad_label = false, bad_region = false;

    for (int img = cluster_id; img < a.b; img += nclusters) {
        // (bins are zero here: zeroed before the loop and again by the push phase of the previous image, and every
        //  thread re-enters pass 1 with the groups it just finished in pass 2 -- no barrier between the two, so the
        //  stores of one image and the loads of the next overlap)

        // ---- pass 1 ---------------------------------------------------------------------
        const long long *lab = a.labels + static_cast<size_t>(img) * a.hw + start;
        const long long *reg = a.regions + static_cast<size_t>(img) * a.hw + start;
        for (int g = tid; g < ngroups; g += nthreads) {
            const u32x8 l0 = ldg256_u32(lab + g * kGroupPx), l1 = ldg256_u32(lab + g * kGroupPx + 4);
            const u32x8 r0 = ldg256_u32(reg + g * kGroupPx), r1 = ldg256_u32(reg + g * kGroupPx + 4);
            // code = label + 1 as a 64-bit add; valid iff the high word becomes 0 and low <= class_num + 1
            unsigned cc[kGroupPx], rr[kGroupPx];
            unsigned hi_or = 0u, lo_max = 0u, rhi_or = 0u, rlo_max = 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                cc[k] = l0.v[2 * k] + 1u;
                cc[k + 4] = l1.v[2 * k] + 1u;
                hi_or |= l0.v[2 * k + 1] + (cc[k] == 0u ? 1u : 0u);
                hi_or |= l1.v[2 * k + 1] + (cc[k + 4] == 0u ? 1u : 0u);
                lo_max = max(lo_max, max(cc[k], cc[k + 4]));
                rr[k] = r0.v[2 * k];
                rr[k + 4] = r1.v[2 * k];
                rhi_or |= r0.v[2 * k + 1] | r1.v[2 * k + 1];
                rlo_max = max(rlo_max, max(rr[k], rr[k + 4]));
            }
            if (hi_or != 0u || lo_max > max_code) {            // rare: sanitise element-wise
                bad_label = true;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (l0.v[2 * k + 1] + (cc[k] == 0u ? 1u : 0u) != 0u || cc[k] > max_code) cc[k] = 0u;
                    if (l1.v[2 * k + 1] + (cc[k + 4] == 0u ? 1u : 0u) != 0u || cc[k + 4] > max_code) cc[k + 4] = 0u;
                }
            }
            if (rhi_or != 0u || rlo_max >= bound) {
                bad_region = true;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (r0.v[2 * k + 1] != 0u || rr[k] >= bound) rr[k] = 0u;
                    if (r1.v[2 * k + 1] != 0u || rr[k + 4] >= bound) rr[k + 4] = 0u;
                }
            }
            reg16[g] = make_uint4(rr[0] | (rr[1] << 16), rr[2] | (rr[3] << 16), rr[4] | (rr[5] << 16), rr[6] | (rr[7] << 16));
            lab8[g] = make_uint2(cc[0] | (cc[1] << 8) | (cc[2] << 16) | (cc[3] << 24), cc[4] | (cc[5] << 8) | (cc[6] << 16) | (cc[7] << 24));
            if (HIST == 1) {
                // Histogram, variant 1: loop over the DISTINCT regions of the 8 pixels (1-3 in
                // practice; contiguity is irrelevant for counting).
                AccT oh[kGroupPx];
#pragma unroll
                for (int k = 0; k < kGroupPx; ++k) oh[k] = static_cast<AccT>(1) << (4 * cc[k]);
                unsigned rem = 0xFFu;
                unsigned r = rr[0];
                do {
                    unsigned m = 0u;
                    AccT acc = 0;
#pragma unroll
                    for (int k = 0; k < kGroupPx; ++k) {
                        const bool e = rr[k] == r;
                        m |= e ? (1u << k) : 0u;
                        acc += e ? oh[k] : static_cast<AccT>(0);
                    }
                    if (r != 0u) flush_acc<CW>(bins_s, r, acc);
                    rem &= ~m;
                    unsigned nr = rr[7];
#pragma unroll
                    for (int k = 6; k >= 1; --k) nr = (rem & (1u << k)) ? rr[k] : nr;
                    r = nr;
                } while (rem != 0u);
            } else {
                // Histogram, variant 0: counting does not care where in the group a region's pixels sit, so a group
                // with at most THREE distinct regions (first pixel's, last pixel's, one more) is accumulated without
                // branches on the pixel position; only >= 4 regions in 8 pixels go pixel by pixel.  (A warp executes
                // the union of its threads' paths, so the rare path must be rare per WARP, not per thread.)
                const unsigned ra = rr[0], rb = rr[7];
                unsigned in_a = 0u, in_b = 0u;
                AccT acc_a = 0, acc_b = 0, acc_all = 0;
#pragma unroll
                for (int k = 0; k < kGroupPx; ++k) {
                    const AccT one = static_cast<AccT>(1) << (4 * cc[k]);
                    const bool ea = rr[k] == ra;
                    const bool eb = !ea && rr[k] == rb;
                    in_a |= ea ? (1u << k) : 0u;
                    in_b |= eb ? (1u << k) : 0u;
                    acc_all += one;
                    acc_a += ea ? one : static_cast<AccT>(0);
                    acc_b += eb ? one : static_cast<AccT>(0);
                }
                const unsigned rest = ~(in_a | in_b) & 0xFFu;
                if (rest == 0u) {
                    if (ra != 0u) flush_acc<CW>(bins_s, ra, acc_a);
                    if (rb != 0u && in_b != 0u) flush_acc<CW>(bins_s, rb, acc_b);
                } else {
                    unsigned rm = rr[6];                       // first pixel that is in neither run (pixels 0 and 7 never are)
#pragma unroll
                    for (int k = 5; k >= 1; --k) rm = (rest & (1u << k)) ? rr[k] : rm;
                    unsigned in_m = 0u;
#pragma unroll
                    for (int k = 1; k < kGroupPx - 1; ++k) in_m |= (rr[k] == rm) ? (1u << k) : 0u;
                    if ((rest & ~in_m) == 0u) {
                        if (ra != 0u) flush_acc<CW>(bins_s, ra, acc_a);
                        if (rb != 0u && in_b != 0u) flush_acc<CW>(bins_s, rb, acc_b);
                        if (rm != 0u) flush_acc<CW>(bins_s, rm, static_cast<AccT>(acc_all - acc_a - acc_b));
                    } else {
#pragma unroll
                        for (int k = 0; k < kGroupPx; ++k)
                            if (rr[k] != 0u) flush_acc<CW>(bins_s, rr[k], static_cast<AccT>(static_cast<AccT>(1) << (4 * cc[k])));
                    }
                }
            }
        }
        __syncthreads();

        // ---- merge: PUSH.  Every CTA adds its non-zero packed bins to the owner CTA's 32-bit table with remote
        // shared-memory reductions (fire and forget: no DSMEM load round trips on the critical path), the owner
        // applies the exact float32 test locally and pushes 16-byte slices of winner codes to all CTAs. ----------
        for (int i = tid; i < nwords; i += nthreads) {
            const unsigned v = bins[i];
            if (v == 0u) continue;
            bins[i] = 0u;                                                    // ready for the next image
            const int r = i / CW, w = i - r * CW;
            const int owner = r / rpc;
            unsigned *dst = cluster.map_shared_rank(acc, owner) + (r - owner * rpc) * (2 * CW) + 2 * w;
            if (v & 0xFFFFu) atomicAdd(dst, v & 0xFFFFu);
            if (v >> 16) atomicAdd(dst + 1, v >> 16);
        }
        {   // keep HBM busy across the two cluster barriers: pull the head of the next image's slice into L2
            const int nimg = img + nclusters;
            if (nimg < a.b) {
                const char *nl = reinterpret_cast<const char *>(a.labels + static_cast<size_t>(nimg) * a.hw + start);
                const char *nr = reinterpret_cast<const char *>(a.regions + static_cast<size_t>(nimg) * a.hw + start);
                const size_t bytes = static_cast<size_t>(end - start) * 8;
                for (int it = 0; it < prefetch_iters; ++it) {
                    const size_t off = (static_cast<size_t>(it) * nthreads + tid) * 128;
                    if (off < bytes) {
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(nl + off));
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(nr + off));
                    }
                }
            }
        }
        cluster.sync();
        for (int r = r_lo + tid; r < r_hi; r += nthreads) {
            unsigned *ar = acc + (r - r_lo) * (2 * CW);
            unsigned tot[2 * CW];
#pragma unroll
            for (int c = 0; c < 2 * CW; ++c) { tot[c] = ar[c]; ar[c] = 0u; }
            const unsigned wv = lrh_winner<2 * CW>(tot, a.class_num, a.percent, a.ignore_label);
            win[r] = static_cast<unsigned char>((wv == kWinNone || r == 0) ? 0u : wv + 1u);
        }
        __syncthreads();
        {
            const int nvec = rpc / 16;                                         // own slice, in 16-byte vectors
            const uint4 *own = reinterpret_cast<const uint4 *>(win + static_cast<size_t>(rank) * rpc);
            for (int i = tid; i < nvec * static_cast<int>(CL); i += nthreads) {
                const unsigned j = static_cast<unsigned>(i / nvec);
                const int vi = i - static_cast<int>(j) * nvec;
                if (j == rank || static_cast<int>(rank) * rpc + vi * 16 >= a.region_bound) continue;
                reinterpret_cast<uint4 *>(cluster.map_shared_rank(win, j) + static_cast<size_t>(rank) * rpc)[vi] = own[vi];
            }
        }
        cluster.sync();

        // ---- pass 2 ---------------------------------------------------------------------
        long long *out = a.out + static_cast<size_t>(img) * a.hw + start;
        for (int g2 = tid; g2 < ngroups; g2 += nthreads) {
            const uint4 rv = reg16[g2];
            const uint2 lv = lab8[g2];
            const unsigned rw[4] = {rv.x, rv.y, rv.z, rv.w};
            u32x8 o0, o1;
#pragma unroll
            for (int k = 0; k < kGroupPx; ++k) {
                const unsigned r = (k & 1) ? (rw[k >> 1] >> 16) : (rw[k >> 1] & 0xFFFFu);
                const unsigned code = ((k < 4 ? lv.x : lv.y) >> ((k & 3) * 8)) & 0xFFu;
                const unsigned wv = win[r];
                const unsigned res = wv ? wv : code;          // code: 0 = ignore(-1), else class + 1
                const unsigned lo = res - 1u, hi = res ? 0u : 0xFFFFFFFFu;
                if (k < 4) { o0.v[2 * k] = lo; o0.v[2 * k + 1] = hi; } else { o1.v[2 * (k - 4)] = lo; o1.v[2 * (k - 4) + 1] = hi; }
            }
            stg256_u32(out + g2 * kGroupPx, o0);
            stg256_u32(out + g2 * kGroupPx + 4, o1);
        }
    }
    if (bad_label) raise_flag(a.flags, REGDA_FLAG_LABEL_RANGE);
    if (bad_region) raise_flag(a.flags, REGDA_FLAG_REGION_RANGE);
}

thread_local int g_path_mode = 0;   // 0 auto, 1 generic, 2 cluster forced, 3 cluster with the loop histogram
thread_local int g_last_path = 0;
thread_local int g_last_cluster = 0;

struct ClusterPlan {
    bool ok = false;
    int cluster = 0;
    int px_cta = 0;
    int cw = 0;
    int threads = 0;
    int prefetch = 0;
    size_t smem = 0;
};

int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

size_t cluster_smem_bytes(int region_bound, int cw, int px_cta, int cl) {
    const size_t bins = (static_cast<size_t>(region_bound) * cw * 4 + 15) & ~static_cast<size_t>(15);
    const size_t rpc = static_cast<size_t>(((region_bound + cl - 1) / cl + 15) & ~15);
    // packed local bins + winner bytes of all regions + 32-bit cluster-wide counts of the owned regions + 3 B/px tile
    return bins + rpc * cl + rpc * (2 * cw) * 4 + static_cast<size_t>(px_cta) * 3;
}

ClusterPlan plan_cluster(int b, int64_t hw, int class_num, int64_t region_bound, bool allow_single) {
    ClusterPlan p;
    if (class_num < 1 || class_num > kMaxFastClasses) return p;
    if (region_bound < 1 || region_bound > 65536) return p;   // u16 region ids in the smem tile
    if (hw < kGroupPx || hw % kGroupPx != 0 || hw > (1ll << 30)) return p;
    p.cw = (class_num + 1) / 2;
    const size_t limit = static_cast<size_t>(max_optin_smem());
    // Candidates (cluster size, CTAs per SM): several CTAs per SM in different phases (load / merge barriers / store)
    // are what keeps HBM busy.  One CTA per SM is only taken when the path is forced: the 3-kernel global-bin path
    // is faster there (measured: 5000 regions/tile, 27 % vs 40 % of the HBM roofline).
    // Tuning knobs (profiling only): REGDA_LRH_CLUSTER, REGDA_LRH_PER_SM, REGDA_LRH_THREADS, REGDA_LRH_PREFETCH.
    // small bin tables: 16-CTA clusters measured faster (50 regions/tile: 72.8 % vs 69.8 % of the HBM roofline)
    const bool small = b * 8 <= sm_count() / 2 || region_bound <= 256;
    int order[6][2] = {{small ? 16 : 8, 2}, {small ? 8 : 16, 2}, {16, 3}, {small ? 16 : 8, 1}, {small ? 8 : 16, 1}, {0, 0}};
    int ncand = allow_single ? 5 : 3;
    const int fc = env_int("REGDA_LRH_CLUSTER", 0), fp = env_int("REGDA_LRH_PER_SM", 0);
    if (fc > 0 && fp > 0) { order[0][0] = fc; order[0][1] = fp; ncand = 1; }
    for (int i = 0; i < ncand; ++i) {
        const int cl = order[i][0], per_sm = order[i][1];
        int px = static_cast<int>((hw + cl - 1) / cl);
        px = (px + kGroupPx - 1) / kGroupPx * kGroupPx;
        if (px > 65535) continue;                              // u16 per-CTA counters
        const size_t s = cluster_smem_bytes(static_cast<int>(region_bound), p.cw, px, cl);
        if (s > (limit + 1024) / per_sm - 1024) continue;      // per_sm CTAs per SM (1 KB reserved each)
        const int groups = px / kGroupPx;
        p.ok = true; p.cluster = cl; p.px_cta = px; p.smem = s;
        // 64 registers/thread: 1024 threads per SM fill the register file
        p.threads = per_sm == 3 ? 256 : (groups >= 512 ? 512 : 256);
        p.threads = env_int("REGDA_LRH_THREADS", p.threads);
        p.prefetch = env_int("REGDA_LRH_PREFETCH", 0);
        return p;
    }
    return p;
}

template <typename Kern, typename... Extra>
int launch_cluster_kernel(Kern kern, const LrhArgs &a, const ClusterPlan &p, cudaStream_t st, bool *launched, Extra... extra) {
    *launched = false;
    REGDA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem)));
    if (p.cluster > 8) REGDA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(p.threads);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3(p.cluster * a.b);
    int max_clusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
    if (e != cudaSuccess || max_clusters < 1) {
        (void)cudaGetLastError();
        return REGDA_OK;                                       // not launchable here: caller falls back
    }
    const int nclusters = a.b < max_clusters ? a.b : max_clusters;
    cfg.gridDim = dim3(p.cluster * nclusters);
    REGDA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, a, p.px_cta, nclusters, extra...));
    *launched = true;
    return REGDA_OK;
}

template <int CW>
int launch_cluster(const LrhArgs &a, const ClusterPlan &p, cudaStream_t st, bool *launched) {
    if (a.ignore_label == -1 && p.threads <= 512) {
        if constexpr (CW <= 3) {
            if (a.class_num <= 6) {
                if (g_path_mode == 3) return launch_cluster_kernel(lrh_cluster_fast<CW, unsigned, 1>, a, p, st, launched, p.prefetch);
                return launch_cluster_kernel(lrh_cluster_fast<CW, unsigned, 0>, a, p, st, launched, p.prefetch);
            }
        }
        if (a.class_num <= 14) return launch_cluster_kernel(lrh_cluster_fast<CW, unsigned long long, 1>, a, p, st, launched, p.prefetch);
    }
    return launch_cluster_kernel(lrh_cluster_kernel<CW>, a, p, st, launched);
}

// ----------------------------------------------------------------------------------------
// generic path
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lrh_hist_global(const LrhArgs a, unsigned *__restrict__ gcnt) {
    const int img = blockIdx.y;
    const long long *lab = a.labels + static_cast<size_t>(img) * a.hw;
    const long long *reg = a.regions + static_cast<size_t>(img) * a.hw;
    unsigned *cnt = gcnt + static_cast<size_t>(img) * a.region_bound * a.class_num;
    bool bad_label = false, bad_region = false;
    const int stride = gridDim.x * blockDim.x;
    const int iters = (a.hw + stride - 1) / stride;
    for (int it = 0; it < iters; ++it) {
        const int i = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
        bool counted = false;
        unsigned key = 0;
        if (i < a.hw) {
            const unsigned code = encode_label(ldg64_stream(lab + i), a, bad_label);
            const unsigned r = checked_region(ldg64_stream(reg + i), a, bad_region);
            if (code < static_cast<unsigned>(a.class_num) && r != 0u) {
                counted = true;
                key = r * static_cast<unsigned>(a.class_num) + code;
            }
        }
        const unsigned voters = __ballot_sync(0xffffffffu, counted);
        if (counted) {
            const unsigned peers = __match_any_sync(voters, key);
            if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(cnt + key, __popc(peers));
        }
    }
    if (bad_label) raise_flag(a.flags, REGDA_FLAG_LABEL_RANGE);
    if (bad_region) raise_flag(a.flags, REGDA_FLAG_REGION_RANGE);
}

__global__ void __launch_bounds__(256)
lrh_winner_global(const LrhArgs a, const unsigned *__restrict__ gcnt, unsigned char *__restrict__ gwin) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(a.b) * a.region_bound) return;
    const unsigned *cnt = gcnt + i * a.class_num;
    unsigned valid = 0, best = 0;
    int arg = 0;
    for (int c = 0; c < a.class_num; ++c) {
        const unsigned v = cnt[c];
        valid += v;
        if (v > best) { best = v; arg = c; }
    }
    const float ratio = __fdiv_rn(__uint2float_rn(best), __fadd_rn(__uint2float_rn(valid), 1e-5f));
    // class ids above 254 cannot be a byte: the generic path stores winner+1 in 16 bits instead
    unsigned short *w16 = reinterpret_cast<unsigned short *>(gwin);
    w16[i] = (ratio < a.percent || static_cast<long long>(arg) == a.ignore_label) ? 0 : static_cast<unsigned short>(arg + 1);
}

__global__ void __launch_bounds__(256)
lrh_apply_global(const LrhArgs a, const unsigned char *__restrict__ gwin) {
    const int img = blockIdx.y;
    const long long *lab = a.labels + static_cast<size_t>(img) * a.hw;
    const long long *reg = a.regions + static_cast<size_t>(img) * a.hw;
    long long *out = a.out + static_cast<size_t>(img) * a.hw;
    const unsigned short *w16 = reinterpret_cast<const unsigned short *>(gwin) + static_cast<size_t>(img) * a.region_bound;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.hw; i += gridDim.x * blockDim.x) {
        const long long l = ldg64_stream(lab + i);
        const long long r = ldg64_stream(reg + i);
        long long res = l;
        if (r > 0 && r < a.region_bound) {
            const unsigned w = w16[r];
            if (w != 0u) res = static_cast<long long>(w) - 1;
        }
        out[i] = res;
    }
}

__global__ void __launch_bounds__(256)
region_bound_kernel(const long long *__restrict__ regions, long long n, long long *bound, int32_t *flags) {
    long long m = -1;
    bool neg = false;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = regions[i];
        neg |= r < 0;
        m = r > m ? r : m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long other = __shfl_xor_sync(0xffffffffu, m, o);
        m = other > m ? other : m;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(bound, m + 1);
    if (neg) raise_flag(flags, REGDA_FLAG_REGION_RANGE);
}


size_t generic_workspace(int b, int class_num, int64_t region_bound) {
    const size_t cnt = align_up(static_cast<size_t>(b) * region_bound * class_num * 4, 256);
    const size_t win = align_up(static_cast<size_t>(b) * region_bound * 2, 256);
    return cnt + win;
}

}  // namespace
}  // namespace regda

using namespace regda;

extern "C" int regda_set_lrh_path(int mode) {
    if (mode < 0 || mode > 3) return fail(REGDA_ERR_INVALID_ARG, "lrh path mode must be 0..3");
    g_path_mode = mode;
    return REGDA_OK;
}

extern "C" int regda_lrh_last_path(int *cluster) {
    if (cluster) *cluster = g_last_cluster;
    return g_last_path;
}

extern "C" size_t regda_lrh_workspace_bytes(int b, int64_t hw, int class_num, int64_t region_bound) {
    (void)hw;
    if (b < 0 || class_num < 1 || region_bound < 1) return 0;
    return generic_workspace(b, class_num, region_bound);
}

extern "C" int regda_lrh_forward(const int64_t *labels, const int64_t *regions, int64_t *out,
                                 int b, int64_t hw, int class_num, int64_t ignore_label, double percent,
                                 int64_t region_bound, int32_t *flags,
                                 void *workspace, size_t workspace_bytes, void *stream) {
    if (b < 0 || hw < 0) return fail(REGDA_ERR_INVALID_ARG, "lrh: negative shape");
    if (b == 0 || hw == 0) return REGDA_OK;
    if (!labels || !regions || !out) return fail(REGDA_ERR_INVALID_ARG, "lrh: null tensor pointer");
    if (class_num < 1 || class_num > 65534) return fail(REGDA_ERR_INVALID_ARG, "lrh: class_num out of range");
    if (region_bound < 1) return fail(REGDA_ERR_INVALID_ARG, "lrh: region_bound must be >= 1");
    if (hw > 0x7fffffffll || region_bound * class_num > 0x7fffffffll)
        return fail(REGDA_ERR_UNSUPPORTED, "lrh: image or bin table too large for 32-bit indexing");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LrhArgs a{reinterpret_cast<const long long *>(labels), reinterpret_cast<const long long *>(regions),
              reinterpret_cast<long long *>(out), b, static_cast<int>(hw), class_num,
              static_cast<long long>(ignore_label), static_cast<float>(percent), static_cast<int>(region_bound), flags};

    const bool aligned = (reinterpret_cast<uintptr_t>(labels) % 32 == 0) && (reinterpret_cast<uintptr_t>(regions) % 32 == 0) &&
                         (reinterpret_cast<uintptr_t>(out) % 32 == 0) && (hw % 4 == 0);
    if (g_path_mode != 1 && aligned) {
        const ClusterPlan p = plan_cluster(b, hw, class_num, region_bound, g_path_mode >= 2);
        if (p.ok) {
            bool launched = false;
            int rc = REGDA_OK;
            switch (p.cw) {
                case 1: rc = launch_cluster<1>(a, p, st, &launched); break;
                case 2: rc = launch_cluster<2>(a, p, st, &launched); break;
                case 3: rc = launch_cluster<3>(a, p, st, &launched); break;
                case 4: rc = launch_cluster<4>(a, p, st, &launched); break;
                case 5: rc = launch_cluster<5>(a, p, st, &launched); break;
                case 6: rc = launch_cluster<6>(a, p, st, &launched); break;
                case 7: rc = launch_cluster<7>(a, p, st, &launched); break;
                default: rc = launch_cluster<8>(a, p, st, &launched); break;
            }
            if (rc != REGDA_OK) return rc;
            if (launched) {
                g_last_path = 2;
                g_last_cluster = p.cluster;
                return REGDA_OK;
            }
        }
    }
    if (g_path_mode >= 2) return fail(REGDA_ERR_UNSUPPORTED, "lrh: cluster path forced but shape/alignment does not fit it");

    const size_t need = generic_workspace(b, class_num, region_bound);
    if (!workspace || workspace_bytes < need) return fail(REGDA_ERR_WORKSPACE, "lrh: workspace too small (%zu < %zu)", workspace_bytes, need);
    unsigned *gcnt = static_cast<unsigned *>(workspace);
    unsigned char *gwin = static_cast<unsigned char *>(workspace) + align_up(static_cast<size_t>(b) * region_bound * class_num * 4, 256);
    REGDA_CUDA_CHECK(cudaMemsetAsync(gcnt, 0, static_cast<size_t>(b) * region_bound * class_num * 4, st));
    const int per_img = static_cast<int>(std::min<int64_t>((hw + 1023) / 1024, std::max(1, 8 * sm_count() / b)));
    lrh_hist_global<<<dim3(per_img, b), 256, 0, st>>>(a, gcnt);
    REGDA_LAUNCH_CHECK();
    const long long nreg = static_cast<long long>(b) * region_bound;
    lrh_winner_global<<<static_cast<unsigned>((nreg + 255) / 256), 256, 0, st>>>(a, gcnt, gwin);
    REGDA_LAUNCH_CHECK();
    lrh_apply_global<<<dim3(per_img, b), 256, 0, st>>>(a, gwin);
    REGDA_LAUNCH_CHECK();
    g_last_path = 1;
    g_last_cluster = 0;
    return REGDA_OK;
}

extern "C" int regda_region_bound(const int64_t *regions, int64_t n, int64_t *bound_out, int32_t *flags, void *stream) {
    if (!bound_out) return fail(REGDA_ERR_INVALID_ARG, "region_bound: null output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    REGDA_CUDA_CHECK(cudaMemsetAsync(bound_out, 0, sizeof(int64_t), st));
    if (n <= 0) return REGDA_OK;
    if (!regions) return fail(REGDA_ERR_INVALID_ARG, "region_bound: null input");
    const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 8ll * sm_count()));
    region_bound_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const long long *>(regions), n, reinterpret_cast<long long *>(bound_out), flags);
    REGDA_LAUNCH_CHECK();
    return REGDA_OK;
}
