// Shared helpers for the regda_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "regda_b200.h"

namespace regda {

constexpr int kSmCountFallback = 148;

// thread-local error text behind regda_last_error()
void set_error(const char *fmt, ...);
int fail(int code, const char *fmt, ...);
int sm_count();
int max_optin_smem();

#define REGDA_CUDA_CHECK(expr)                                                              \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return ::regda::fail(REGDA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,            \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);               \
    } while (0)

#define REGDA_LAUNCH_CHECK()                                                                \
    do {                                                                                    \
        cudaError_t _e = cudaPeekAtLastError();                                             \
        if (_e != cudaSuccess)                                                              \
            return ::regda::fail(REGDA_ERR_CUDA, "kernel launch failed: %s (%s:%d)",        \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);               \
    } while (0)

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// A kernel that calls pdl_trigger() at its start lets the NEXT kernel of the stream be scheduled while this one is still
// running; that kernel runs its prologue (barrier / tensor-memory / descriptor set-up) and then blocks in pdl_wait() until
// this grid has completed and its memory is visible.  Rule for every PDL-launched kernel: NO global-memory access before
// pdl_wait().  REGDA_PDL=0 launches them as plain stream-ordered kernels (both intrinsics are then no-ops).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// PDL level: 0 = off, 1 = kernels of level 1 (the tensor-core kernels, which have a real prologue to overlap), 2 = all.
// Default 1 (measured on the step, round 2: 0 -> 1010.8, 1 -> 1027.2, 2 -> 999.2 images/s with the trigger at the kernel start);
// REGDA_PDL overrides for sweeps.  (Moving the tensor-core kernels' trigger from their start to "last MMA issued" measured
// neutral at level 1 -- 1018 vs 1017 images/s -- and did not rescue level 2 -- 999; profiles/pdl_sweep_round2.txt.)
inline int pdl_level() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("REGDA_PDL");
        v = e ? atoi(e) : 1;
    }
    return v;
}
template <int LEVEL = 1, typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_level() >= LEVEL ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- 256-bit streaming global accesses (LDG.E.256 / STG.E.256 on sm_100) --------------
struct alignas(32) i64x4 { long long v[4]; };

__device__ __forceinline__ i64x4 ldg256_stream(const void *p) {
    i64x4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.v[0]), "=l"(r.v[1]), "=l"(r.v[2]), "=l"(r.v[3]) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg256_stream(void *p, const i64x4 &r) {
    asm volatile("st.global.L1::no_allocate.v4.s64 [%4], {%0,%1,%2,%3};"
                 :: "l"(r.v[0]), "l"(r.v[1]), "l"(r.v[2]), "l"(r.v[3]), "l"(p) : "memory");
}
__device__ __forceinline__ long long ldg64_stream(const long long *p) {
    long long r;
    asm volatile("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}

__device__ __forceinline__ void raise_flag(int32_t *flags, int bit) {
    if (flags != nullptr) atomicOr(flags, bit);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace regda
