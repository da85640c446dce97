// Shared helpers for the regda_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

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
