// Phase trace of the BatchNorm apply kernel (debug harness, not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DREGDA_BN_TRACE -Iinclude -Iregda_b200/csrc \
//        scripts/bn_trace.cu regda_b200/csrc/abi.cu -o scripts/bn_trace
// Block (0..) stamps %globaltimer at entry / constants ready / end of its first and second batch / exit.
#include "../regda_b200/csrc/norm.cu"

#include <algorithm>
#include <vector>

int main() {
    struct Sh { const char *name; long long npix; int c; } shapes[] = {{"layer3 inner (16x32x32 x 256)", 16384, 256}, {"layer3 outer (16x32x32 x 1024)", 16384, 1024},
                                                                    {"layer2 inner (16x64x64 x 128)", 65536, 128}};
    for (const Sh &sh : shapes) {
        const size_t n = static_cast<size_t>(sh.npix) * sh.c;
        __nv_bfloat16 *y, *out;
        float *stats, *gamma, *beta, *rm, *rv;
        long long *nbt;
        unsigned char *mask;
        cudaMalloc(&y, n * 2); cudaMalloc(&out, n * 2); cudaMalloc(&mask, n / 8);
        cudaMalloc(&stats, 4 * sh.c * 4); cudaMalloc(&gamma, sh.c * 4); cudaMalloc(&beta, sh.c * 4); cudaMalloc(&rm, sh.c * 4); cudaMalloc(&rv, sh.c * 4);
        cudaMalloc(&nbt, 8);
        cudaMemset(y, 0, n * 2); cudaMemset(stats, 0, 4 * sh.c * 4); cudaMemset(gamma, 0, sh.c * 4); cudaMemset(beta, 0, sh.c * 4);
        cudaMemset(rm, 0, sh.c * 4); cudaMemset(rv, 0, sh.c * 4); cudaMemset(nbt, 0, 8);
        cudaStream_t st;
        cudaStreamCreate(&st);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int reps = 30;
        for (int i = 0; i < 5; ++i)
            regda_bn_forward_bf16(y, nullptr, out, sh.npix, sh.c, 2, gamma, beta, rm, rv, reinterpret_cast<int64_t *>(nbt), 1e-5, 0.1, 1, stats, 1, 1, mask, st);
        cudaStreamSynchronize(st);
        cudaEventRecord(e0, st);
        for (int i = 0; i < reps; ++i)
            if (regda_bn_forward_bf16(y, nullptr, out, sh.npix, sh.c, 2, gamma, beta, rm, rv, reinterpret_cast<int64_t *>(nbt), 1e-5, 0.1, 1, stats, 1, 1, mask, st)) {
                printf("launch failed: %s\n", regda_last_error());
                return 1;
            }
        cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<unsigned long long> tr(kBnTraceBlocks * kBnTraceSlots);
        cudaMemcpyFromSymbol(tr.data(), g_bn_trace, tr.size() * 8);
        const int blocks = apply_grid(sh.npix / 2 * sh.c, 2) * 2;
        unsigned long long first = ~0ull, last = 0;
        for (int b = 0; b < blocks; ++b) { first = std::min(first, tr[b * kBnTraceSlots]); last = std::max(last, tr[b * kBnTraceSlots + 4]); }
        printf("== %s: %zu MB in, %d blocks: %.2f us per launch back to back (events); last launch: first entry -> last exit %.2f us\n", sh.name,
               n * 2 >> 20, blocks, ms * 1e3 / reps, (last - first) / 1e3);
        const char *nm[5] = {"entry", "constants ready", "batch 1 stored", "batch 2 stored", "exit"};
        for (int s = 0; s < 5; ++s) {
            std::vector<double> v;
            for (int b = 0; b < blocks; ++b) if (tr[b * kBnTraceSlots + s] >= first) v.push_back((tr[b * kBnTraceSlots + s] - first) / 1e3);
            std::sort(v.begin(), v.end());
            if (!v.empty()) printf("   %-18s min %6.2f  median %6.2f  max %6.2f us\n", nm[s], v.front(), v[v.size() / 2], v.back());
        }
        cudaFree(y); cudaFree(out); cudaFree(mask); cudaFree(stats);
    }
    return 0;
}
