// Phase trace of the persistent convolution kernel (debug harness, not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DREGDA_CONV_TRACE -Iinclude -Iregda_b200/csrc \
//        scripts/conv_trace.cu regda_b200/csrc/abi.cu -o scripts/conv_trace -lcuda
// Each CTA of conv_persistent_kernel stamps %globaltimer / clock64 at its phase boundaries (CONV_TRACE slots in conv_tc.cu);
// the harness launches the same convolution back to back and prints, per launch, where the time between "first CTA entered"
// and "next launch's first CTA entered" goes.
#include "../regda_b200/csrc/conv_tc.cu"

#include <algorithm>
#include <vector>

namespace {

struct Shape { const char *name; int n, h, w, cin, cout, r, pad; };

const char *kSlotName[14] = {"entry", "prologue done", "pdl_wait passed", "mma: 1st stage landed", "mma: last commit", "epi: 1st tile ready",
                             "epi: last tile ready", "epi: tiles drained", "epi: stats flushed", "epi: stores read", "cta sync", "tmem freed",
                             "tma: 1st issue", "tma: all issued"};

template <bool STATS>
int run(const Shape &sh, int launches, int grid_override) {
    ConvGeom g;
    geom_init(g, sh.n, sh.h, sh.w, sh.cin, sh.cout, sh.r, sh.r, 1, sh.pad, 1);
    const size_t xe = static_cast<size_t>(sh.n) * sh.h * sh.w * sh.cin, we = static_cast<size_t>(sh.cout) * sh.r * sh.r * sh.cin;
    const size_t ye = static_cast<size_t>(sh.n) * g.oh * g.ow * sh.cout;
    __nv_bfloat16 *x, *w, *y;
    float *stats;
    cudaMalloc(&x, xe * 2); cudaMalloc(&w, we * 2); cudaMalloc(&y, ye * 2); cudaMalloc(&stats, 4 * sh.cout * sizeof(float));
    cudaMemset(x, 0, xe * 2); cudaMemset(w, 0, we * 2); cudaMemset(stats, 0, 4 * sh.cout * sizeof(float));
    CUtensorMap tx, tw;
    if (make_tmap_x(&tx, x, g)) return 1;
    const int block_n = pick_block_n(g);
    if (make_tmap_w(&tw, w, g.cout, sh.r * sh.r * g.wct, block_n)) return 1;
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const bool early = getenv("REGDA_TRACE_EARLY") && atoi(getenv("REGDA_TRACE_EARLY")) != 0;
    auto launch = [&](int id) {
        const int ipg = (getenv("REGDA_TRACE_ONE_GROUP") ? sh.n : sh.n / 2) | (id << 16);
        if (early) regda_conv_hint_static_weights();
        if (block_n == 256) return launch_persistent_impl<256, 4, false, STATS, false>(tx, tw, y, g, st, stats, ipg, nullptr);
        if (block_n == 128) return launch_persistent_impl<128, 6, false, STATS, false>(tx, tw, y, g, st, stats, ipg, nullptr);
        return launch_persistent_impl<64, 8, false, STATS, false>(tx, tw, y, g, st, stats, ipg, nullptr);
    };
    (void)grid_override;
    for (int i = 0; i < 5; ++i) launch(0);
    cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    for (int i = 0; i < launches; ++i)
        if (launch(i)) { printf("launch failed: %s\n", regda_last_error()); return 1; }
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<unsigned long long> tr(static_cast<size_t>(kTraceLaunches) * kTraceCtas * kTraceSlots * 2);
    cudaMemcpyFromSymbol(tr.data(), g_conv_trace, tr.size() * 8);
    const int n_tiles = (g.cout / block_n) * g.tiles_img * g.tiles_h * g.tiles_w;
    const int grid = std::min(n_tiles, sm_count());
    printf("== %s%s: M=%d N=%d K=%d  block_n=%d tiles=%d grid=%d  %.2f us per launch (events over %d launches, PDL level %d)\n", sh.name,
           STATS ? " +stats" : "", sh.n * g.oh * g.ow, sh.cout, sh.cin * sh.r * sh.r, block_n, n_tiles, grid, ms * 1e3 / launches, launches, pdl_level());
    auto at = [&](int l, int c, int s, int k) { return tr[((static_cast<size_t>(l) * kTraceCtas + c) * kTraceSlots + s) * 2 + k]; };
    // per launch: first entry, last exit; per slot the median / max over CTAs relative to the launch's first entry (globaltimer ns)
    std::vector<double> med(14, 0), mx(14, 0), clk_epi(0);
    double gap = 0, span = 0, period = 0;
    int cnt = 0;
    for (int l = launches / 2; l < launches - 1; ++l) {
        unsigned long long first = ~0ull, last = 0, next_first = ~0ull;
        for (int c = 0; c < grid; ++c) {
            first = std::min(first, at(l, c, 0, 0));
            last = std::max(last, at(l, c, 11, 0));
            next_first = std::min(next_first, at(l + 1, c, 0, 0));
        }
        for (int s = 0; s < 14; ++s) {
            std::vector<double> v;
            for (int c = 0; c < grid; ++c) v.push_back(static_cast<double>(static_cast<long long>(at(l, c, s, 0) - first)));
            std::sort(v.begin(), v.end());
            med[s] += v[v.size() / 2];
            mx[s] += v.back();
        }
        gap += static_cast<double>(static_cast<long long>(next_first - last));
        span += static_cast<double>(last - first);
        period += static_cast<double>(next_first - first);
        ++cnt;
    }
    printf("   launch period %.2f us = span (first entry -> last tmem free) %.2f us + gap to the next launch's first entry %.2f us\n",
           period / cnt / 1e3, span / cnt / 1e3, gap / cnt / 1e3);
    const int order[14] = {0, 1, 2, 12, 3, 13, 5, 4, 6, 7, 8, 9, 10, 11};
    for (int i = 0; i < 14; ++i) {
        const int s = order[i];
        printf("   %-24s median %8.2f us   max %8.2f us\n", kSlotName[s], med[s] / cnt / 1e3, mx[s] / cnt / 1e3);
    }
    // SM-clock deltas inside a CTA (cycles): median over CTAs of the last traced launch
    const int l = launches - 2;
    auto cyc = [&](int a, int b) {
        std::vector<long long> v;
        for (int c = 0; c < grid; ++c) v.push_back(static_cast<long long>(at(l, c, b, 1) - at(l, c, a, 1)));
        std::sort(v.begin(), v.end());
        return v[v.size() / 2];
    };
    printf("   cycles (median CTA): entry->prologue %lld, prologue->pdl_wait %lld, pdl_wait->1st stage %lld, 1st stage->last commit %lld, "
           "last commit->last tile ready %lld, last tile ready->drained %lld, drained->flushed %lld, flushed->stores read %lld, ->sync %lld, ->tmem freed %lld\n",
           cyc(0, 1), cyc(1, 2), cyc(2, 3), cyc(3, 4), cyc(4, 6), cyc(6, 7), cyc(7, 8), cyc(8, 9), cyc(9, 10), cyc(10, 11));
    cudaFree(x); cudaFree(w); cudaFree(y); cudaFree(stats);
    return 0;
}

}  // namespace

int main() {
    const Shape shapes[] = {
        {"l3.conv1", 16, 32, 32, 1024, 256, 1, 0}, {"l3.conv2", 16, 32, 32, 256, 256, 3, 1}, {"l3.conv3", 16, 32, 32, 256, 1024, 1, 0},
        {"l2.conv1", 16, 64, 64, 512, 128, 1, 0},  {"l2.conv3", 16, 64, 64, 128, 512, 1, 0}, {"l1.conv3", 16, 128, 128, 64, 256, 1, 0},
        {"l4.conv1", 16, 32, 32, 2048, 512, 1, 0},
    };
    for (const Shape &sh : shapes) {
        if (run<true>(sh, 40, 0)) return 1;
        if (run<false>(sh, 40, 0)) return 1;
    }
    return 0;
}
