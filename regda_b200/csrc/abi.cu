// ABI bookkeeping: version, per-thread error text, cached device properties.
#include <cstdarg>

#include "common.cuh"

namespace regda {
namespace {
thread_local char g_err[512] = "";
int g_sm_count = 0;
int g_optin_smem = 0;

void query_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return; }
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) g_sm_count = v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess) g_optin_smem = v;
    (void)cudaGetLastError();
}
}  // namespace

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int sm_count() {
    if (g_sm_count == 0) query_device();
    return g_sm_count > 0 ? g_sm_count : kSmCountFallback;
}

int max_optin_smem() {
    if (g_optin_smem == 0) query_device();
    return g_optin_smem > 0 ? g_optin_smem : 232448;  // 227 KB on sm_100
}
}  // namespace regda

extern "C" int regda_abi_version(void) { return REGDA_ABI_VERSION; }
extern "C" const char *regda_last_error(void) { return regda::g_err; }
