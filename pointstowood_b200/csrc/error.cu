// error.cu -- last-error slot and device query for the C ABI.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace p2w {
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return P2W_ECUDA;
    }
    return P2W_OK;
}
}  // namespace p2w

extern "C" int p2w_version(void) { return 100; }
extern "C" const char *p2w_last_error(void) { return p2w::g_err; }

extern "C" int p2w_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) {
        p2w::set_error("p2w_device_info: %s", cudaGetErrorString(e));
        return P2W_ENOGPU;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return P2W_OK;
}

extern "C" long long p2w_launch_count(void) { return p2w::launches(); }
