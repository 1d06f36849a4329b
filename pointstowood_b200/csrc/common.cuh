// common.cuh -- shared helpers for libp2w.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/p2w.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libp2w targets sm_100a only"
#endif

namespace p2w {

constexpr int kNumSMs = 148;   // B200; grids for persistent kernels are sized from the device query

void set_error(const char *fmt, ...);
int check_launch(const char *what);
void count_launches(int n);   // kernels enqueued by libp2w since load (bench.py's gpu_launches)
long long launches();

#define P2W_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            p2w::set_error(__VA_ARGS__);       \
            return P2W_EINVAL;                 \
        }                                      \
    } while (0)

// every kernel launch goes through this macro so that p2w_launch_count() is exact
#define P2W_LAUNCH(kernel, ...) p2w::count_launches(1), kernel<<<__VA_ARGS__>>>

static inline cudaStream_t as_stream(p2w_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// first b with ptr[b+1] > i  (ptr ascending, ptr[0] = 0, i < ptr[B])
__device__ __forceinline__ int find_tile(const int64_t *__restrict__ ptr, int B, int64_t i) {
    int lo = 0, hi = B - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (ptr[mid + 1] > i) hi = mid; else lo = mid + 1;
    }
    return lo;
}
#endif

}  // namespace p2w
