// Experiment (not part of libp2w): issue rate of tcgen05.mma.cta_group::1.kind::f16 M=128 N=128 K=16 with
// both operands in shared memory, as a function of the smem layout type of the descriptors.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu ; run on a B200.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// mode: 0 none K-major (LBO 2048, SBO 128), 1 SW128 K-major, 2 SW64 K-major, 3 SW32 K-major,
//       4 none with B MN-major (as libp2w's hid tile), 5 none K-major, padded LBO (2064, as the msg tile)
__global__ void __launch_bounds__(128) rate_kernel(int mode, int n_mma, int nacc, int N, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    unsigned char *A = smem, *B = smem + 65536;
    for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((mode == 4 ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
        uint64_t ad, bd;
        uint32_t kstep_a, kstep_b;          // descriptor increment per K=16 step (in 16-byte units)
        const uint32_t a = smem_u32(A), b = smem_u32(B);
        if (mode == 1) { ad = make_desc(a, 16, 1024, 2); bd = make_desc(b, 16, 1024, 2); kstep_a = kstep_b = 2; }
        else if (mode == 2) { ad = make_desc(a, 16, 512, 4); bd = make_desc(b, 16, 512, 4); kstep_a = kstep_b = 2; }
        else if (mode == 3) { ad = make_desc(a, 16, 256, 6); bd = make_desc(b, 16, 256, 6); kstep_a = kstep_b = 2; }
        else if (mode == 4) { ad = make_desc(a, 2048, 128, 0); bd = make_desc(b, 128, 192 * 16, 0); kstep_a = 256; kstep_b = 16; }
        else if (mode == 5) { ad = make_desc(a, 2048, 128, 0); bd = make_desc(b, 2064, 128, 0); kstep_a = 256; kstep_b = 258; }
        else { ad = make_desc(a, 2048, 128, 0); bd = make_desc(b, 2048, 128, 0); kstep_a = kstep_b = 256; }
        const long long t0 = clock64();
        for (int i = 0; i < n_mma; i++) {
            const uint32_t k = i & 3;      // walk 4 K-steps of a 64-wide tile
            umma(tmem + (i % nacc) * N, ad + k * kstep_a, bd + k * kstep_b, idesc, i >= nacc ? 1u : 0u);
        }
        commit(&bar);
        const long long t1 = clock64();
        while (!try_wait(&bar, 0)) {}
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 4096);
    const char *names[] = {"none K-major", "SW128 K-major", "SW64 K-major", "SW32 K-major", "none, B MN-major", "none K-major padded LBO"};
    for (int N : {128, 256}) {
        for (int nacc : {1, 2}) {
            for (int mode = 0; mode < 6; mode++) {
                for (int rep = 0; rep < 2; rep++) {
                    rate_kernel<<<148, 128, 131072 + 2048>>>(mode, 2048, nacc, N, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
                }
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("N=%d acc=%d %-26s issue %.1f cyc/mma   complete %.1f cyc/mma\n", N, nacc, names[mode], h[0] / 2048.0, h[1] / 2048.0);
            }
        }
    }
    return 0;
}
