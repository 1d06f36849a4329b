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
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
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
__global__ void __launch_bounds__(128) rate_kernel(int mode, int n_mma, int nacc, int N, int M, int same_a, long long *out,
                                                   const unsigned char *gsrc, int tma_on) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint64_t cbar[4];
    __shared__ volatile int stop_flag;
    __shared__ uint32_t tmem_slot;
    unsigned char *A = smem, *B = smem + 65536;
    for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar[i])));
        stop_flag = 0;
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
    if (threadIdx.x < 32) {
        // warp-uniform issue loop, one elected lane per group of 4 MMAs (the pattern libp2w's conv_tc uses)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((mode == 4 ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
                               ((uint32_t)(M >> 4) << 24);
        uint64_t ad, bd;
        uint32_t kstep_a, kstep_b;          // descriptor increment per K=16 step (in 16-byte units)
        const uint32_t a = smem_u32(A), b = smem_u32(B);
        if (mode == 1) { ad = make_desc(a, 16, 1024, 2); bd = make_desc(b, 16, 1024, 2); kstep_a = kstep_b = 2; }
        else if (mode == 4) { ad = make_desc(a, 2048, 128, 0); bd = make_desc(b, 128, 192 * 16, 0); kstep_a = 256; kstep_b = 16; }
        else if (mode == 5) { ad = make_desc(a, 2048, 128, 0); bd = make_desc(b, 2064, 128, 0); kstep_a = 256; kstep_b = 258; }
        else if (mode == 0) { ad = make_desc(a, 2048, 128, 0); bd = make_desc(b, 2048, 128, 0); kstep_a = kstep_b = 256; }
        else { ad = make_desc(a, 2048, 128, 0); bd = make_desc(b, 4096, 128, 0); kstep_a = 256; kstep_b = 512; }
        const uint32_t d0 = tmem, d1 = tmem + (nacc > 1 ? N : 0);
        const uint64_t a2 = ad + 2 * kstep_a, a3 = ad + 3 * kstep_a;
        const uint64_t b2 = bd + 2 * kstep_b, b3 = bd + 3 * kstep_b;
        __syncwarp();
        const long long t0 = clock64();
        // same_a doubles as the overhead mode here: bit 0 = tcgen05.commit after every group of four MMAs (the ring
        // slot release of conv_tc), bit 1 = a (satisfied) mbarrier try_wait before every group (the ring_full wait)
        const int ovh = same_a;
        uint32_t rdy_phase = 0;
        if (ovh & 2) {          // cbar[1] completes one phase now; waiting on parity 0 is satisfied from then on
            if (elect_one()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&cbar[1])) : "memory");
            __syncwarp();
        }
        for (int i = 0; i < n_mma; i += 4) {
            if (ovh & 2) { while (!try_wait(&cbar[1], rdy_phase)) {} }
            if (elect_one()) {
                umma(d0, ad, bd, idesc, i ? 1u : 0u);
                umma(d1, ad + kstep_a, bd + kstep_b, idesc, (i || nacc == 1) ? 1u : 0u);
                umma(d0, a2, b2, idesc, 1u);
                umma(d1, a3, b3, idesc, 1u);
                if (ovh & 1) commit(&cbar[0]);
            }
            __syncwarp();
        }
        if (elect_one()) commit(&bar);
        const long long t1 = clock64();
        while (!try_wait(&bar, 0)) {}
        const long long t2 = clock64();
        stop_flag = 1;
        if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    } else if (threadIdx.x == 32 && tma_on) {
        // concurrent writer: 8 KB bulk copies global -> shared, four in flight (the weight ring of conv_tc)
        long long copies = 0;
        uint32_t ph[4] = {0, 0, 0, 0};
        unsigned char *dst = smem + 140 * 1024;
        for (int i = 0; !stop_flag; i++) {
            const int sl = i & 3;
            if (i >= 4) { while (!try_wait(&cbar[sl], ph[sl])) {} ph[sl] ^= 1; }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cbar[sl])), "r"(8192) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst + sl * 8192)),
                         "l"(gsrc + (size_t)((i * 37 + blockIdx.x) & 63) * 8192), "r"(8192), "r"(smem_u32(&cbar[sl])) : "memory");
            copies++;
        }
        if (blockIdx.x == 0) out[2] = copies;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long *d, h[3];
    unsigned char *g;
    cudaMalloc(&d, 24);
    cudaMalloc(&g, 64 * 8192);
    cudaMemset(g, 0, 64 * 8192);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int N : {128}) {
        for (int ovh = 0; ovh < 4; ovh++) {
            for (int rep = 0; rep < 2; rep++) {
                cudaMemset(d, 0, 24);
                rate_kernel<<<148, 128, 200 * 1024>>>(6, 8192, 2, N, 128, ovh, d, g, 0);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
            printf("M=128 N=%d  groups of 4 MMAs%s%s: %.1f cyc/mma (floor %d)\n", N, (ovh & 2) ? " + satisfied try_wait before" : "",
                   (ovh & 1) ? " + tcgen05.commit after" : "", h[1] / 8192.0, N / 2);
        }
    }
    return 0;
}
