// Experiment (not part of libp2w): throughput of tcgen05.ld.32x32b.xN (TMEM -> registers) per SM as a function of
// the number of warps reading (one lane quarter each; warps w and w+4 share a quarter) and the shape xN.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_rate tmem_ld_rate.cu ; run on a B200.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t &sink) {
    if constexpr (X == 32) {
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                       "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                       "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; i++) sink ^= r[i];
    } else {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; i++) sink ^= r[i];
    }
}

// two loads in flight before the wait (what a pipelined epilogue would do)
__device__ __forceinline__ void ld2(uint32_t taddr, uint32_t &sink) {
    uint32_t r[32], q[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9]),
                   "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]), "=r"(q[17]), "=r"(q[18]),
                   "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]), "=r"(q[25]), "=r"(q[26]), "=r"(q[27]),
                   "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
                 : "r"(taddr + 32));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) sink ^= r[i] ^ q[i];
}

__global__ void __launch_bounds__(256) rate_kernel(int mode, int iters, long long *out, uint32_t *sinkp) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t sink = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        const uint32_t col = (uint32_t)((i * 64 + (warp >> 2) * 256) & 511);
        if (mode == 0) ld<32>(tmem + (col & 480), sink);
        else if (mode == 1) ld<16>(tmem + (col & 496), sink);
        else ld2(tmem + (col & 448), sink);
    }
    const long long t1 = clock64();
    if (sink == 0x12345678u) *sinkp = sink;
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[warp] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512) : "memory");
}

int main() {
    long long *d, h[8];
    uint32_t *s;
    cudaMalloc(&d, 64);
    cudaMalloc(&s, 4);
    const char *names[] = {"x32, wait each", "x16, wait each", "2 x x32, then wait"};
    const int bytes[] = {4096, 2048, 8192};
    for (int warps : {1, 4, 8}) {
        for (int mode = 0; mode < 3; mode++) {
            for (int rep = 0; rep < 2; rep++) {
                rate_kernel<<<148, warps * 32, 0>>>(mode, 4096, d, s);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int w = 0; w < warps; w++) mx = h[w] > mx ? h[w] : mx;
            printf("warps=%d %-20s %.1f cyc per warp-load, %.1f B/cyc per SM\n", warps, names[mode], mx / 4096.0,
                   (double)bytes[mode] * warps * 4096.0 / mx);
        }
    }
    return 0;
}
