// Experiment (not part of libp2w): per-SM throughput of a 32 x 32 "column max" -- every lane holds 32 FP32 values,
// lane c ends up with the maximum of value c over the warp -- by (0) redux.sync.max.f32 (CREDUX) + select and
// (1) a shuffle butterfly (31 SHFL + 31 FMNMX + selects).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float redux_max(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}

__global__ void __launch_bounds__(256) rate_kernel(int mode, int iters, long long *out, float *sinkp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = (float)((threadIdx.x * 37 + i * 11) & 255);
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        float res = 0.f;
        if (mode == 0) {
#pragma unroll
            for (int c = 0; c < 32; c++) {
                const float r = redux_max(v[c]);
                res = lane == c ? r : res;
            }
        } else {
            float w[32];
#pragma unroll
            for (int i = 0; i < 32; i++) w[i] = v[i];
#pragma unroll
            for (int s = 16; s; s >>= 1) {
                const bool up = lane & s;
#pragma unroll
                for (int i = 0; i < s; i++) {
                    const float send = up ? w[i] : w[i + s], keep = up ? w[i + s] : w[i];
                    w[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, s));
                }
            }
            res = w[0];
        }
        acc += res;
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] += res * 1e-9f;      // loop-carried, keeps every iteration live
    }
    const long long t1 = clock64();
    if (acc == 123.456f) *sinkp = acc;
    if (blockIdx.x == 0 && lane == 0) out[warp] = t1 - t0;
}

int main() {
    long long *d, h[8];
    float *s;
    cudaMalloc(&d, 64);
    cudaMalloc(&s, 4);
    const char *names[] = {"CREDUX.MAX.F32 + select", "shuffle butterfly"};
    for (int warps : {1, 4, 8}) {
        for (int mode = 0; mode < 2; mode++) {
            for (int rep = 0; rep < 2; rep++) {
                rate_kernel<<<148, warps * 32>>>(mode, 2048, d, s);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int w = 0; w < warps; w++) mx = h[w] > mx ? h[w] : mx;
            printf("warps=%d %-26s %.1f cyc per 32x32 block per warp (incl. 32 FFMA), %.2f cyc per column-max per SM\n", warps,
                   names[mode], mx / 2048.0, mx / 2048.0 / 32.0 / warps);
        }
    }
    return 0;
}
