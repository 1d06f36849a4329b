// dense.cu -- per-channel affine + ReLU epilogues between the dense per-point GEMMs of
// InvertedResidualBlock (src/model.py:18-85).  In eval mode every BatchNorm1d that follows a
// k=1 convolution folds into the convolution's weights; what is left between two GEMMs is
//     depthwise(k=1) -> BN -> ReLU                       y = relu(x * s1 + t1)
//     BN -> ReLU -> depthwise(k=1) -> BN -> ReLU         y = relu(relu(x * s1 + t1) * s2 + t2)
// on [N, 4C] activations.  These are HBM-bound streaming kernels: one pass, 16-byte accesses,
// FP32 or BF16 activations, FP32 per-channel constants.
#include <cuda_bf16.h>

#include "common.cuh"

namespace p2w {
namespace {

template <bool TWO>
__device__ __forceinline__ float chain(float v, float s1, float t1, float s2, float t2) {
    v = fmaxf(fmaf(v, s1, t1), 0.f);
    if (TWO) v = fmaxf(fmaf(v, s2, t2), 0.f);
    return v;
}

// A thread owns ONE group of 8 channels and walks down the rows: its 16 / 32 per-channel constants stay in
// registers, and consecutive threads cover consecutive 16-byte pieces of a row (coalesced); two rows per
// iteration keep two loads in flight per thread.
template <bool TWO, bool BF16>
__global__ void __launch_bounds__(256) affine_relu_kernel(const void *__restrict__ xin, void *__restrict__ yout,
                                                          int64_t n, int c, const float *__restrict__ s1,
                                                          const float *__restrict__ t1, const float *__restrict__ s2,
                                                          const float *__restrict__ t2) {
    const int c8 = c >> 3;
    const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t nthreads = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int g = static_cast<int>(tid % c8);                     // nthreads is a multiple of c8 (see the launch)
    const int64_t row0 = tid / c8, row_step = nthreads / c8;
    const int ch = g * 8;
    float a1[8], b1[8], a2[8], b2[8];
    *reinterpret_cast<float4 *>(a1) = __ldg(reinterpret_cast<const float4 *>(s1 + ch));
    *reinterpret_cast<float4 *>(a1 + 4) = __ldg(reinterpret_cast<const float4 *>(s1 + ch + 4));
    *reinterpret_cast<float4 *>(b1) = __ldg(reinterpret_cast<const float4 *>(t1 + ch));
    *reinterpret_cast<float4 *>(b1 + 4) = __ldg(reinterpret_cast<const float4 *>(t1 + ch + 4));
    if (TWO) {
        *reinterpret_cast<float4 *>(a2) = __ldg(reinterpret_cast<const float4 *>(s2 + ch));
        *reinterpret_cast<float4 *>(a2 + 4) = __ldg(reinterpret_cast<const float4 *>(s2 + ch + 4));
        *reinterpret_cast<float4 *>(b2) = __ldg(reinterpret_cast<const float4 *>(t2 + ch));
        *reinterpret_cast<float4 *>(b2 + 4) = __ldg(reinterpret_cast<const float4 *>(t2 + ch + 4));
    }
    for (int64_t r = row0; r < n; r += 2 * row_step) {
        const int64_t i0 = r * c8 + g, i1 = (r + row_step) * c8 + g;
        const bool second = r + row_step < n;
        if (BF16) {
            uint4 raw0 = reinterpret_cast<const uint4 *>(xin)[i0];
            uint4 raw1 = second ? reinterpret_cast<const uint4 *>(xin)[i1] : make_uint4(0, 0, 0, 0);
            __nv_bfloat162 *h0 = reinterpret_cast<__nv_bfloat162 *>(&raw0), *h1 = reinterpret_cast<__nv_bfloat162 *>(&raw1);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                float2 f = __bfloat1622float2(h0[u]), e = __bfloat1622float2(h1[u]);
                f.x = chain<TWO>(f.x, a1[2 * u], b1[2 * u], a2[2 * u], b2[2 * u]);
                f.y = chain<TWO>(f.y, a1[2 * u + 1], b1[2 * u + 1], a2[2 * u + 1], b2[2 * u + 1]);
                e.x = chain<TWO>(e.x, a1[2 * u], b1[2 * u], a2[2 * u], b2[2 * u]);
                e.y = chain<TWO>(e.y, a1[2 * u + 1], b1[2 * u + 1], a2[2 * u + 1], b2[2 * u + 1]);
                h0[u] = __floats2bfloat162_rn(f.x, f.y);
                h1[u] = __floats2bfloat162_rn(e.x, e.y);
            }
            reinterpret_cast<uint4 *>(yout)[i0] = raw0;
            if (second) reinterpret_cast<uint4 *>(yout)[i1] = raw1;
        } else {
#pragma unroll
            for (int half = 0; half < 2; half++) {
                if (half && !second) break;
                const int64_t i = half ? i1 : i0;
                float v[8];
                *reinterpret_cast<float4 *>(v) = reinterpret_cast<const float4 *>(xin)[2 * i];
                *reinterpret_cast<float4 *>(v + 4) = reinterpret_cast<const float4 *>(xin)[2 * i + 1];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = chain<TWO>(v[u], a1[u], b1[u], a2[u], b2[u]);
                reinterpret_cast<float4 *>(yout)[2 * i] = *reinterpret_cast<float4 *>(v);
                reinterpret_cast<float4 *>(yout)[2 * i + 1] = *reinterpret_cast<float4 *>(v + 4);
            }
        }
    }
}

// out[r] = dot(x[r, :], w) + bias: the 1-channel head (conv2, src/model.py:243).  Warp per row, 16-byte loads,
// weights in registers (c <= 1024): one pass over x at HBM speed instead of a GEMM with N = 1.
template <bool BF16>
__global__ void __launch_bounds__(256) rowdot_kernel(const void *__restrict__ xin, int64_t n, int c,
                                                     const float *__restrict__ w, float bias, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int c8 = c >> 3;
    float wr[4][8];                                   // this lane's channel groups: lane, lane + 32, ...
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int g = lane + 32 * u;
#pragma unroll
        for (int e = 0; e < 8; e++) wr[u][e] = g < c8 ? __ldg(w + g * 8 + e) : 0.f;
    }
    for (int64_t r = warp; r < n; r += nwarps) {
        float acc = 0.f;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int g = lane + 32 * u;
            if (g < c8) {
                float v[8];
                if (BF16) {
                    const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(xin) + r * c8 + g);
                    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float2 f = __bfloat1622float2(h[e]);
                        v[2 * e] = f.x;
                        v[2 * e + 1] = f.y;
                    }
                } else {
                    *reinterpret_cast<float4 *>(v) = __ldg(reinterpret_cast<const float4 *>(xin) + 2 * (r * c8 + g));
                    *reinterpret_cast<float4 *>(v + 4) = __ldg(reinterpret_cast<const float4 *>(xin) + 2 * (r * c8 + g) + 1);
                }
#pragma unroll
                for (int e = 0; e < 8; e++) acc = fmaf(v[e], wr[u][e], acc);
            }
        }
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[r] = acc + bias;
    }
}

// out = relu(a + b) over n elements of FP32 / BF16 rows, 16 bytes per thread: the shortcut add + ReLU that closes
// InvertedResidualBlock (src/model.py:84) in one pass instead of an add and a clamp.
template <bool BF16>
__global__ void __launch_bounds__(256) add_relu_kernel(const void *__restrict__ a, const void *__restrict__ b,
                                                       void *__restrict__ out, int64_t nvec) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const uint4 va = __ldg(reinterpret_cast<const uint4 *>(a) + i), vb = __ldg(reinterpret_cast<const uint4 *>(b) + i);
        uint4 vo;
        if (BF16) {
            const __nv_bfloat162 *pa = reinterpret_cast<const __nv_bfloat162 *>(&va);
            const __nv_bfloat162 *pb = reinterpret_cast<const __nv_bfloat162 *>(&vb);
            __nv_bfloat162 *po = reinterpret_cast<__nv_bfloat162 *>(&vo);
            const __nv_bfloat162 zero = __float2bfloat162_rn(0.f);
#pragma unroll
            for (int e = 0; e < 4; e++) po[e] = __hmax2(__hadd2(pa[e], pb[e]), zero);   // one bf16 rounding, as torch's add_
        } else {
            const float *pa = reinterpret_cast<const float *>(&va), *pb = reinterpret_cast<const float *>(&vb);
            float *po = reinterpret_cast<float *>(&vo);
#pragma unroll
            for (int e = 0; e < 4; e++) po[e] = fmaxf(pa[e] + pb[e], 0.f);
        }
        reinterpret_cast<uint4 *>(out)[i] = vo;
    }
}

}  // namespace
}  // namespace p2w

using namespace p2w;

extern "C" int p2w_rowdot(const void *x, int32_t dtype, int64_t n, int32_t c, const float *w, float bias, float *out,
                          p2w_stream_t stream) {
    P2W_REQUIRE(c >= 8 && c % 8 == 0 && c <= 1024, "p2w_rowdot: c=%d must be a multiple of 8, at most 1024", c);
    P2W_REQUIRE(dtype == P2W_F32 || dtype == P2W_BF16, "p2w_rowdot: unknown dtype %d", dtype);
    P2W_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15u) == 0, "p2w_rowdot: rows must be 16-byte aligned");
    if (n == 0) return P2W_OK;
    int64_t blocks = (n * 32 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cudaStream_t st = as_stream(stream);
    if (dtype == P2W_BF16) P2W_LAUNCH(rowdot_kernel<true>, (unsigned)blocks, 256, 0, st)(x, n, c, w, bias, out);
    else P2W_LAUNCH(rowdot_kernel<false>, (unsigned)blocks, 256, 0, st)(x, n, c, w, bias, out);
    return check_launch("p2w_rowdot");
}

extern "C" int p2w_add_relu(const void *a, const void *b, void *out, int64_t n, int32_t dtype, p2w_stream_t stream) {
    P2W_REQUIRE(dtype == P2W_F32 || dtype == P2W_BF16, "p2w_add_relu: unknown dtype %d", dtype);
    const int per = dtype == P2W_BF16 ? 8 : 4;
    P2W_REQUIRE(n >= 0 && n % per == 0, "p2w_add_relu: n=%lld must be a multiple of %d", (long long)n, per);
    P2W_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0,
                "p2w_add_relu: operands must be 16-byte aligned");
    if (n == 0) return P2W_OK;
    const int64_t nvec = n / per;
    int64_t blocks = (nvec + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cudaStream_t st = as_stream(stream);
    if (dtype == P2W_BF16) P2W_LAUNCH(add_relu_kernel<true>, (unsigned)blocks, 256, 0, st)(a, b, out, nvec);
    else P2W_LAUNCH(add_relu_kernel<false>, (unsigned)blocks, 256, 0, st)(a, b, out, nvec);
    return check_launch("p2w_add_relu");
}

extern "C" int p2w_affine_relu(const void *x, void *y, int64_t n, int32_t c, const float *s1, const float *t1,
                               const float *s2, const float *t2, int32_t dtype, p2w_stream_t stream) {
    P2W_REQUIRE(c >= 8 && c % 8 == 0, "p2w_affine_relu: c=%d must be a multiple of 8", c);
    P2W_REQUIRE(dtype == P2W_F32 || dtype == P2W_BF16, "p2w_affine_relu: unknown dtype %d", dtype);
    P2W_REQUIRE((s2 == nullptr) == (t2 == nullptr), "p2w_affine_relu: s2 and t2 go together");
    P2W_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(s1) |
                  reinterpret_cast<uintptr_t>(t1) | reinterpret_cast<uintptr_t>(s2) | reinterpret_cast<uintptr_t>(t2)) &
                 15u) == 0,
                "p2w_affine_relu: pointers must be 16-byte aligned");
    if (n == 0) return P2W_OK;
    cudaStream_t st = as_stream(stream);
    // the thread count is a multiple of the channel groups per row, so that a thread keeps its group
    const int c8 = c >> 3;
    const int64_t total = n * c8;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    {
        int64_t lcm_blocks = c8;                       // blocks * 256 % c8 == 0  <=  blocks % (c8 / gcd(c8, 256)) == 0
        int64_t a = c8, b = 256;
        while (b) { const int64_t t = a % b; a = b; b = t; }
        lcm_blocks = c8 / a;
        blocks = (blocks + lcm_blocks - 1) / lcm_blocks * lcm_blocks;
    }
    const bool two = s2 != nullptr, bf = dtype == P2W_BF16;
    if (two && bf) P2W_LAUNCH((affine_relu_kernel<true, true>), (unsigned)blocks, 256, 0, st)(x, y, n, c, s1, t1, s2, t2);
    else if (two) P2W_LAUNCH((affine_relu_kernel<true, false>), (unsigned)blocks, 256, 0, st)(x, y, n, c, s1, t1, s2, t2);
    else if (bf) P2W_LAUNCH((affine_relu_kernel<false, true>), (unsigned)blocks, 256, 0, st)(x, y, n, c, s1, t1, s2, t2);
    else P2W_LAUNCH((affine_relu_kernel<false, false>), (unsigned)blocks, 256, 0, st)(x, y, n, c, s1, t1, s2, t2);
    return check_launch("p2w_affine_relu");
}
