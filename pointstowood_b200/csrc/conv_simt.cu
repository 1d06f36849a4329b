// conv_simt.cu -- K5 in FP32: fused gather -> per-edge MLP -> max (the 1e-3 parity mode).
//
// Replaces MessagePassing.propagate(aggr='max') around PointNetConv.message
// (src/pointnet.py:108,116-132).  One 128-thread CTA owns one target and its <= 32 edges:
// lane = edge.  The gathered message tile [32, C+4] and the hidden tile [32, H] live in
// shared memory (odd row strides: conflict-free column reads); the weights are read as
// warp-uniform 128-bit loads from L1/L2 in a [K][N] layout, 8 output channels per lane in
// registers.  BatchNorm is applied per edge BEFORE the max (its scale may be negative),
// padded / missing edges are replaced by a duplicate of a valid one, so the [E, C'] edge
// tensor never exists in HBM.
#include <cuda_bf16.h>

#include <stdlib.h>

#include "common.cuh"

namespace p2w {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int CONV_WARPS = 4;

__global__ void transpose_kernel(const float *__restrict__ w, int rows, int cols, float *__restrict__ wt) {
    // w [rows, cols] -> wt [cols, rows]
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<int64_t>(rows) * cols) return;
    const int r = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
    wt[static_cast<int64_t>(c) * rows + r] = w[i];
}

// acc[8] += a * wt[k][n0..n0+8) for k in [0, K)
__device__ __forceinline__ void dot8(const float *__restrict__ tile_row, int K, const float *__restrict__ wt, int N,
                                     int n0, float acc[8]) {
#pragma unroll 4
    for (int k = 0; k < K; k++) {
        const float a = tile_row[k];
        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(wt + static_cast<int64_t>(k) * N + n0));
        const float4 w1 = __ldg(reinterpret_cast<const float4 *>(wt + static_cast<int64_t>(k) * N + n0 + 4));
        acc[0] = fmaf(a, w0.x, acc[0]);
        acc[1] = fmaf(a, w0.y, acc[1]);
        acc[2] = fmaf(a, w0.z, acc[2]);
        acc[3] = fmaf(a, w0.w, acc[3]);
        acc[4] = fmaf(a, w1.x, acc[4]);
        acc[5] = fmaf(a, w1.y, acc[5]);
        acc[6] = fmaf(a, w1.z, acc[6]);
        acc[7] = fmaf(a, w1.w, acc[7]);
    }
}

__global__ void __launch_bounds__(CONV_WARPS * 32)
    conv_simt_kernel(const float *__restrict__ x, const float *__restrict__ pos_src, const float *__restrict__ pos_tgt,
                     const int32_t *__restrict__ nbr, int64_t n_tgt, int K, int C, int H, int Co,
                     const float *__restrict__ w1t, const float *__restrict__ b1, const float *__restrict__ w2t,
                     const float *__restrict__ b2, const float *__restrict__ bn_scale,
                     const float *__restrict__ bn_shift, float *__restrict__ out) {
    extern __shared__ float smem[];
    const int K1 = C + 4;
    const int ldm = K1 | 1, ldh = H | 1;
    float *msg = smem;                 // [32][ldm]
    float *hid = smem + 32 * ldm;      // [32][ldh]
    __shared__ int s_j[32];
    __shared__ int s_any;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (int64_t t = blockIdx.x; t < n_tgt; t += gridDim.x) {
        // ---- edges of this target; padded slots duplicate the first valid neighbour
        if (warp == 0) {
            int j = (lane < K) ? nbr[t * K + lane] : -1;
            const unsigned m = __ballot_sync(FULL, j >= 0);
            if (m) {
                const int jf = __shfl_sync(FULL, j, __ffs(m) - 1);
                if (j < 0) j = jf;
                const float dx = pos_src[static_cast<int64_t>(j) * 4 + 0] - pos_tgt[t * 4 + 0];
                const float dy = pos_src[static_cast<int64_t>(j) * 4 + 1] - pos_tgt[t * 4 + 1];
                const float dz = pos_src[static_cast<int64_t>(j) * 4 + 2] - pos_tgt[t * 4 + 2];
                float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
                for (int o = 16; o; o >>= 1) nrm = fmaxf(nrm, __shfl_xor_sync(FULL, nrm, o));
                const float den = nrm + 1e-8f;
                msg[lane * ldm + C + 0] = dx / den;
                msg[lane * ldm + C + 1] = dy / den;
                msg[lane * ldm + C + 2] = dz / den;
                msg[lane * ldm + C + 3] = pos_src[static_cast<int64_t>(j) * 4 + 3];
            }
            s_j[lane] = j;
            if (lane == 0) s_any = m ? 1 : 0;
        }
        __syncthreads();
        if (!s_any) {   // no edge: PyG's max aggregation leaves 0
            for (int c = threadIdx.x; c < Co; c += CONV_WARPS * 32) out[t * Co + c] = 0.f;
            __syncthreads();
            continue;
        }
        // ---- gather x_j rows (coalesced along channels)
        for (int e = warp; e < 32; e += CONV_WARPS) {
            const float *row = x + static_cast<int64_t>(s_j[e]) * C;
            for (int c = lane; c < C; c += 32) msg[e * ldm + c] = row[c];
        }
        __syncthreads();
        // ---- layer 1: hid = relu(msg W1^T + b1)
        for (int blk = warp; blk < H / 8; blk += CONV_WARPS) {
            const int n0 = blk * 8;
            float acc[8];
#pragma unroll
            for (int u = 0; u < 8; u++) acc[u] = b1[n0 + u];
            dot8(msg + lane * ldm, K1, w1t, H, n0, acc);
#pragma unroll
            for (int u = 0; u < 8; u++) hid[lane * ldh + n0 + u] = fmaxf(acc[u], 0.f);
        }
        __syncthreads();
        // ---- layer 2 + ReLU + BN, then max over the 32 edges
        for (int blk = warp; blk < Co / 8; blk += CONV_WARPS) {
            const int n0 = blk * 8;
            float acc[8];
#pragma unroll
            for (int u = 0; u < 8; u++) acc[u] = b2[n0 + u];
            dot8(hid + lane * ldh, H, w2t, Co, n0, acc);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                float v = fmaf(fmaxf(acc[u], 0.f), bn_scale[n0 + u], bn_shift[n0 + u]);
                for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
                acc[u] = v;
            }
#pragma unroll
            for (int u = 0; u < 8; u++)
                if (lane == u) out[t * Co + n0 + u] = acc[u];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ knn_interpolate
// warp per query; lanes over channels.  w = 1 / max(|p_x - p_y|^2, 1e-16).  TI / TO: float or bf16.
__device__ __forceinline__ float ld_f(const float *p) { return *p; }
__device__ __forceinline__ float ld_f(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_f(float *p, float v) { *p = v; }
__device__ __forceinline__ void st_f(__nv_bfloat16 *p, float v) { *p = __float2bfloat16(v); }

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) interp_kernel(const TI *__restrict__ x, const float *__restrict__ pos_x,
                                                     const float *__restrict__ pos_y, const int32_t *__restrict__ nbr,
                                                     int64_t ny, int k, int c, int ld_out, TO *__restrict__ out) {
    const int64_t q = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= ny) return;
    int j = -1;
    float w = 0.f;
    if (lane < k) {
        j = nbr[q * k + lane];
        if (j >= 0) {
            const float dx = __fsub_rn(pos_x[static_cast<int64_t>(j) * 3 + 0], pos_y[q * 3 + 0]);
            const float dy = __fsub_rn(pos_x[static_cast<int64_t>(j) * 3 + 1], pos_y[q * 3 + 1]);
            const float dz = __fsub_rn(pos_x[static_cast<int64_t>(j) * 3 + 2], pos_y[q * 3 + 2]);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            w = __fdiv_rn(1.0f, fmaxf(d2, 1e-16f));
        }
    }
    float den = 0.f;
    for (int e = 0; e < k; e++) den = __fadd_rn(den, __shfl_sync(FULL, w, e));
    for (int c0 = 0; c0 < c; c0 += 32) {
        const int ch = c0 + lane;
        float num = 0.f;
        for (int e = 0; e < k; e++) {
            const int je = __shfl_sync(FULL, j, e);
            const float we = __shfl_sync(FULL, w, e);
            if (je >= 0 && ch < c) num = __fadd_rn(num, __fmul_rn(ld_f(x + static_cast<int64_t>(je) * c + ch), we));
        }
        if (ch < c) st_f(out + q * ld_out + ch, __fdiv_rn(num, den));
    }
}

// FPModule.forward (src/model.py:148-153) up to its MLP: out[q] = [knn_interpolate(x)[q], x_skip[q]] in one
// pass.  Warp per query row; a lane owns groups of 8 channels (16-byte accesses on bf16 rows), so the
// k source rows stream through full sectors and the concatenated row is written once.
__device__ __forceinline__ void ld8(const float *p, float v[8]) {
    *reinterpret_cast<float4 *>(v) = __ldg(reinterpret_cast<const float4 *>(p));
    *reinterpret_cast<float4 *>(v + 4) = __ldg(reinterpret_cast<const float4 *>(p) + 1);
}
__device__ __forceinline__ void ld8(const __nv_bfloat16 *p, float v[8]) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(p));
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const float2 f = __bfloat1622float2(h[u]);
        v[2 * u] = f.x;
        v[2 * u + 1] = f.y;
    }
}
__device__ __forceinline__ void st8(float *p, const float v[8]) {
    reinterpret_cast<float4 *>(p)[0] = *reinterpret_cast<const float4 *>(v);
    reinterpret_cast<float4 *>(p)[1] = *reinterpret_cast<const float4 *>(v + 4);
}
__device__ __forceinline__ void st8(__nv_bfloat16 *p, const float v[8]) {
    uint4 raw;
    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&raw);
#pragma unroll
    for (int u = 0; u < 4; u++) h[u] = __floats2bfloat162_rn(v[2 * u], v[2 * u + 1]);
    *reinterpret_cast<uint4 *>(p) = raw;
}

template <typename TI, typename TS, typename TO>
__global__ void __launch_bounds__(256) interp_cat_kernel(const TI *__restrict__ x, const float *__restrict__ pos_x,
                                                         const float *__restrict__ pos_y,
                                                         const int32_t *__restrict__ nbr, int64_t ny, int k, int c,
                                                         const TS *__restrict__ skip, int cs, int ld_out,
                                                         TO *__restrict__ out) {
    const int64_t q = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= ny) return;
    int j = -1;
    float w = 0.f;
    if (lane < k) {
        j = nbr[q * k + lane];
        if (j >= 0) {
            const float dx = __fsub_rn(pos_x[static_cast<int64_t>(j) * 3 + 0], pos_y[q * 3 + 0]);
            const float dy = __fsub_rn(pos_x[static_cast<int64_t>(j) * 3 + 1], pos_y[q * 3 + 1]);
            const float dz = __fsub_rn(pos_x[static_cast<int64_t>(j) * 3 + 2], pos_y[q * 3 + 2]);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            w = __fdiv_rn(1.0f, fmaxf(d2, 1e-16f));
        }
    }
    float den = 0.f;
    for (int e = 0; e < k; e++) den = __fadd_rn(den, __shfl_sync(FULL, w, e));
    const int groups = (c + cs) >> 3, gi = c >> 3;
    // BF16 output: the weights are normalised once and the sum is a fused multiply-add chain (the rounding to
    // bf16 swamps the difference); FP32 output keeps upstream's sum-then-divide, operation for operation.
    constexpr bool kFast = sizeof(TO) == 2;
    const float wn = kFast ? __fdividef(w, den) : w;
    for (int g0 = 0; g0 < groups; g0 += 32) {
        const int g = g0 + lane;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int e = 0; e < k; e++) {          // warp-uniform trip count: the shuffles stay converged
            const int je = __shfl_sync(FULL, j, e);
            const float we = __shfl_sync(FULL, wn, e);
            if (je >= 0 && g < gi) {
                float v[8];
                ld8(x + static_cast<int64_t>(je) * c + g * 8, v);
#pragma unroll
                for (int u = 0; u < 8; u++) acc[u] = kFast ? fmaf(v[u], we, acc[u]) : __fadd_rn(acc[u], __fmul_rn(v[u], we));
            }
        }
        if (g < gi) {
            if (!kFast) {
#pragma unroll
                for (int u = 0; u < 8; u++) acc[u] = __fdiv_rn(acc[u], den);
            }
            st8(out + q * ld_out + g * 8, acc);
        } else if (g < groups) {
            ld8(skip + q * cs + (g - gi) * 8, acc);
            st8(out + q * ld_out + g * 8, acc);
        }
    }
}

// out[q, :] = act( knn_interpolate(y)[q, :] + z[q, :] ): the first Linear of an FPModule applied BEFORE the
// interpolation.  knn_interpolate is linear with weights that sum to one, so
//     relu([interp(x), x_skip] W^T + b) = relu(interp(x Wc^T) + (x_skip Ws^T + b)),
// and x Wc^T runs over the COARSE rows (2.3-4x fewer than the fine ones) while the [n_fine, C + C_skip] concatenation
// is never built.  One warp per fine row, 16-byte channel groups; out may alias z.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) interp_add_kernel(const TI *__restrict__ y, const float *__restrict__ pos_x,
                                                         const float *__restrict__ pos_y,
                                                         const int32_t *__restrict__ nbr, int64_t ny, int k, int c,
                                                         const TO *z, int relu, TO *out) {
    const int64_t q = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= ny) return;
    int j = -1;
    float w = 0.f;
    if (lane < k) {
        j = nbr[q * k + lane];
        if (j >= 0) {
            const float dx = __fsub_rn(pos_x[static_cast<int64_t>(j) * 3 + 0], pos_y[q * 3 + 0]);
            const float dy = __fsub_rn(pos_x[static_cast<int64_t>(j) * 3 + 1], pos_y[q * 3 + 1]);
            const float dz = __fsub_rn(pos_x[static_cast<int64_t>(j) * 3 + 2], pos_y[q * 3 + 2]);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            w = __fdiv_rn(1.0f, fmaxf(d2, 1e-16f));
        }
    }
    float den = 0.f;
    for (int e = 0; e < k; e++) den = __fadd_rn(den, __shfl_sync(FULL, w, e));
    const float wn = den > 0.f ? __fdiv_rn(w, den) : 0.f;
    const int groups = c >> 3;
    for (int g0 = 0; g0 < groups; g0 += 32) {
        const int g = g0 + lane;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int e = 0; e < k; e++) {          // warp-uniform trip count: the shuffles stay converged
            const int je = __shfl_sync(FULL, j, e);
            const float we = __shfl_sync(FULL, wn, e);
            if (je >= 0 && g < groups) {
                float v[8];
                ld8(y + static_cast<int64_t>(je) * c + g * 8, v);
#pragma unroll
                for (int u = 0; u < 8; u++) acc[u] = fmaf(v[u], we, acc[u]);
            }
        }
        if (g < groups) {
            float zz[8];
            ld8(z + q * c + g * 8, zz);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const float t = acc[u] + zz[u];
                acc[u] = relu ? fmaxf(t, 0.f) : t;
            }
            st8(out + q * c + g * 8, acc);
        }
    }
}

// ------------------------------------------------------------------ segment max (global_max_pool)
__global__ void __launch_bounds__(128) segment_max_kernel(const float *__restrict__ x, const int64_t *__restrict__ ptr,
                                                          int c, float *__restrict__ out) {
    const int b = blockIdx.x;
    const int ch = blockIdx.y * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    const int64_t r0 = ptr[b], r1 = ptr[b + 1];
    float m = __int_as_float(0xff800000);
    for (int64_t r = r0; r < r1; r++) m = fmaxf(m, x[r * c + ch]);
    out[static_cast<int64_t>(b) * c + ch] = (r1 > r0) ? m : 0.f;
}

// rows in FP32 or BF16, optional per-channel affine applied to every element before the max (the eval-mode
// BatchNorm that follows the last MLP of GlobalSAModule, src/model.py:134-136): one pass over the rows instead
// of a cast, a multiply and an add over [n, c] in front of the pooling.  A block owns 64 channels of one
// segment: lanes = channel pairs, the four warps take every fourth row (four rows in flight each) and meet
// in shared memory.
template <bool BF16>
__global__ void __launch_bounds__(128) segment_max_affine_kernel(const void *__restrict__ xin, const int64_t *__restrict__ ptr,
                                                                 int c, const float *__restrict__ scale,
                                                                 const float *__restrict__ shift, float *__restrict__ out) {
    __shared__ float2 part[4][32];
    const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ch = 2 * (blockIdx.y * 32 + lane);
    const bool live = ch < c;
    const int64_t r0 = ptr[b], r1 = ptr[b + 1];
    const float s0 = live && scale ? scale[ch] : 1.f, s1 = live && scale ? scale[ch + 1] : 1.f;
    const float t0 = live && shift ? shift[ch] : 0.f, t1 = live && shift ? shift[ch + 1] : 0.f;
    float m0 = __int_as_float(0xff800000), m1 = m0;
    auto load = [&](int64_t r, float &a, float &bb) {
        if (BF16) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(
                static_cast<const __nv_bfloat16 *>(xin) + r * c + ch));
            a = f.x; bb = f.y;
        } else {
            const float2 f = *reinterpret_cast<const float2 *>(static_cast<const float *>(xin) + r * c + ch);
            a = f.x; bb = f.y;
        }
    };
    if (live) {
        int64_t r = r0 + w;
        for (; r + 12 < r1; r += 16) {
            float a[4], bb[4];
#pragma unroll
            for (int u = 0; u < 4; u++) load(r + 4 * u, a[u], bb[u]);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                m0 = fmaxf(m0, __fadd_rn(__fmul_rn(a[u], s0), t0));      // mul then add, as the tensor expression h * s + t
                m1 = fmaxf(m1, __fadd_rn(__fmul_rn(bb[u], s1), t1));
            }
        }
        for (; r < r1; r += 4) {
            float a, bb;
            load(r, a, bb);
            m0 = fmaxf(m0, __fadd_rn(__fmul_rn(a, s0), t0));
            m1 = fmaxf(m1, __fadd_rn(__fmul_rn(bb, s1), t1));
        }
    }
    part[w][lane] = make_float2(m0, m1);
    __syncthreads();
    if (w == 0 && live) {
#pragma unroll
        for (int k = 1; k < 4; k++) {
            m0 = fmaxf(m0, part[k][lane].x);
            m1 = fmaxf(m1, part[k][lane].y);
        }
        const bool any = r1 > r0;
        *reinterpret_cast<float2 *>(out + static_cast<int64_t>(b) * c + ch) = make_float2(any ? m0 : 0.f, any ? m1 : 0.f);
    }
}

// ------------------------------------------------------------------ scatter max / min with arg
__device__ __forceinline__ void atomic_max_f(float *a, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int *>(a), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned *>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_f(float *a, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int *>(a), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned *>(a), __float_as_uint(v));
}

__global__ void scatter_init_kernel(float *out, int64_t *arg, int64_t total, int is_max, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    out[i] = __int_as_float(is_max ? 0xff800000 : 0x7f800000);
    if (arg) arg[i] = n;
}
__global__ void scatter_reduce_kernel(const float *__restrict__ src, const int64_t *__restrict__ index, int64_t n,
                                      int c, int is_max, float *__restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n * c) return;
    const int64_t r = i / c;
    const int ch = static_cast<int>(i % c);
    float *dst = out + index[r] * c + ch;
    if (is_max) atomic_max_f(dst, src[i]); else atomic_min_f(dst, src[i]);
}
__global__ void scatter_arg_kernel(const float *__restrict__ src, const int64_t *__restrict__ index, int64_t n, int c,
                                   const float *__restrict__ out, int64_t *__restrict__ arg) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n * c) return;
    const int64_t r = i / c;
    const int ch = static_cast<int>(i % c);
    const int64_t o = index[r] * c + ch;
    if (src[i] == out[o]) atomicMin(reinterpret_cast<long long *>(arg + o), static_cast<long long>(r));
}
__global__ void scatter_final_kernel(float *out, const int64_t *arg, int64_t total, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float v = out[i];
    if (isinf(v) && (arg ? arg[i] == n : true)) out[i] = 0.f;
}

}  // namespace
}  // namespace p2w

using namespace p2w;

int p2w_conv_tc_launch(const void *x, int x_bf16, const float *pos_src, const float *pos_tgt, const int32_t *nbr,
                       int64_t n_src, int64_t n_tgt, int32_t k, int32_t c_in, int32_t hidden, int32_t c_out,
                       const float *w1, const float *b1, const float *w2, const float *b2, const float *bn_scale,
                       const float *bn_shift, void *out, int out_bf16, void *ws, size_t ws_bytes, cudaStream_t st,
                       bool packed, const int64_t *tgt_index);
size_t p2w_conv_tc_ws_bytes(int32_t c_in, int32_t hidden, int32_t c_out);
// conv_tc_g2.cu: the two-gather-group variant, used for hidden <= 64 (SA1) unless P2W_CONV_G2=0
int p2w_conv_tc2_launch(const void *x, int x_bf16, const float *pos_src, const float *pos_tgt, const int32_t *nbr,
                        int64_t n_src, int64_t n_tgt, int32_t k, int32_t c_in, int32_t hidden, int32_t c_out,
                        const float *w1, const float *b1, const float *w2, const float *b2, const float *bn_scale,
                        const float *bn_shift, void *out, int out_bf16, void *ws, size_t ws_bytes, cudaStream_t st,
                        bool packed, const int64_t *tgt_index);
size_t p2w_conv_tc2_ws_bytes(int32_t c_in, int32_t hidden, int32_t c_out);
static bool use_g2(int hidden) {
    static const int v = [] { const char *e = getenv("P2W_CONV_G2"); return e ? atoi(e) : 1; }();
    return v != 0 && hidden <= 64;
}

extern "C" size_t p2w_pointnet_conv_ws_bytes(int32_t c_in, int32_t hidden, int32_t c_out, int32_t mode) {
    if (mode == P2W_CONV_BF16_TC) {
        const size_t a = p2w_conv_tc_ws_bytes(c_in, hidden, c_out), b = p2w_conv_tc2_ws_bytes(c_in, hidden, c_out);
        return a > b ? a : b;
    }
    return sizeof(float) * (static_cast<size_t>(c_in + 4) * hidden + static_cast<size_t>(hidden) * c_out) + 256;
}

static int conv_dispatch(const void *xv, int32_t x_dtype, const float *pos_src, const float *pos_tgt,
                         const int32_t *nbr, int64_t n_src, int64_t n_tgt, int32_t k, int32_t c_in, int32_t hidden,
                         int32_t c_out, const float *w1, const float *b1, const float *w2, const float *b2,
                         const float *bn_scale, const float *bn_shift, void *outv, int32_t out_dtype, int32_t mode,
                         void *ws, size_t ws_bytes, p2w_stream_t stream, bool packed, const int64_t *tgt_index = nullptr) {
    P2W_REQUIRE((x_dtype == P2W_F32 || x_dtype == P2W_BF16) && (out_dtype == P2W_F32 || out_dtype == P2W_BF16),
                "p2w_pointnet_conv_max: unknown dtype");
    P2W_REQUIRE(mode == P2W_CONV_BF16_TC || (x_dtype == P2W_F32 && out_dtype == P2W_F32),
                "p2w_pointnet_conv_max: the FP32 mode takes and returns FP32 rows");
    const float *x = static_cast<const float *>(xv);
    float *out = static_cast<float *>(outv);
    P2W_REQUIRE(k >= 1 && k <= 32, "p2w_pointnet_conv_max: k=%d outside [1,32]", k);
    P2W_REQUIRE(c_in >= 1 && hidden % 8 == 0 && c_out % 8 == 0 && hidden > 0 && c_out > 0,
                "p2w_pointnet_conv_max: hidden=%d and c_out=%d must be positive multiples of 8", hidden, c_out);
    P2W_REQUIRE(mode == P2W_CONV_FP32 || mode == P2W_CONV_BF16_TC, "p2w_pointnet_conv_max: unknown mode %d", mode);
    P2W_REQUIRE(ws_bytes >= p2w_pointnet_conv_ws_bytes(c_in, hidden, c_out, mode),
                "p2w_pointnet_conv_max: workspace too small");
    if (n_tgt == 0) return P2W_OK;
    cudaStream_t st = as_stream(stream);
    if (mode == P2W_CONV_BF16_TC && use_g2(hidden))
        return p2w_conv_tc2_launch(xv, x_dtype == P2W_BF16, pos_src, pos_tgt, nbr, n_src, n_tgt, k, c_in, hidden, c_out,
                                   w1, b1, w2, b2, bn_scale, bn_shift, outv, out_dtype == P2W_BF16, ws, ws_bytes, st,
                                   packed, tgt_index);
    if (mode == P2W_CONV_BF16_TC)
        return p2w_conv_tc_launch(xv, x_dtype == P2W_BF16, pos_src, pos_tgt, nbr, n_src, n_tgt, k, c_in, hidden, c_out,
                                  w1, b1, w2, b2, bn_scale, bn_shift, outv, out_dtype == P2W_BF16, ws, ws_bytes, st,
                                  packed, tgt_index);
    P2W_REQUIRE(tgt_index == nullptr, "p2w_pointnet_conv_max: tgt_index needs the tensor-core mode");
    const int K1 = c_in + 4;
    float *w1t = static_cast<float *>(ws);
    float *w2t = w1t + static_cast<size_t>(K1) * hidden;
    // 16-byte alignment of the second panel for the float4 loads
    w2t = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(w2t) + 15) & ~uintptr_t(15));
    if (!packed) {
        P2W_LAUNCH(transpose_kernel, (K1 * hidden + 255) / 256, 256, 0, st)(w1, hidden, K1, w1t);
        P2W_LAUNCH(transpose_kernel, (hidden * c_out + 255) / 256, 256, 0, st)(w2, c_out, hidden, w2t);
    }
    const size_t smem = sizeof(float) * 32 * ((K1 | 1) + (hidden | 1));
    P2W_REQUIRE(smem <= 200 * 1024, "p2w_pointnet_conv_max: layer too wide for the FP32 kernel");
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaFuncSetAttribute(conv_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        smem_set = smem;
    }
    int64_t grid = n_tgt < 148 * 16 ? n_tgt : 148 * 16;
    P2W_LAUNCH(conv_simt_kernel, (unsigned)grid, CONV_WARPS * 32, smem, st)(x, pos_src, pos_tgt, nbr, n_tgt, k, c_in, hidden, c_out, w1t, b1, w2t, b2, bn_scale, bn_shift, out);
    return check_launch("p2w_pointnet_conv_max");
}

extern "C" int p2w_pointnet_conv_max(const float *x, const float *pos_src, const float *pos_tgt, const int32_t *nbr,
                                     int64_t n_src, int64_t n_tgt, int32_t k, int32_t c_in, int32_t hidden,
                                     int32_t c_out, const float *w1, const float *b1, const float *w2, const float *b2,
                                     const float *bn_scale, const float *bn_shift, float *out, int32_t mode, void *ws,
                                     size_t ws_bytes, p2w_stream_t stream) {
    return conv_dispatch(x, P2W_F32, pos_src, pos_tgt, nbr, n_src, n_tgt, k, c_in, hidden, c_out, w1, b1, w2, b2,
                         bn_scale, bn_shift, out, P2W_F32, mode, ws, ws_bytes, stream, false);
}

extern "C" int p2w_pointnet_conv_max_ex(const void *x, int32_t x_dtype, const float *pos_src, const float *pos_tgt,
                                        const int32_t *nbr, int64_t n_src, int64_t n_tgt, int32_t k, int32_t c_in,
                                        int32_t hidden, int32_t c_out, const float *w1, const float *b1,
                                        const float *w2, const float *b2, const float *bn_scale,
                                        const float *bn_shift, void *out, int32_t out_dtype, int32_t mode, void *ws,
                                        size_t ws_bytes, int32_t flags, const int64_t *tgt_index,
                                        p2w_stream_t stream) {
    return conv_dispatch(x, x_dtype, pos_src, pos_tgt, nbr, n_src, n_tgt, k, c_in, hidden, c_out, w1, b1, w2, b2,
                         bn_scale, bn_shift, out, out_dtype, mode, ws, ws_bytes, stream,
                         (flags & P2W_CONV_WS_PACKED) != 0, tgt_index);
}

extern "C" int p2w_knn_interpolate_ex(const void *x, int32_t x_dtype, const float *pos_x, const float *pos_y,
                                      const int32_t *nbr, int64_t ny, int32_t k, int32_t c, int32_t ld_out, void *out,
                                      int32_t out_dtype, p2w_stream_t stream) {
    P2W_REQUIRE(k >= 1 && k <= 32, "p2w_knn_interpolate: k=%d outside [1,32]", k);
    P2W_REQUIRE(c >= 1 && ld_out >= c, "p2w_knn_interpolate: bad channel count / stride");
    P2W_REQUIRE((x_dtype == P2W_F32 || x_dtype == P2W_BF16) && (out_dtype == P2W_F32 || out_dtype == P2W_BF16),
                "p2w_knn_interpolate: unknown dtype");
    if (ny == 0) return P2W_OK;
    cudaStream_t st = as_stream(stream);
    const unsigned blocks = (unsigned)((ny * 32 + 255) / 256);
    const float *xf = static_cast<const float *>(x);
    const __nv_bfloat16 *xh = static_cast<const __nv_bfloat16 *>(x);
    float *of = static_cast<float *>(out);
    __nv_bfloat16 *oh = static_cast<__nv_bfloat16 *>(out);
    if (x_dtype == P2W_F32 && out_dtype == P2W_F32)
        P2W_LAUNCH((interp_kernel<float, float>), blocks, 256, 0, st)(xf, pos_x, pos_y, nbr, ny, k, c, ld_out, of);
    else if (x_dtype == P2W_F32)
        P2W_LAUNCH((interp_kernel<float, __nv_bfloat16>), blocks, 256, 0, st)(xf, pos_x, pos_y, nbr, ny, k, c, ld_out, oh);
    else if (out_dtype == P2W_F32)
        P2W_LAUNCH((interp_kernel<__nv_bfloat16, float>), blocks, 256, 0, st)(xh, pos_x, pos_y, nbr, ny, k, c, ld_out, of);
    else
        P2W_LAUNCH((interp_kernel<__nv_bfloat16, __nv_bfloat16>), blocks, 256, 0, st)(xh, pos_x, pos_y, nbr, ny, k, c, ld_out, oh);
    return check_launch("p2w_knn_interpolate");
}

extern "C" int p2w_knn_interpolate(const float *x, const float *pos_x, const float *pos_y, const int32_t *nbr,
                                   int64_t ny, int32_t k, int32_t c, int32_t ld_out, float *out,
                                   p2w_stream_t stream) {
    return p2w_knn_interpolate_ex(x, P2W_F32, pos_x, pos_y, nbr, ny, k, c, ld_out, out, P2W_F32, stream);
}

extern "C" int p2w_knn_interpolate_cat(const void *x, int32_t x_dtype, const float *pos_x, const float *pos_y,
                                       const int32_t *nbr, int64_t ny, int32_t k, int32_t c, const void *skip,
                                       int32_t skip_dtype, int32_t c_skip, int32_t ld_out, void *out,
                                       int32_t out_dtype, p2w_stream_t stream) {
    P2W_REQUIRE(k >= 1 && k <= 32, "p2w_knn_interpolate_cat: k=%d outside [1,32]", k);
    P2W_REQUIRE(c >= 8 && c % 8 == 0 && c_skip >= 0 && c_skip % 8 == 0 && ld_out >= c + c_skip && ld_out % 8 == 0,
                "p2w_knn_interpolate_cat: channel counts and the row stride must be multiples of 8");
    P2W_REQUIRE((c_skip == 0) || skip != nullptr, "p2w_knn_interpolate_cat: skip rows missing");
    auto okdt = [](int d) { return d == P2W_F32 || d == P2W_BF16; };
    P2W_REQUIRE(okdt(x_dtype) && okdt(skip_dtype) && okdt(out_dtype), "p2w_knn_interpolate_cat: unknown dtype");
    P2W_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(skip) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0,
                "p2w_knn_interpolate_cat: rows must be 16-byte aligned");
    if (ny == 0) return P2W_OK;
    cudaStream_t st = as_stream(stream);
    const unsigned blocks = (unsigned)((ny * 32 + 255) / 256);
    typedef __nv_bfloat16 bf;
#define P2W_IC(TI, TS, TO)                                                                                       \
    P2W_LAUNCH((interp_cat_kernel<TI, TS, TO>), blocks, 256, 0, st)(static_cast<const TI *>(x), pos_x, pos_y, nbr, ny, k, c, \
                                                                    static_cast<const TS *>(skip), c_skip, ld_out,   \
                                                                    static_cast<TO *>(out))
    const int sel = x_dtype * 4 + skip_dtype * 2 + out_dtype;
    switch (sel) {
        case 0: P2W_IC(float, float, float); break;
        case 1: P2W_IC(float, float, bf); break;
        case 2: P2W_IC(float, bf, float); break;
        case 3: P2W_IC(float, bf, bf); break;
        case 4: P2W_IC(bf, float, float); break;
        case 5: P2W_IC(bf, float, bf); break;
        case 6: P2W_IC(bf, bf, float); break;
        default: P2W_IC(bf, bf, bf); break;
    }
#undef P2W_IC
    return check_launch("p2w_knn_interpolate_cat");
}

extern "C" int p2w_knn_interpolate_add(const void *y, int32_t y_dtype, const float *pos_x, const float *pos_y,
                                       const int32_t *nbr, int64_t ny, int32_t k, int32_t c, const void *z, void *out,
                                       int32_t out_dtype, int32_t relu, p2w_stream_t stream) {
    P2W_REQUIRE(k >= 1 && k <= 32, "p2w_knn_interpolate_add: k=%d outside [1,32]", k);
    P2W_REQUIRE(c >= 8 && c % 8 == 0, "p2w_knn_interpolate_add: the channel count must be a multiple of 8");
    auto okdt = [](int d) { return d == P2W_F32 || d == P2W_BF16; };
    P2W_REQUIRE(okdt(y_dtype) && okdt(out_dtype), "p2w_knn_interpolate_add: unknown dtype");
    P2W_REQUIRE(((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0,
                "p2w_knn_interpolate_add: rows must be 16-byte aligned");
    if (ny == 0) return P2W_OK;
    cudaStream_t st = as_stream(stream);
    const unsigned blocks = (unsigned)((ny * 32 + 255) / 256);
    typedef __nv_bfloat16 bf;
#define P2W_IA(TI, TO)                                                                                                  \
    P2W_LAUNCH((interp_add_kernel<TI, TO>), blocks, 256, 0, st)(static_cast<const TI *>(y), pos_x, pos_y, nbr, ny, k, c, \
                                                                static_cast<const TO *>(z), relu, static_cast<TO *>(out))
    if (y_dtype == P2W_F32 && out_dtype == P2W_F32) P2W_IA(float, float);
    else if (y_dtype == P2W_F32) P2W_IA(float, bf);
    else if (out_dtype == P2W_F32) P2W_IA(bf, float);
    else P2W_IA(bf, bf);
#undef P2W_IA
    return check_launch("p2w_knn_interpolate_add");
}

extern "C" int p2w_segment_max(const float *x, const int64_t *ptr, int32_t num_segments, int32_t c, float *out,
                               p2w_stream_t stream) {
    P2W_REQUIRE(num_segments >= 1 && c >= 1, "p2w_segment_max: bad sizes");
    dim3 grid(num_segments, (c + 127) / 128);
    P2W_LAUNCH(segment_max_kernel, grid, 128, 0, as_stream(stream))(x, ptr, c, out);
    return check_launch("p2w_segment_max");
}

extern "C" int p2w_segment_max_ex(const void *x, int32_t dtype, const int64_t *ptr, int32_t num_segments, int32_t c,
                                  const float *scale, const float *shift, float *out, p2w_stream_t stream) {
    P2W_REQUIRE(num_segments >= 1 && c >= 2 && c % 2 == 0, "p2w_segment_max_ex: c=%d must be even, segments >= 1", c);
    P2W_REQUIRE(dtype == P2W_F32 || dtype == P2W_BF16, "p2w_segment_max_ex: unknown dtype %d", dtype);
    P2W_REQUIRE((reinterpret_cast<uintptr_t>(x) & 7u) == 0 && (reinterpret_cast<uintptr_t>(out) & 7u) == 0,
                "p2w_segment_max_ex: rows must be 8-byte aligned");
    dim3 grid(num_segments, (c / 2 + 31) / 32);
    if (dtype == P2W_BF16)
        P2W_LAUNCH(segment_max_affine_kernel<true>, grid, 128, 0, as_stream(stream))(x, ptr, c, scale, shift, out);
    else
        P2W_LAUNCH(segment_max_affine_kernel<false>, grid, 128, 0, as_stream(stream))(x, ptr, c, scale, shift, out);
    return check_launch("p2w_segment_max_ex");
}

extern "C" int p2w_scatter_minmax(const float *src, const int64_t *index, int64_t n, int32_t c, int64_t dim_size,
                                  int32_t is_max, float *out, int64_t *arg, p2w_stream_t stream) {
    P2W_REQUIRE(c >= 1 && dim_size >= 0 && n >= 0, "p2w_scatter_minmax: bad sizes");
    cudaStream_t st = as_stream(stream);
    const int64_t total = dim_size * c;
    if (total == 0) return P2W_OK;
    P2W_LAUNCH(scatter_init_kernel, (unsigned)((total + 255) / 256), 256, 0, st)(out, arg, total, is_max, n);
    if (n > 0) {
        const unsigned blocks = (unsigned)((n * c + 255) / 256);
        P2W_LAUNCH(scatter_reduce_kernel, blocks, 256, 0, st)(src, index, n, c, is_max, out);
        if (arg) P2W_LAUNCH(scatter_arg_kernel, blocks, 256, 0, st)(src, index, n, c, out, arg);
    }
    P2W_LAUNCH(scatter_final_kernel, (unsigned)((total + 255) / 256), 256, 0, st)(out, arg, total, n);
    return check_launch("p2w_scatter_minmax");
}
