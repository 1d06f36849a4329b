// misc.cu -- SA glue, K7 batch packing and K8 per-point write-back: streaming, coalesced,
// HBM-bound kernels around the hot path (src/model.py:109,122,124; src/predicter.py:78-94,
// 199-214).
#include "common.cuh"

namespace p2w {
namespace {

constexpr unsigned FULL = 0xffffffffu;

// pos4[i] = (pos[i]/sf[b], refl[i]);  pos_back[i] = (pos[i]/sf[b])*sf[b]
__global__ void __launch_bounds__(256) sa_prepare_kernel(const float *__restrict__ pos, int ld,
                                                         const float *__restrict__ refl,
                                                         const int64_t *__restrict__ ptr, const float *__restrict__ sf,
                                                         int B, int64_t n, float *__restrict__ pos4,
                                                         float *__restrict__ pos_back) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float s = sf[find_tile(ptr, B, i)];
    float4 o;
    o.x = __fdiv_rn(pos[i * ld + 0], s);
    o.y = __fdiv_rn(pos[i * ld + 1], s);
    o.z = __fdiv_rn(pos[i * ld + 2], s);
    o.w = refl[i];
    reinterpret_cast<float4 *>(pos4)[i] = o;
    if (pos_back) {
        pos_back[i * 3 + 0] = __fmul_rn(o.x, s);
        pos_back[i * 3 + 1] = __fmul_rn(o.y, s);
        pos_back[i * 3 + 2] = __fmul_rn(o.z, s);
    }
}

__device__ __forceinline__ double block_sum(double v, double *sh) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double w = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0.0;
        for (int o = 16; o; o >>= 1) w += __shfl_xor_sync(FULL, w, o);
        if (lane == 0) sh[0] = w;
    }
    __syncthreads();
    const double r = sh[0];
    __syncthreads();
    return r;
}

// One CTA per tile.  local_shift = mean(pos) with FP64 accumulation in a fixed tree (exact for
// tile-sized sums of FP32 coordinates, hence order independent), pos -= shift,
// sf = max sqrt((x*x + y*y) + z*z) in FP32.
__global__ void __launch_bounds__(1024) pack_kernel(const float *__restrict__ cloud, int ld,
                                                    const int64_t *__restrict__ index,
                                                    const int64_t *__restrict__ ptr, float *__restrict__ pos,
                                                    float *__restrict__ refl, int64_t *__restrict__ batch,
                                                    float *__restrict__ local_shift, float *__restrict__ sf) {
    __shared__ double sh[32];
    __shared__ float shf[32];
    const int b = blockIdx.x;
    const int64_t r0 = ptr[b], r1 = ptr[b + 1];
    double sx = 0, sy = 0, sz = 0;
    for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
        const int64_t g = index ? index[r] : r;
        sx += static_cast<double>(cloud[g * ld + 0]);
        sy += static_cast<double>(cloud[g * ld + 1]);
        sz += static_cast<double>(cloud[g * ld + 2]);
    }
    sx = block_sum(sx, sh);
    sy = block_sum(sy, sh);
    sz = block_sum(sz, sh);
    const double cnt = static_cast<double>(r1 - r0);
    const float mx = (r1 > r0) ? static_cast<float>(sx / cnt) : 0.f;
    const float my = (r1 > r0) ? static_cast<float>(sy / cnt) : 0.f;
    const float mz = (r1 > r0) ? static_cast<float>(sz / cnt) : 0.f;
    float best = 0.f;
    for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
        const int64_t g = index ? index[r] : r;
        const float px = __fsub_rn(cloud[g * ld + 0], mx);
        const float py = __fsub_rn(cloud[g * ld + 1], my);
        const float pz = __fsub_rn(cloud[g * ld + 2], mz);
        pos[r * 3 + 0] = px;
        pos[r * 3 + 1] = py;
        pos[r * 3 + 2] = pz;
        refl[r] = cloud[g * ld + 3];
        batch[r] = b;
        const float nrm =
            __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
        best = fmaxf(best, nrm);
    }
    for (int o = 16; o; o >>= 1) best = fmaxf(best, __shfl_xor_sync(FULL, best, o));
    if ((threadIdx.x & 31) == 0) shf[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        float w = threadIdx.x < (blockDim.x >> 5) ? shf[threadIdx.x] : 0.f;
        for (int o = 16; o; o >>= 1) w = fmaxf(w, __shfl_xor_sync(FULL, w, o));
        if (threadIdx.x == 0) {
            sf[b] = w;
            local_shift[b * 3 + 0] = mx;
            local_shift[b * 3 + 1] = my;
            local_shift[b * 3 + 2] = mz;
        }
    }
}

__global__ void __launch_bounds__(256) writeback_kernel(const float *__restrict__ logits,
                                                        const float *__restrict__ pos,
                                                        const int64_t *__restrict__ ptr,
                                                        const float *__restrict__ local_shift, int B, int64_t m,
                                                        float is_wood, double *__restrict__ out64,
                                                        float *__restrict__ prob, uint8_t *__restrict__ pred,
                                                        float *__restrict__ xyz32) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= m) return;
    float z = logits[i];
    if (isnan(z)) z = 0.f;                       // torch.nan_to_num
    else if (isinf(z)) z = z > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
    const float p = 1.0f / (1.0f + expf(-z));
    const int lab = p >= is_wood ? 1 : 0;
    if (prob) prob[i] = p;
    if (pred) pred[i] = static_cast<uint8_t>(lab);
    if (out64 || xyz32) {
        const int b = find_tile(ptr, B, i);
        const double x = static_cast<double>(pos[i * 3 + 0]) + static_cast<double>(local_shift[b * 3 + 0]);
        const double y = static_cast<double>(pos[i * 3 + 1]) + static_cast<double>(local_shift[b * 3 + 1]);
        const double z = static_cast<double>(pos[i * 3 + 2]) + static_cast<double>(local_shift[b * 3 + 2]);
        if (out64) {
            out64[i * 5 + 0] = x;
            out64[i * 5 + 1] = y;
            out64[i * 5 + 2] = z;
            out64[i * 5 + 3] = static_cast<double>(lab);
            out64[i * 5 + 4] = static_cast<double>(p);
        }
        if (xyz32) {   // the same un-shifted coordinates rounded to FP32: the vote's search runs in FP32
            xyz32[i * 3 + 0] = static_cast<float>(x);
            xyz32[i * 3 + 1] = static_cast<float>(y);
            xyz32[i * 3 + 2] = static_cast<float>(z);
        }
    }
}

}  // namespace
}  // namespace p2w

using namespace p2w;

extern "C" int p2w_sa_prepare(const float *pos, int32_t ld_pos, const float *refl, const int64_t *ptr, const float *sf,
                              int32_t num_tiles, int64_t n, float *pos4, float *pos_back, p2w_stream_t stream) {
    P2W_REQUIRE(ld_pos >= 3 && num_tiles >= 1, "p2w_sa_prepare: bad sizes");
    P2W_REQUIRE((reinterpret_cast<uintptr_t>(pos4) & 15u) == 0, "p2w_sa_prepare: pos4 must be 16-byte aligned");
    if (n == 0) return P2W_OK;
    P2W_LAUNCH(sa_prepare_kernel, (unsigned)((n + 255) / 256), 256, 0, as_stream(stream))(pos, ld_pos, refl, ptr, sf, num_tiles, n, pos4, pos_back);
    return check_launch("p2w_sa_prepare");
}

extern "C" int p2w_pack(const float *cloud, int32_t ld, const int64_t *index, const int64_t *ptr, int32_t num_tiles,
                        int64_t m, float *pos, float *refl, int64_t *batch, float *local_shift, float *sf,
                        p2w_stream_t stream) {
    P2W_REQUIRE(ld >= 4 && num_tiles >= 1, "p2w_pack: cloud needs x,y,z,reflectance columns");
    (void)m;
    P2W_LAUNCH(pack_kernel, num_tiles, 1024, 0, as_stream(stream))(cloud, ld, index, ptr, pos, refl, batch, local_shift, sf);
    return check_launch("p2w_pack");
}

extern "C" int p2w_writeback(const float *logits, const float *pos, const int64_t *ptr, const float *local_shift,
                             int32_t num_tiles, int64_t m, float is_wood, double *out64, float *prob, uint8_t *pred,
                             float *xyz32, p2w_stream_t stream) {
    P2W_REQUIRE(num_tiles >= 1, "p2w_writeback: bad sizes");
    if (m == 0) return P2W_OK;
    P2W_LAUNCH(writeback_kernel, (unsigned)((m + 255) / 256), 256, 0, as_stream(stream))(logits, pos, ptr, local_shift, num_tiles, m, is_wood, out64, prob, pred, xyz32);
    return check_launch("p2w_writeback");
}
