// tiling.cu -- K6: the vectorised front end of Voxelise (src/preprocessing.py:18-64,116-120):
// height normalisation, quantile normalisation of reflectance, the 5-column feature array the
// reference voxelises, and the priority keys that stand in for torch.multinomial.  The grouping
// itself is p2w_grid + p2w_sort_pairs + p2w_unique_last (voxel.cu): one sort per grid size
// replaces the reference's O(V*N) mask-per-voxel loop.
#include "common.cuh"

namespace p2w {
namespace {

constexpr unsigned FULL = 0xffffffffu;
const float kPosInf = __builtin_huge_valf();
const float kNegInf = -__builtin_huge_valf();

__device__ __forceinline__ void atomic_min_f(float *a, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int *>(a), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned *>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float *a, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int *>(a), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned *>(a), __float_as_uint(v));
}

// torch.bucketize(v, edges) with edges_i = lo + cell*i: number of edges strictly below v
__device__ __forceinline__ int bucket(float v, float lo, float cell, int nb) {
    int i = static_cast<int>(floorf((v - lo) / cell));
    if (i < 0) i = 0;
    if (i > nb) i = nb;
    while (i < nb && __fadd_rn(lo, __fmul_rn(cell, static_cast<float>(i))) < v) i++;
    while (i > 0 && !(__fadd_rn(lo, __fmul_rn(cell, static_cast<float>(i - 1))) < v)) i--;
    return i;
}

__global__ void fill_kernel(float *a, int64_t n, float v) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

__global__ void __launch_bounds__(256) ground_min_kernel(const float *__restrict__ cloud, int ld, int64_t n,
                                                         const float *__restrict__ mn_xy, float cell, int nbx,
                                                         int nby, float *__restrict__ cell_min) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = bucket(cloud[i * ld + 0], mn_xy[0], cell, nbx);
    const int cy = bucket(cloud[i * ld + 1], mn_xy[1], cell, nby);
    atomic_min_f(&cell_min[cx * (nby + 1) + cy], cloud[i * ld + 2]);
}

__global__ void __launch_bounds__(256) ground_apply_kernel(const float *__restrict__ cloud, int ld, int64_t n,
                                                           const float *__restrict__ mn_xy, float cell, int nbx,
                                                           int nby, const float *__restrict__ cell_min,
                                                           float *__restrict__ n_z) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = bucket(cloud[i * ld + 0], mn_xy[0], cell, nbx);
    const int cy = bucket(cloud[i * ld + 1], mn_xy[1], cell, nby);
    n_z[i] = __fsub_rn(cloud[i * ld + 2], cell_min[cx * (nby + 1) + cy]);
}

// order-preserving map float -> uint32
__device__ __forceinline__ uint32_t sortable(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(256) refl_keys_kernel(const float *__restrict__ cloud, int ld, int col, int64_t n,
                                                        uint64_t *__restrict__ keys) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = sortable(cloud[i * ld + col]);
}

__global__ void __launch_bounds__(256) refl_normal_kernel(const int32_t *__restrict__ sorted_idx, int64_t n,
                                                          float *__restrict__ v, float *__restrict__ mnmx) {
    const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    float val = 0.f;
    const bool ok = p < n;
    if (ok) {
        float q = __fdiv_rn(__fadd_rn(static_cast<float>(p), 1.0f), static_cast<float>(n + 1));
        q = fminf(fmaxf(q, 1e-7f), 1.0f - 1e-7f);
        val = __fmul_rn(erfinvf(__fsub_rn(__fmul_rn(2.0f, q), 1.0f)), 1.41421354f);
        v[sorted_idx[p]] = val;
    }
    float lo = ok ? val : __int_as_float(0x7f800000), hi = ok ? val : __int_as_float(0xff800000);
    for (int o = 16; o; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(FULL, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(FULL, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { atomic_min_f(&mnmx[0], lo); atomic_max_f(&mnmx[1], hi); }
}

__global__ void __launch_bounds__(256) refl_scale_kernel(const float *__restrict__ v, int64_t n,
                                                         const float *__restrict__ mnmx, float *__restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float mn = mnmx[0], mx = mnmx[1];
    out[i] = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fsub_rn(v[i], mn)), __fsub_rn(mx, mn)), 1.0f);
}

__global__ void __launch_bounds__(256) assemble5_kernel(const float *__restrict__ cloud, int ld,
                                                        const float *__restrict__ refl, const float *__restrict__ n_z,
                                                        int64_t n, float *__restrict__ feat) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    feat[i * 5 + 0] = cloud[i * ld + 0];
    feat[i * 5 + 1] = cloud[i * ld + 1];
    feat[i * 5 + 2] = cloud[i * ld + 2];
    feat[i * 5 + 3] = refl ? refl[i] : cloud[i * ld + 3];
    feat[i * 5 + 4] = n_z[i];
}

__device__ __forceinline__ uint32_t mix32(uint32_t h) {   // murmur3 finaliser
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

__global__ void __launch_bounds__(256) priority_keys_kernel(const float *__restrict__ feat,
                                                            const int32_t *__restrict__ members,
                                                            const int32_t *__restrict__ member_tile, int64_t m,
                                                            float refl_min, uint32_t seed,
                                                            uint64_t *__restrict__ keys) {
    const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const int32_t i = members[r];
    const float w = __fadd_rn(__fsub_rn(feat[static_cast<int64_t>(i) * 5 + 3], refl_min), 1e-8f);
    const uint32_t h = mix32(static_cast<uint32_t>(i) * 0x9e3779b9u + seed);
    const float u = __fmul_rn(static_cast<float>((h >> 8) + 1u), 5.9604644775390625e-08f);   // (0,1], exact
    const float pr = __fdiv_rn(w, u);
    keys[r] = (static_cast<uint64_t>(static_cast<uint32_t>(member_tile[r])) << 32) |
              static_cast<uint64_t>(~sortable(pr));
}

}  // namespace
}  // namespace p2w

using namespace p2w;


extern "C" int p2w_ground_normalize(const float *cloud, int32_t ld, int64_t n, const float *mn_xy, float cell,
                                    int32_t nbx, int32_t nby, float *cell_min, float *n_z, p2w_stream_t stream) {
    P2W_REQUIRE(ld >= 3 && nbx >= 1 && nby >= 1 && cell > 0.f, "p2w_ground_normalize: bad arguments");
    cudaStream_t st = as_stream(stream);
    if (n == 0) return P2W_OK;
    const int64_t cells = static_cast<int64_t>(nbx + 1) * (nby + 1);
    P2W_LAUNCH(fill_kernel, (unsigned)(((cells) + 255) / 256), 256, 0, st)(cell_min, cells, kPosInf);
    P2W_LAUNCH(ground_min_kernel, (unsigned)(((n) + 255) / 256), 256, 0, st)(cloud, ld, n, mn_xy, cell, nbx, nby, cell_min);
    P2W_LAUNCH(ground_apply_kernel, (unsigned)(((n) + 255) / 256), 256, 0, st)(cloud, ld, n, mn_xy, cell, nbx, nby, cell_min, n_z);
    return check_launch("p2w_ground_normalize");
}

extern "C" int p2w_reflectance_keys(const float *cloud, int32_t ld, int32_t col, int64_t n, uint64_t *keys,
                                    p2w_stream_t stream) {
    cudaStream_t st = as_stream(stream);
    if (n == 0) return P2W_OK;
    P2W_LAUNCH(refl_keys_kernel, (unsigned)(((n) + 255) / 256), 256, 0, st)(cloud, ld, col, n, keys);
    return check_launch("p2w_reflectance_keys");
}

extern "C" int p2w_reflectance_normalize(const int32_t *sorted_idx, int64_t n, float *v, float *mnmx_ws, float *out,
                                         p2w_stream_t stream) {
    cudaStream_t st = as_stream(stream);
    if (n == 0) return P2W_OK;
    P2W_LAUNCH(fill_kernel, 1, 32, 0, st)(mnmx_ws, 1, kPosInf);
    P2W_LAUNCH(fill_kernel, 1, 32, 0, st)(mnmx_ws + 1, 1, kNegInf);
    P2W_LAUNCH(refl_normal_kernel, (unsigned)(((n) + 255) / 256), 256, 0, st)(sorted_idx, n, v, mnmx_ws);
    P2W_LAUNCH(refl_scale_kernel, (unsigned)(((n) + 255) / 256), 256, 0, st)(v, n, mnmx_ws, out);
    return check_launch("p2w_reflectance_normalize");
}

extern "C" int p2w_assemble5(const float *cloud, int32_t ld, const float *refl, const float *n_z, int64_t n,
                             float *feat, p2w_stream_t stream) {
    cudaStream_t st = as_stream(stream);
    if (n == 0) return P2W_OK;
    P2W_LAUNCH(assemble5_kernel, (unsigned)(((n) + 255) / 256), 256, 0, st)(cloud, ld, refl, n_z, n, feat);
    return check_launch("p2w_assemble5");
}

extern "C" int p2w_priority_keys(const float *feat, const int32_t *members, const int32_t *member_tile, int64_t m,
                                 float refl_min, uint32_t seed, uint64_t *keys, p2w_stream_t stream) {
    cudaStream_t st = as_stream(stream);
    if (m == 0) return P2W_OK;
    P2W_LAUNCH(priority_keys_kernel, (unsigned)(((m) + 255) / 256), 256, 0, st)(feat, members, member_tile, m, refl_min, seed, keys);
    return check_launch("p2w_priority_keys");
}
