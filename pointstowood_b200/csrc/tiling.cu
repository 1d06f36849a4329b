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

// Few cells (a 1 M-point plot has 25, a 100 M-point plot 1 681): one million atomics on a handful of global words
// serialise in the L2 (259 us per 1 M points).  A CTA reduces its points into a copy of the cell table in shared
// memory first and touches every global cell at most once.
__global__ void __launch_bounds__(256) ground_min_smem_kernel(const float *__restrict__ cloud, int ld, int64_t n,
                                                              const float *__restrict__ mn_xy, float cell, int nbx,
                                                              int nby, int cells, float *__restrict__ cell_min) {
    extern __shared__ float s_min[];
    for (int c = threadIdx.x; c < cells; c += blockDim.x) s_min[c] = __int_as_float(0x7f800000);
    __syncthreads();
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int cx = bucket(cloud[i * ld + 0], mn_xy[0], cell, nbx);
        const int cy = bucket(cloud[i * ld + 1], mn_xy[1], cell, nby);
        atomic_min_f(&s_min[cx * (nby + 1) + cy], cloud[i * ld + 2]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cells; c += blockDim.x) {
        const float v = s_min[c];
        if (v < __int_as_float(0x7f800000)) atomic_min_f(&cell_min[c], v);
    }
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

// sorted position p of this call holds global rank rank0 + p of n_total values (a sharded plot ranks a key range per GPU)
__global__ void __launch_bounds__(256) refl_normal_kernel(const int32_t *__restrict__ sorted_idx, int64_t n,
                                                          int64_t rank0, int64_t n_total, float *__restrict__ v,
                                                          float *__restrict__ mnmx) {
    const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    float val = 0.f;
    const bool ok = p < n;
    if (ok) {
        float q = __fdiv_rn(__fadd_rn(static_cast<float>(rank0 + p), 1.0f), static_cast<float>(n_total + 1));
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
__device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
    z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ull; z ^= z >> 27; z *= 0x94d049bb133111ebull; z ^= z >> 31;
    return z;
}

// Efraimidis-Spirakis: ascending -log(u_i)/w_i lists the members in the order sequential weighted sampling
// WITHOUT replacement draws them (the law of torch.multinomial, src/preprocessing.py:118).  u is a
// counter-based hash of (seed, global point index) in (0,1]; the key is the raw bit pattern of the
// non-negative float64, which sorts like the value.
__global__ void __launch_bounds__(256) sampling_keys_kernel(const float *__restrict__ feat, int ld, int col,
                                                            const int32_t *__restrict__ members,
                                                            const int32_t *__restrict__ global_index, int64_t m,
                                                            float refl_min, uint32_t seed,
                                                            uint64_t *__restrict__ keys) {
    const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const int32_t i = members[r];
    const uint32_t g = static_cast<uint32_t>(global_index ? global_index[i] : i);
    const float w = __fadd_rn(__fsub_rn(feat[static_cast<int64_t>(i) * ld + col], refl_min), 1e-8f);   // :99,104
    const uint32_t h = mix32(g * 0x9e3779b9u + seed);
    const double u = static_cast<double>((h >> 8) + 1u) * 5.9604644775390625e-08;     // (0,1], exact
    const double key = -log(u) / static_cast<double>(w);
    keys[r] = static_cast<uint64_t>(__double_as_longlong(key > 0.0 ? key : 0.0));
}

// torch.randint(0, n_t, (max_pts,)) of src/preprocessing.py:120: max_pts draws WITH replacement per
// oversized tile; draw s of the tile with voxel id v takes member hash(seed, v, s) mod n_t.
__global__ void __launch_bounds__(256) replacement_picks_kernel(const int32_t *__restrict__ members,
                                                                const int64_t *__restrict__ seg,
                                                                const int64_t *__restrict__ voxel_id, int num_tiles,
                                                                int max_pts, uint64_t seed,
                                                                int32_t *__restrict__ picks) {
    const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= static_cast<int64_t>(num_tiles) * max_pts) return;
    const int t = static_cast<int>(r / max_pts);
    const uint64_t s = static_cast<uint64_t>(r - static_cast<int64_t>(t) * max_pts);
    const int64_t lo = seg[t], n_t = seg[t + 1] - lo;
    const uint64_t h = mix64(mix64(seed ^ (static_cast<uint64_t>(voxel_id[t]) * 0x9e3779b97f4a7c15ull)) + s);
    picks[r] = members[lo + static_cast<int64_t>(h % static_cast<uint64_t>(n_t))];
}

}  // namespace
}  // namespace p2w

using namespace p2w;


extern "C" int p2w_ground_min(const float *cloud, int32_t ld, int64_t n, const float *mn_xy, float cell,
                              int32_t nbx, int32_t nby, float *cell_min, p2w_stream_t stream) {
    P2W_REQUIRE(ld >= 3 && nbx >= 1 && nby >= 1 && cell > 0.f, "p2w_ground_min: bad arguments");
    cudaStream_t st = as_stream(stream);
    const int64_t cells = static_cast<int64_t>(nbx + 1) * (nby + 1);
    P2W_LAUNCH(fill_kernel, (unsigned)(((cells) + 255) / 256), 256, 0, st)(cell_min, cells, kPosInf);
    if (n && cells <= 8192) {
        int64_t blocks = (n + 256 * 16 - 1) / (256 * 16);                  // >= 16 points per thread before the flush
        if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        P2W_LAUNCH(ground_min_smem_kernel, (unsigned)blocks, 256, static_cast<size_t>(cells) * sizeof(float), st)(
            cloud, ld, n, mn_xy, cell, nbx, nby, static_cast<int>(cells), cell_min);
    } else if (n) {
        P2W_LAUNCH(ground_min_kernel, (unsigned)(((n) + 255) / 256), 256, 0, st)(cloud, ld, n, mn_xy, cell, nbx, nby, cell_min);
    }
    return check_launch("p2w_ground_min");
}

extern "C" int p2w_ground_apply(const float *cloud, int32_t ld, int64_t n, const float *mn_xy, float cell,
                                int32_t nbx, int32_t nby, const float *cell_min, float *n_z, p2w_stream_t stream) {
    P2W_REQUIRE(ld >= 3 && nbx >= 1 && nby >= 1 && cell > 0.f, "p2w_ground_apply: bad arguments");
    if (n == 0) return P2W_OK;
    P2W_LAUNCH(ground_apply_kernel, (unsigned)(((n) + 255) / 256), 256, 0, as_stream(stream))(cloud, ld, n, mn_xy, cell, nbx, nby, cell_min, n_z);
    return check_launch("p2w_ground_apply");
}

extern "C" int p2w_ground_normalize(const float *cloud, int32_t ld, int64_t n, const float *mn_xy, float cell,
                                    int32_t nbx, int32_t nby, float *cell_min, float *n_z, p2w_stream_t stream) {
    if (n == 0) return P2W_OK;
    const int rc = p2w_ground_min(cloud, ld, n, mn_xy, cell, nbx, nby, cell_min, stream);
    return rc ? rc : p2w_ground_apply(cloud, ld, n, mn_xy, cell, nbx, nby, cell_min, n_z, stream);
}

extern "C" int p2w_reflectance_keys(const float *cloud, int32_t ld, int32_t col, int64_t n, uint64_t *keys,
                                    p2w_stream_t stream) {
    cudaStream_t st = as_stream(stream);
    if (n == 0) return P2W_OK;
    P2W_LAUNCH(refl_keys_kernel, (unsigned)(((n) + 255) / 256), 256, 0, st)(cloud, ld, col, n, keys);
    return check_launch("p2w_reflectance_keys");
}

extern "C" int p2w_reflectance_values(const int32_t *sorted_idx, int64_t n, int64_t rank0, int64_t n_total, float *v,
                                      float *mnmx, p2w_stream_t stream) {
    P2W_REQUIRE(rank0 >= 0 && rank0 + n <= n_total, "p2w_reflectance_values: ranks outside [0, n_total)");
    cudaStream_t st = as_stream(stream);
    P2W_LAUNCH(fill_kernel, 1, 32, 0, st)(mnmx, 1, kPosInf);
    P2W_LAUNCH(fill_kernel, 1, 32, 0, st)(mnmx + 1, 1, kNegInf);
    if (n) P2W_LAUNCH(refl_normal_kernel, (unsigned)(((n) + 255) / 256), 256, 0, st)(sorted_idx, n, rank0, n_total, v, mnmx);
    return check_launch("p2w_reflectance_values");
}

extern "C" int p2w_reflectance_scale(const float *v, int64_t n, const float *mnmx, float *out, p2w_stream_t stream) {
    if (n == 0) return P2W_OK;
    P2W_LAUNCH(refl_scale_kernel, (unsigned)(((n) + 255) / 256), 256, 0, as_stream(stream))(v, n, mnmx, out);
    return check_launch("p2w_reflectance_scale");
}

extern "C" int p2w_reflectance_normalize(const int32_t *sorted_idx, int64_t n, float *v, float *mnmx_ws, float *out,
                                         p2w_stream_t stream) {
    if (n == 0) return P2W_OK;
    const int rc = p2w_reflectance_values(sorted_idx, n, 0, n, v, mnmx_ws, stream);
    return rc ? rc : p2w_reflectance_scale(v, n, mnmx_ws, out, stream);
}

extern "C" int p2w_assemble5(const float *cloud, int32_t ld, const float *refl, const float *n_z, int64_t n,
                             float *feat, p2w_stream_t stream) {
    cudaStream_t st = as_stream(stream);
    if (n == 0) return P2W_OK;
    P2W_LAUNCH(assemble5_kernel, (unsigned)(((n) + 255) / 256), 256, 0, st)(cloud, ld, refl, n_z, n, feat);
    return check_launch("p2w_assemble5");
}

extern "C" int p2w_sampling_keys(const float *feat, int32_t ld, int32_t col, const int32_t *members,
                                 const int32_t *global_index, int64_t m, float refl_min, uint32_t seed, uint64_t *keys,
                                 p2w_stream_t stream) {
    P2W_REQUIRE(ld >= 1 && col >= 0 && col < ld, "p2w_sampling_keys: bad column");
    if (m == 0) return P2W_OK;
    P2W_LAUNCH(sampling_keys_kernel, (unsigned)(((m) + 255) / 256), 256, 0, as_stream(stream))(feat, ld, col, members, global_index, m, refl_min, seed, keys);
    return check_launch("p2w_sampling_keys");
}

extern "C" int p2w_replacement_picks(const int32_t *members, const int64_t *seg, const int64_t *voxel_id,
                                     int32_t num_tiles, int32_t max_pts, uint64_t seed, int32_t *picks,
                                     p2w_stream_t stream) {
    P2W_REQUIRE(num_tiles >= 0 && max_pts >= 1, "p2w_replacement_picks: bad sizes");
    const int64_t total = static_cast<int64_t>(num_tiles) * max_pts;
    if (total == 0) return P2W_OK;
    P2W_LAUNCH(replacement_picks_kernel, (unsigned)((total + 255) / 256), 256, 0, as_stream(stream))(members, seg, voxel_id, num_tiles, max_pts, seed, picks);
    return check_launch("p2w_replacement_picks");
}
