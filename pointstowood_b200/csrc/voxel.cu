// voxel.cu -- K4: voxel-grid ids (torch_cluster::grid), stable radix sort, and
// consecutive_cluster (one representative per occupied voxel, clusters ascending by id).
// Call sites in the reference: src/model.py:103-106 (voxelsample), src/preprocessing.py:33,58.
//
// All of it is HBM-bound integer work: coalesced streaming kernels, grids sized from the
// SM count, no host synchronisation (counts stay on the device).
#include "common.cuh"

namespace p2w {
namespace {

constexpr unsigned FULL = 0xffffffffu;

// ------------------------------------------------------------------ column min / max
__device__ __forceinline__ void atomic_max_f(float *a, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int *>(a), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned *>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_f(float *a, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int *>(a), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned *>(a), __float_as_uint(v));
}

__global__ void minmax_init_kernel(float *mn, float *mx, int dim) {
    if (threadIdx.x < dim) {
        mn[threadIdx.x] = __int_as_float(0x7f800000);
        mx[threadIdx.x] = __int_as_float(0xff800000);
    }
}

template <int DIM>
__global__ void __launch_bounds__(256) minmax_kernel(const float *__restrict__ pos, int64_t n, int ld,
                                                     float *__restrict__ mn, float *__restrict__ mx) {
    float lo[DIM], hi[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) { lo[d] = __int_as_float(0x7f800000); hi[d] = __int_as_float(0xff800000); }
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            const float v = pos[i * ld + d];
            lo[d] = fminf(lo[d], v);
            hi[d] = fmaxf(hi[d], v);
        }
    }
#pragma unroll
    for (int d = 0; d < DIM; d++) {
        for (int o = 16; o; o >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(FULL, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(FULL, hi[d], o));
        }
        if ((threadIdx.x & 31) == 0) { atomic_min_f(&mn[d], lo[d]); atomic_max_f(&mx[d], hi[d]); }
    }
}

// ------------------------------------------------------------------ grid ids
// id = sum_d (int64)((pos[d]-start[d]) / size[d]) * prod_{d'<d} ((int64)((end-start)/size)+1)
__global__ void __launch_bounds__(256) grid_kernel(const float *__restrict__ pos, int64_t n, int dim, int ld,
                                                   const int64_t *__restrict__ batch, const float *__restrict__ size,
                                                   const float *__restrict__ start, const float *__restrict__ end,
                                                   int64_t *__restrict__ ids) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t c = 0, k = 1;
    for (int d = 0; d < dim; d++) {
        const float p = __fsub_rn(pos[i * ld + d], start[d]);
        c += static_cast<int64_t>(__fdiv_rn(p, size[d])) * k;
        k *= static_cast<int64_t>(__fdiv_rn(__fsub_rn(end[d], start[d]), size[d])) + 1;
    }
    if (batch) {
        const float p = __fsub_rn(static_cast<float>(batch[i]), start[dim]);
        c += static_cast<int64_t>(__fdiv_rn(p, size[dim])) * k;
    }
    ids[i] = c;
}


// ------------------------------------------------------------------ grouped voxel keys (super-batches)
// Several reference batches ("groups" of consecutive tiles) travel through one launch.  The reference
// voxelises every batch on its own grid: start/end are the column min/max over the points of THAT batch
// (torch_geometric.nn.voxel_grid), so the origin is per group.
// One CTA per group reduces its points' xyz extent.
__global__ void __launch_bounds__(512) group_minmax_kernel(const float *__restrict__ pos, int ld,
                                                           const int64_t *__restrict__ tile_ptr,
                                                           const int64_t *__restrict__ group_ptr,
                                                           float *__restrict__ gmn, float *__restrict__ gmx) {
    __shared__ float s_lo[16][3], s_hi[16][3];
    const int g = blockIdx.x;
    const int64_t lo_i = tile_ptr[group_ptr[g]], hi_i = tile_ptr[group_ptr[g + 1]];
    float lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; d++) { lo[d] = __int_as_float(0x7f800000); hi[d] = __int_as_float(0xff800000); }
    for (int64_t i = lo_i + threadIdx.x; i < hi_i; i += blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const float v = pos[i * ld + d];
            lo[d] = fminf(lo[d], v);
            hi[d] = fmaxf(hi[d], v);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        for (int o = 16; o; o >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(FULL, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(FULL, hi[d], o));
        }
        if (lane == 0) { s_lo[warp][d] = lo[d]; s_hi[warp][d] = hi[d]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int d = threadIdx.x;
        float a = s_lo[0][d], b = s_hi[0][d];
        for (int w = 1; w < 16; w++) { a = fminf(a, s_lo[w][d]); b = fmaxf(b, s_hi[w][d]); }
        gmn[g * 3 + d] = a;
        gmx[g * 3 + d] = b;
    }
}

// key = (tile << spatial_bits) | spatial id on the group's grid.  Inside one group this orders points
// exactly like the reference's batch-major id (spatial + local_tile * cells); across groups it is
// tile-major.  *overflow is raised when a spatial id does not fit spatial_bits.
__global__ void __launch_bounds__(256) grid_grouped_kernel(const float *__restrict__ pos, int64_t n, int ld,
                                                           const int64_t *__restrict__ tile_ptr, int num_tiles,
                                                           const int64_t *__restrict__ group_ptr, int num_groups,
                                                           const float *__restrict__ gmn,
                                                           const float *__restrict__ gmx, float size,
                                                           int spatial_bits, uint64_t *__restrict__ keys,
                                                           int32_t *__restrict__ overflow) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = find_tile(tile_ptr, num_tiles, i);
    const int g = find_tile(group_ptr, num_groups, t);
    int64_t c = 0, k = 1;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const float s = gmn[g * 3 + d];
        const float p = __fsub_rn(pos[i * ld + d], s);
        c += static_cast<int64_t>(__fdiv_rn(p, size)) * k;
        k *= static_cast<int64_t>(__fdiv_rn(__fsub_rn(gmx[g * 3 + d], s), size)) + 1;
    }
    if (c >> spatial_bits) atomicExch(overflow, 1);
    keys[i] = (static_cast<uint64_t>(t) << spatial_bits) | static_cast<uint64_t>(c);
}

// ------------------------------------------------------------------ multi-block exclusive scan (int64)
constexpr int SCAN_T = 256, SCAN_V = 8, SCAN_TILE = SCAN_T * SCAN_V;

__device__ __forceinline__ int64_t block_excl_scan(int64_t v, int64_t *wsum, int64_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t s = v;
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(FULL, s, o);
        if (lane >= o) s += t;
    }
    if (lane == 31) wsum[warp] = s;
    __syncthreads();
    if (warp == 0) {
        int64_t w = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t t = __shfl_up_sync(FULL, w, o);
            if (lane >= o) w += t;
        }
        wsum[lane] = w;
    }
    __syncthreads();
    total = wsum[(blockDim.x >> 5) - 1];
    const int64_t r = (warp ? wsum[warp - 1] : 0) + s - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_T) scan_partial_kernel(const int64_t *__restrict__ in, int64_t n,
                                                              int64_t *__restrict__ bsum) {
    __shared__ int64_t wsum[32];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_TILE + threadIdx.x * SCAN_V;
    int64_t s = 0;
#pragma unroll
    for (int u = 0; u < SCAN_V; u++) s += (base + u < n) ? in[base + u] : 0;
    int64_t total;
    block_excl_scan(s, wsum, total);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

// single CTA: exclusive scan of a[0..n) in place, a[n] = total
__global__ void __launch_bounds__(1024) scan_single_kernel(int64_t *__restrict__ a, int64_t n) {
    __shared__ int64_t wsum[32];
    __shared__ int64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = i < n ? a[i] : 0;
        int64_t total;
        const int64_t e = block_excl_scan(v, wsum, total);
        const int64_t carry = carry_s;
        if (i < n) a[i] = carry + e;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) a[n] = carry_s;
}

__global__ void __launch_bounds__(SCAN_T) scan_apply_kernel(const int64_t *__restrict__ in, int64_t n,
                                                            const int64_t *__restrict__ bsum,
                                                            int64_t *__restrict__ out, int64_t *__restrict__ total_out) {
    __shared__ int64_t wsum[32];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_TILE + threadIdx.x * SCAN_V;
    int64_t v[SCAN_V], s = 0;
#pragma unroll
    for (int u = 0; u < SCAN_V; u++) { v[u] = (base + u < n) ? in[base + u] : 0; s += v[u]; }
    int64_t total;
    int64_t run = block_excl_scan(s, wsum, total) + bsum[blockIdx.x];
#pragma unroll
    for (int u = 0; u < SCAN_V; u++) {
        if (base + u < n) out[base + u] = run;
        run += v[u];
    }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total_out = bsum[gridDim.x];
}

inline int64_t scan_blocks(int64_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

// out[i] = sum_{j<i} in[i]; *total_out = sum.  bsum: scan_blocks(n)+1 int64.  in may alias out.
int scan_exclusive(const int64_t *in, int64_t *out, int64_t n, int64_t *bsum, int64_t *total_out, cudaStream_t st) {
    const int64_t nb = scan_blocks(n);
    if (nb == 0) return P2W_OK;
    P2W_LAUNCH(scan_partial_kernel, (unsigned)nb, SCAN_T, 0, st)(in, n, bsum);
    P2W_LAUNCH(scan_single_kernel, 1, 1024, 0, st)(bsum, nb);
    P2W_LAUNCH(scan_apply_kernel, (unsigned)nb, SCAN_T, 0, st)(in, n, bsum, out, total_out);
    return check_launch("scan_exclusive");
}

// ------------------------------------------------------------------ stable LSD radix sort
constexpr int RS_T = 256;              // threads per CTA (8 warps)
constexpr int RS_R = 8;                // rounds: each warp ranks 8 x 32 consecutive elements
constexpr int RS_TILE = RS_T * RS_R;   // 2048 elements per sub-tile
constexpr int RS_MAXB = 148 * 8;

struct SortPlan { int nb; int64_t per_block; };
inline SortPlan sort_plan(int64_t n) {
    SortPlan p;
    int64_t tiles = (n + RS_TILE - 1) / RS_TILE;
    p.nb = (int)(tiles < RS_MAXB ? tiles : RS_MAXB);
    if (p.nb < 1) p.nb = 1;
    int64_t tiles_per_block = (tiles + p.nb - 1) / p.nb;
    p.per_block = tiles_per_block * RS_TILE;
    p.nb = (int)((n + p.per_block - 1) / p.per_block);
    if (p.nb < 1) p.nb = 1;
    return p;
}

__global__ void __launch_bounds__(RS_T) rs_hist_kernel(const uint64_t *__restrict__ keys, int64_t n,
                                                       int64_t per_block, int shift, int nb,
                                                       int64_t *__restrict__ counts) {
    __shared__ unsigned hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t lo = static_cast<int64_t>(blockIdx.x) * per_block;
    const int64_t hi = (lo + per_block < n) ? lo + per_block : n;
    for (int64_t i = lo + threadIdx.x; i < hi; i += RS_T) atomicAdd(&hist[(keys[i] >> shift) & 255u], 1u);
    __syncthreads();
    counts[static_cast<int64_t>(threadIdx.x) * nb + blockIdx.x] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(RS_T) rs_scatter_kernel(const uint64_t *__restrict__ keys_in,
                                                          const int32_t *__restrict__ vals_in, int64_t n,
                                                          int64_t per_block, int shift, int nb,
                                                          const int64_t *__restrict__ offsets,
                                                          uint64_t *__restrict__ keys_out,
                                                          int32_t *__restrict__ vals_out) {
    __shared__ unsigned wcnt[RS_T / 32][256];   // per-warp digit counts of the current sub-tile
    __shared__ int64_t digit_base[256];         // global position of the next element of each digit
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    digit_base[threadIdx.x] = offsets[static_cast<int64_t>(threadIdx.x) * nb + blockIdx.x];
    const int64_t lo = static_cast<int64_t>(blockIdx.x) * per_block;
    const int64_t hi = (lo + per_block < n) ? lo + per_block : n;
    for (int64_t t0 = lo; t0 < hi; t0 += RS_TILE) {
        for (int w = 0; w < RS_T / 32; w++) wcnt[w][threadIdx.x] = 0;
        __syncthreads();
        uint64_t key[RS_R];
        int32_t val[RS_R];
        unsigned rank[RS_R];
        const int64_t wbase = t0 + warp * (32 * RS_R);
#pragma unroll
        for (int r = 0; r < RS_R; r++) {
            const int64_t i = wbase + r * 32 + lane;
            const bool ok = i < hi;
            key[r] = ok ? keys_in[i] : 0;
            val[r] = ok ? (vals_in ? vals_in[i] : static_cast<int32_t>(i)) : 0;
            const unsigned d = static_cast<unsigned>(key[r] >> shift) & 255u;
            const unsigned peers = __match_any_sync(FULL, ok ? d : 256u + 0u);
            const unsigned prior = ok ? wcnt[warp][d] : 0;
            __syncwarp();
            rank[r] = prior + __popc(peers & ((1u << lane) - 1u));
            if (ok && (peers & ((1u << lane) - 1u)) == 0) wcnt[warp][d] = prior + __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        // exclusive scan over warps for digit = threadIdx.x
        unsigned run = 0;
        for (int w = 0; w < RS_T / 32; w++) {
            const unsigned c = wcnt[w][threadIdx.x];
            wcnt[w][threadIdx.x] = run;
            run += c;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RS_R; r++) {
            const int64_t i = wbase + r * 32 + lane;
            if (i < hi) {
                const unsigned d = static_cast<unsigned>(key[r] >> shift) & 255u;
                const int64_t p = digit_base[d] + wcnt[warp][d] + rank[r];
                keys_out[p] = key[r];
                vals_out[p] = val[r];
            }
        }
        __syncthreads();
        digit_base[threadIdx.x] += run;
        __syncthreads();
    }
}

// ------------------------------------------------------------------ consecutive_cluster on sorted keys
__global__ void __launch_bounds__(256) head_flag_kernel(const uint64_t *__restrict__ keys, int64_t n,
                                                        int64_t *__restrict__ flag) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

__global__ void __launch_bounds__(256) unique_last_kernel(const uint64_t *__restrict__ keys,
                                                          const int32_t *__restrict__ idx, int64_t n,
                                                          const int64_t *__restrict__ excl,
                                                          int64_t *__restrict__ perm, int64_t *__restrict__ inverse,
                                                          int64_t *__restrict__ seg_start) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool head = (i == 0 || keys[i] != keys[i - 1]);
    const int64_t u = excl[i] + (head ? 1 : 0) - 1;
    if (inverse) inverse[idx[i]] = u;
    if (seg_start && head) seg_start[u] = i;
    if (i == n - 1 || keys[i] != keys[i + 1]) {
        if (perm) perm[u] = idx[i];
        if (seg_start && i == n - 1) seg_start[u + 1] = n;
    }
}

}  // namespace
}  // namespace p2w

using namespace p2w;

extern "C" int p2w_colminmax(const float *pos, int64_t n, int32_t dim, int32_t ld, float *mn, float *mx,
                             p2w_stream_t stream) {
    P2W_REQUIRE(dim >= 1 && dim <= 8 && ld >= dim, "p2w_colminmax: dim=%d ld=%d unsupported", dim, ld);
    cudaStream_t st = as_stream(stream);
    P2W_LAUNCH(minmax_init_kernel, 1, 32, 0, st)(mn, mx, dim);
    if (n > 0) {
        int64_t blocks = (n + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        switch (dim) {
#define P2W_MM(D) case D: P2W_LAUNCH(minmax_kernel<D>, (unsigned)blocks, 256, 0, st)(pos, n, ld, mn, mx); break;
            P2W_MM(1) P2W_MM(2) P2W_MM(3) P2W_MM(4) P2W_MM(5) P2W_MM(6) P2W_MM(7) P2W_MM(8)
#undef P2W_MM
        }
    }
    return check_launch("p2w_colminmax");
}

extern "C" int p2w_grid(const float *pos, int64_t n, int32_t dim, int32_t ld, const int64_t *batch, const float *size,
                        const float *start, const float *end, int64_t *ids, p2w_stream_t stream) {
    P2W_REQUIRE(dim >= 1 && ld >= dim, "p2w_grid: dim=%d ld=%d unsupported", dim, ld);
    if (n == 0) return P2W_OK;
    P2W_LAUNCH(grid_kernel, (unsigned)((n + 255) / 256), 256, 0, as_stream(stream))(pos, n, dim, ld, batch, size, start, end, ids);
    return check_launch("p2w_grid");
}

extern "C" int p2w_voxel_keys_grouped(const float *pos, int64_t n, int32_t ld, const int64_t *tile_ptr,
                                      int32_t num_tiles, const int64_t *group_ptr, int32_t num_groups, float size,
                                      int32_t spatial_bits, float *group_min, float *group_max, uint64_t *keys,
                                      int32_t *overflow, p2w_stream_t stream) {
    P2W_REQUIRE(ld >= 3 && num_tiles >= 1 && num_groups >= 1, "p2w_voxel_keys_grouped: bad sizes");
    P2W_REQUIRE(spatial_bits >= 1 && spatial_bits <= 48, "p2w_voxel_keys_grouped: spatial_bits=%d", spatial_bits);
    P2W_REQUIRE(size > 0.f, "p2w_voxel_keys_grouped: size must be positive");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(overflow, 0, sizeof(int32_t), st);
    if (n == 0) return check_launch("p2w_voxel_keys_grouped");
    P2W_LAUNCH(group_minmax_kernel, num_groups, 512, 0, st)(pos, ld, tile_ptr, group_ptr, group_min, group_max);
    P2W_LAUNCH(grid_grouped_kernel, (unsigned)((n + 255) / 256), 256, 0, st)(pos, n, ld, tile_ptr, num_tiles, group_ptr,
                                                                             num_groups, group_min, group_max, size,
                                                                             spatial_bits, keys, overflow);
    return check_launch("p2w_voxel_keys_grouped");
}

extern "C" size_t p2w_sort_ws_bytes(int64_t n) {
    const SortPlan p = sort_plan(n);
    const int64_t ncnt = 256 * static_cast<int64_t>(p.nb);
    // counters (+1 total) + scan block sums + ping-pong key/value buffers
    return static_cast<size_t>(8 * (ncnt + 1) + 8 * (scan_blocks(ncnt) + 2) + 8 * n + 4 * n + 64);
}

extern "C" int p2w_sort_pairs(const uint64_t *keys_in, const int32_t *vals_in, uint64_t *keys_out, int32_t *vals_out,
                              int64_t n, int32_t key_bits, void *ws, p2w_stream_t stream) {
    P2W_REQUIRE(key_bits >= 0 && key_bits <= 64, "p2w_sort_pairs: key_bits=%d", key_bits);
    P2W_REQUIRE(n < (int64_t(1) << 31), "p2w_sort_pairs: n must fit int32 values");
    if (n == 0) return P2W_OK;
    cudaStream_t st = as_stream(stream);
    const SortPlan p = sort_plan(n);
    const int64_t ncnt = 256 * static_cast<int64_t>(p.nb);
    int64_t *counts = static_cast<int64_t *>(ws);
    int64_t *bsum = counts + ncnt + 1;
    uint64_t *ktmp = reinterpret_cast<uint64_t *>(bsum + scan_blocks(ncnt) + 2);
    int32_t *vtmp = reinterpret_cast<int32_t *>(ktmp + n);
    int passes = (key_bits + 7) / 8;
    if (passes == 0) passes = 1;
    // ping-pong so that the last pass lands in (keys_out, vals_out)
    const uint64_t *kin = keys_in;
    const int32_t *vin = vals_in;   // NULL: values are the identity permutation
    for (int ps = 0; ps < passes; ps++) {
        const bool to_out = ((passes - 1 - ps) % 2) == 0;
        uint64_t *kout = to_out ? keys_out : ktmp;
        int32_t *vout = to_out ? vals_out : vtmp;
        P2W_LAUNCH(rs_hist_kernel, p.nb, RS_T, 0, st)(kin, n, p.per_block, ps * 8, p.nb, counts);
        int rc = scan_exclusive(counts, counts, ncnt, bsum, nullptr, st);
        if (rc) return rc;
        P2W_LAUNCH(rs_scatter_kernel, p.nb, RS_T, 0, st)(kin, vin, n, p.per_block, ps * 8, p.nb, counts, kout, vout);
        kin = kout;
        vin = vout;
    }
    return check_launch("p2w_sort_pairs");
}

extern "C" size_t p2w_unique_ws_bytes(int64_t n) { return static_cast<size_t>(8 * (n + 1) + 8 * (scan_blocks(n) + 2)); }

extern "C" int p2w_unique_last(const uint64_t *sorted_keys, const int32_t *sorted_idx, int64_t n, int64_t *perm,
                               int64_t *inverse, int64_t *seg_start, int64_t *num_unique, void *ws,
                               p2w_stream_t stream) {
    cudaStream_t st = as_stream(stream);
    if (n == 0) {
        cudaMemsetAsync(num_unique, 0, 8, st);
        return check_launch("p2w_unique_last");
    }
    int64_t *flag = static_cast<int64_t *>(ws);
    int64_t *bsum = flag + n + 1;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    P2W_LAUNCH(head_flag_kernel, blocks, 256, 0, st)(sorted_keys, n, flag);
    int rc = scan_exclusive(flag, flag, n, bsum, num_unique, st);
    if (rc) return rc;
    P2W_LAUNCH(unique_last_kernel, blocks, 256, 0, st)(sorted_keys, sorted_idx, n, flag, perm, inverse, seg_start);
    return check_launch("p2w_unique_last");
}
