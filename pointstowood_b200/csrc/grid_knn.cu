// grid_knn.cu -- K1 / K2 with a per-tile cell list: exact kNN and radius search that examine a
// few hundred candidates per query instead of the whole tile, with results bit-identical to the
// brute-force sweep (neighbors.cu) and therefore to torch_cluster's CUDA kernels
// (call sites src/model.py:118,120,149).
//
// Why it is exact.  The result of both searches is a function of the SET of (FP32 distance, index)
// pairs -- kNN keeps the k smallest pairs in lexicographic order, radius keeps the k smallest
// indices with d < r^2 -- so the order in which candidates are offered is irrelevant (topk.cuh).
// A query visits the cells of its tile's uniform grid in growing Chebyshev shells around its own
// cell and stops once every unvisited cell is provably farther than the current k-th distance
// (kNN) or than r (radius): an unvisited point lies beyond a face of the visited block, so its
// distance is at least the query's distance to the nearest face that is not a grid boundary.  The
// test is made with a relative + absolute safety margin far above the FP32 rounding of the cell
// assignment and of the distance, so no point that belongs to the result can be skipped.
//
// Build per call (sources of all tiles at once), a counting sort into a dense cell table:
//   grid_plan_kernel    one CTA per tile: bounding box, then an occupancy pyramid (64^3 bitmap in
//                       Morton order in shared memory, OR-reduced level by level) picks the cell size
//                       so that an occupied cell holds ~tau sources (tau ~ 0.45 k covers both surface-
//                       like and volume-like tiles), capped at 4 cells per source;
//   grid_base_kernel    exclusive scan of the tiles' cell counts (table offsets);
//   grid_count_kernel   cell id per source + histogram;  scan32 -> cell_start (global positions);
//   grid_scatter_kernel sources re-ordered by cell as float4 (x, y, z, index).
// Query: one warp per query.  The lanes look up the row segments of the current shell in the cell
// table (a run of cells along x is one contiguous range; nearest rows first), a warp scan flattens the
// segments so that every step evaluates 32 candidates, and the survivors of the threshold test
// enter the warp-distributed top-k -- one by one, or through a bitonic sort + merge of the whole
// step when many survive.
#include <stdlib.h>

#include "common.cuh"

namespace p2w {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int GL_MAX = 6;        // finest planning level: 64 cells along the longest axis
constexpr int GDIM_MAX = 64;        // cells per axis when the cell size is planned from the occupancy pyramid
constexpr int GDIM_HINT_MAX = 1024;
constexpr int HINT_CELLS_PER_SOURCE = 4;    // table cap of the hinted mode, same as the planned mode (16 measured slower) // ... and when the caller gives the cell size (plot-wide searches)
constexpr int MERGE_MIN = 5;      // survivors per step above which the step is sorted and merged at once

struct __align__(16) GridTile {
    float ox, oy, oz, inv_h;
    float h;
    int nx, ny, nz;
    int64_t base;      // offset of this tile's cells in the cell table
    int64_t ncells;    // nx * ny * nz
};

__device__ __forceinline__ int axis_cell(float p, float o, float inv_h, int n) {
    // the clamp also places queries outside the sources' bounding box in a boundary cell
    const float f = __fmul_rn(__fsub_rn(p, o), inv_h);
    return static_cast<int>(fminf(fmaxf(f, 0.f), static_cast<float>(n - 1)));
}

__device__ __forceinline__ unsigned spread6(unsigned v) {   // 6 bits -> every third bit
    v = (v | (v << 8)) & 0x0000F00Fu;   // not needed beyond 6 bits, but keep the classic ladder
    v = (v | (v << 4)) & 0x000C30C3u;
    v = (v | (v << 2)) & 0x00249249u;
    return v;
}

// ------------------------------------------------------------------ plan: one CTA per tile
// ---- bounding box of ONE huge tile (plot-wide searches) by all SMs; grid_plan_kernel then skips its own pass
__device__ __forceinline__ void atomic_max_f(float *a, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int *>(a), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned *>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_f(float *a, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int *>(a), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned *>(a), __float_as_uint(v));
}
__global__ void box_init_kernel(float *box) {
    if (threadIdx.x < 3) box[threadIdx.x] = __int_as_float(0x7f800000);
    else if (threadIdx.x < 6) box[threadIdx.x] = __int_as_float(0xff800000);
}
__global__ void __launch_bounds__(256) box_kernel(const float *__restrict__ x, int64_t n, float *__restrict__ box) {
    float lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; d++) { lo[d] = __int_as_float(0x7f800000); hi[d] = __int_as_float(0xff800000); }
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const float v = x[i * 3 + d];
            lo[d] = fminf(lo[d], v);
            hi[d] = fmaxf(hi[d], v);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
        for (int o = 16; o; o >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(FULL, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(FULL, hi[d], o));
        }
        if ((threadIdx.x & 31) == 0) { atomic_min_f(&box[d], lo[d]); atomic_max_f(&box[3 + d], hi[d]); }
    }
}

__global__ void __launch_bounds__(512) grid_plan_kernel(const float *__restrict__ x,
                                                        const int64_t *__restrict__ ptr_x, float tau, float cell_hint,
                                                        const float *__restrict__ pre_box,
                                                        GridTile *__restrict__ grid) {
    __shared__ unsigned bits[8192];          // 64^3 occupancy bits, Morton order
    __shared__ float s_lo[16][3], s_hi[16][3];
    __shared__ float s_box[6];
    __shared__ unsigned s_occ[GL_MAX + 1];
    const int b = blockIdx.x;
    const int64_t i0 = ptr_x[b], i1 = ptr_x[b + 1];
    const int64_t n = i1 - i0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; d++) { lo[d] = __int_as_float(0x7f800000); hi[d] = __int_as_float(0xff800000); }
    if (pre_box == nullptr) {
        for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const float v = x[i * 3 + d];
                lo[d] = fminf(lo[d], v);
                hi[d] = fmaxf(hi[d], v);
            }
        }
    } else {
#pragma unroll
        for (int d = 0; d < 3; d++) { lo[d] = pre_box[d]; hi[d] = pre_box[3 + d]; }
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
        for (int o = 16; o; o >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(FULL, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(FULL, hi[d], o));
        }
        if (lane == 0) { s_lo[warp][d] = lo[d]; s_hi[warp][d] = hi[d]; }
    }
    // pyramid depth: a level with more than 4 cells per source can never be chosen (the table cap), so small
    // tiles get a shallow pyramid (8^LB bits to clear, fill and reduce instead of 64^3)
    int LB = GL_MAX;
    while (LB > 2 && (int64_t(1) << (3 * (LB - 1))) >= 4 * n) LB--;
    const int words0 = (1 << (3 * LB)) >> 5;
    if (threadIdx.x <= GL_MAX) s_occ[threadIdx.x] = 0;
    if (n > 32 && cell_hint <= 0.f)
        for (int w = threadIdx.x; w < words0; w += blockDim.x) bits[w] = 0;
    __syncthreads();
    if (threadIdx.x < 3) {
        const int d = threadIdx.x;
        float a = s_lo[0][d], c = s_hi[0][d];
        for (int w = 1; w < 16; w++) { a = fminf(a, s_lo[w][d]); c = fmaxf(c, s_hi[w][d]); }
        s_box[d] = a;
        s_box[3 + d] = c;
    }
    __syncthreads();
    const float ox = s_box[0], oy = s_box[1], oz = s_box[2];
    const float ex = s_box[3] - ox, ey = s_box[4] - oy, ez = s_box[5] - oz;
    const float maxext = fmaxf(ex, fmaxf(ey, ez));
    GridTile g;
    g.ox = ox; g.oy = oy; g.oz = oz;
    g.nx = g.ny = g.nz = 1;
    g.h = 1.f;
    if (n <= 0) {
        g.ox = g.oy = g.oz = 0.f;
    } else if (n <= 32 || !(maxext > 0.f) || !(maxext < 1e30f)) {
        g.h = (maxext > 0.f && maxext < 1e30f) ? maxext * 1.0001f : 1.f;      // one cell: brute force inside the tile
    } else if (cell_hint > 0.f) {
        // the caller knows the scale (plot-wide searches): start from its cell size, respect the table cap
        float h = cell_hint;
        const int64_t cap = n * HINT_CELLS_PER_SOURCE > 64 ? n * HINT_CELLS_PER_SOURCE : 64;
        for (;;) {
            const float fx = ex / h, fy = ey / h, fz = ez / h;
            if (fx < static_cast<float>(GDIM_HINT_MAX) && fy < static_cast<float>(GDIM_HINT_MAX) &&
                fz < static_cast<float>(GDIM_HINT_MAX)) {
                g.nx = static_cast<int>(fx) + 1;
                g.ny = static_cast<int>(fy) + 1;
                g.nz = static_cast<int>(fz) + 1;
                if (static_cast<int64_t>(g.nx) * g.ny * g.nz <= cap) break;
            }
            h *= 1.26f;
        }
        g.h = h;
    } else {
        const int nb = 1 << LB;
        const float scale = static_cast<float>(nb) / maxext;
        for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
            const unsigned cx = static_cast<unsigned>(axis_cell(x[i * 3 + 0], ox, scale, nb));
            const unsigned cy = static_cast<unsigned>(axis_cell(x[i * 3 + 1], oy, scale, nb));
            const unsigned cz = static_cast<unsigned>(axis_cell(x[i * 3 + 2], oz, scale, nb));
            const unsigned m = spread6(cx) | (spread6(cy) << 1) | (spread6(cz) << 2);
            atomicOr(&bits[m >> 5], 1u << (m & 31u));
        }
        __syncthreads();
        // occupancy pyramid: level L has 8^L bits; 8 sibling bits are one byte
        int words = words0;
        for (int L = LB; L >= 1; L--) {
            unsigned c = 0;
            for (int w = threadIdx.x; w < words; w += blockDim.x) c += __popc(bits[w]);
            for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
            if (lane == 0 && c) atomicAdd(&s_occ[L], c);
            const int nwords = words >= 8 ? words / 8 : 1;
            unsigned nw[2] = {0, 0};   // a thread owns at most 2 words of the next level (1024 words / 512 threads)
            for (int t = 0; t < 2; t++) {
                const int j = threadIdx.x + t * blockDim.x;
                if (j < nwords) {
                    unsigned v = 0;
                    for (int u = 0; u < 8; u++) {
                        const int src = j * 8 + u;
                        const unsigned wv = src < words ? bits[src] : 0u;
#pragma unroll
                        for (int by = 0; by < 4; by++)
                            if ((wv >> (8 * by)) & 0xffu) v |= 1u << (u * 4 + by);
                    }
                    nw[t] = v;
                }
            }
            __syncthreads();
            for (int t = 0; t < 2; t++) {
                const int j = threadIdx.x + t * blockDim.x;
                if (j < nwords) bits[j] = nw[t];
            }
            words = nwords;
            __syncthreads();
        }
        // finest level whose occupied cells hold >= tau sources on average
        int L = 0;
        for (int l = LB; l >= 1; l--) {
            if (static_cast<float>(n) >= tau * static_cast<float>(s_occ[l])) { L = l; break; }
        }
        if (L == 0) {
            g.h = maxext * 1.0001f;
        } else {
            float h = maxext / static_cast<float>(1 << L);
            if (L < LB) {
                // between two dyadic levels: shrink h by the local dimension (2 = surface, 3 = volume)
                const float m = static_cast<float>(n) / static_cast<float>(s_occ[L]);
                float dim = log2f(static_cast<float>(s_occ[L + 1]) / static_cast<float>(s_occ[L]));
                dim = fminf(fmaxf(dim, 1.5f), 3.f);
                const float f = fminf(fmaxf(powf(tau / m, 1.f / dim), 0.5f), 1.f);
                h *= f;
            }
            // dense cell table: at most max(4 n, 64) cells per tile (grid_table_bound)
            const int64_t cap = n * 4 > 64 ? n * 4 : 64;
            for (;;) {
                g.nx = min(static_cast<int>(ex / h) + 1, GDIM_MAX);
                g.ny = min(static_cast<int>(ey / h) + 1, GDIM_MAX);
                g.nz = min(static_cast<int>(ez / h) + 1, GDIM_MAX);
                if (static_cast<int64_t>(g.nx) * g.ny * g.nz <= cap) break;
                h *= 1.26f;
            }
            g.h = h;
        }
    }
    g.inv_h = 1.f / g.h;
    g.base = 0;
    g.ncells = static_cast<int64_t>(g.nx) * g.ny * g.nz;
    if (threadIdx.x == 0) grid[b] = g;
}

// table offsets: base[b] = sum of the cell counts of the tiles before b (single CTA; T is small)
__global__ void __launch_bounds__(1024) grid_base_kernel(GridTile *__restrict__ grid, int T) {
    __shared__ int64_t wsum[32];
    __shared__ int64_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int t0 = 0; t0 < T; t0 += 1024) {
        const int t = t0 + threadIdx.x;
        const int64_t v = t < T ? grid[t].ncells : 0;
        int64_t sc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t u = __shfl_up_sync(FULL, sc, o);
            if (lane >= o) sc += u;
        }
        if (lane == 31) wsum[warp] = sc;
        __syncthreads();
        if (warp == 0) {
            int64_t w = wsum[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t u = __shfl_up_sync(FULL, w, o);
                if (lane >= o) w += u;
            }
            wsum[lane] = w;
        }
        __syncthreads();
        const int64_t carry = carry_s;
        if (t < T) grid[t].base = carry + (warp ? wsum[warp - 1] : 0) + sc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wsum[31];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) grid_count_kernel(const float *__restrict__ x,
                                                         const int64_t *__restrict__ ptr_x, int T, int64_t n,
                                                         const GridTile *__restrict__ grid,
                                                         uint32_t *__restrict__ slot, uint32_t *__restrict__ count) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = find_tile(ptr_x, T, i);
    const GridTile g = grid[b];
    const int cx = axis_cell(x[i * 3 + 0], g.ox, g.inv_h, g.nx);
    const int cy = axis_cell(x[i * 3 + 1], g.oy, g.inv_h, g.ny);
    const int cz = axis_cell(x[i * 3 + 2], g.oz, g.inv_h, g.nz);
    const int64_t c = g.base + cx + g.nx * (cy + g.ny * cz);
    slot[i] = static_cast<uint32_t>(c);
    atomicAdd(&count[c], 1u);
}

// ---- exclusive scan of uint32 counters (three launches: block sums, their scan, apply)
constexpr int SC_T = 256, SC_V = 8, SC_TILE = SC_T * SC_V;

__device__ __forceinline__ uint32_t block_scan32(uint32_t v, uint32_t *wsum, uint32_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t sc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, sc, o);
        if (lane >= o) sc += t;
    }
    if (lane == 31) wsum[warp] = sc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, w, o);
            if (lane >= o) w += t;
        }
        wsum[lane] = w;
    }
    __syncthreads();
    total = wsum[(blockDim.x >> 5) - 1];
    const uint32_t r = (warp ? wsum[warp - 1] : 0) + sc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SC_T) scan32_partial_kernel(const uint32_t *__restrict__ in, int64_t n,
                                                              uint32_t *__restrict__ bsum) {
    __shared__ uint32_t wsum[32];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * SC_TILE + threadIdx.x * SC_V;
    uint32_t sum = 0;
    if (base + SC_V <= n) {
        const uint4 a = *reinterpret_cast<const uint4 *>(in + base), c = *reinterpret_cast<const uint4 *>(in + base + 4);
        sum = a.x + a.y + a.z + a.w + c.x + c.y + c.z + c.w;
    } else {
        for (int u = 0; u < SC_V; u++) sum += (base + u < n) ? in[base + u] : 0u;
    }
    uint32_t total;
    block_scan32(sum, wsum, total);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan32_single_kernel(uint32_t *__restrict__ a, int64_t n) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = i < n ? a[i] : 0u;
        uint32_t total;
        const uint32_t e = block_scan32(v, wsum, total);
        const uint32_t carry = carry_s;
        if (i < n) a[i] = carry + e;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
}

// out[i] = offset + exclusive prefix; out may alias in.  out[n] = grand total (+ offset) for the last block.
__global__ void __launch_bounds__(SC_T) scan32_apply_kernel(const uint32_t *__restrict__ in, int64_t n,
                                                            const uint32_t *__restrict__ bsum,
                                                            uint32_t *__restrict__ out) {
    __shared__ uint32_t wsum[32];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * SC_TILE + threadIdx.x * SC_V;
    uint32_t v[SC_V], sum = 0;
#pragma unroll
    for (int u = 0; u < SC_V; u++) { v[u] = (base + u < n) ? in[base + u] : 0u; sum += v[u]; }
    uint32_t total;
    uint32_t run = block_scan32(sum, wsum, total) + bsum[blockIdx.x];
#pragma unroll
    for (int u = 0; u < SC_V; u++) {
        if (base + u < n) out[base + u] = run;
        run += v[u];
    }
}

// spts[cell_start[c] + k] = (x, y, z, index) for the k-th source that claims cell c (order inside a
// cell is arbitrary: the search result does not depend on it)
__global__ void __launch_bounds__(256) grid_scatter_kernel(const float *__restrict__ x, int64_t n,
                                                           const uint32_t *__restrict__ slot,
                                                           const uint32_t *__restrict__ cell_start,
                                                           uint32_t *__restrict__ fill, float4 *__restrict__ spts) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = slot[i];
    const uint32_t p = cell_start[c] + atomicAdd(&fill[c], 1u);
    spts[p] = make_float4(x[i * 3 + 0], x[i * 3 + 1], x[i * 3 + 2], __int_as_float(static_cast<int>(i)));
}

// ---- selection state.  A candidate is ONE 64-bit key: (FP32 bits of d) << 32 | index.  d >= +0, so
// the unsigned order of the keys is the lexicographic order of (d, index) -- a compare is two
// integer instructions and a warp minimum is two REDUX.  Radius mode uses key = index.
using key_t = unsigned long long;
constexpr key_t KEY_NONE = ~0ull;                                   // filler, above every real key
__device__ __forceinline__ key_t key_sentinel() {                   // upstream's (1e10, -1) initial entry
    return (static_cast<key_t>(__float_as_uint(1e10f)) << 32) | 0xFFFFFFFFull;
}

template <int S>
struct TopK64 {
    key_t key[S];   // entry e of the ascending list lives in slot e/32 of lane e%32
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < S; s++) key[s] = key_sentinel();
    }
    __device__ __forceinline__ void insert(key_t ck, int lane) {
        int pos = 0;
#pragma unroll
        for (int s = 0; s < S; s++) pos += __popc(__ballot_sync(FULL, key[s] < ck));
#pragma unroll
        for (int s = S - 1; s >= 0; s--) {
            key_t up = __shfl_up_sync(FULL, key[s], 1);
            if (s > 0) {
                const key_t w = __shfl_sync(FULL, key[s - 1], 31);
                if (lane == 0) up = w;
            }
            const int e = s * 32 + lane;
            if (e == pos) key[s] = ck;
            else if (e > pos) key[s] = up;
        }
    }
    __device__ __forceinline__ key_t kth(int k) const {
        const int e = k - 1;
        key_t v = key[0];
#pragma unroll
        for (int s = 1; s < S; s++)
            if ((e >> 5) == s) v = key[s];
        return __shfl_sync(FULL, v, e & 31);
    }
};

__device__ __forceinline__ void cmpx64(key_t &v, int j, bool take_min) {
    const key_t o = __shfl_xor_sync(FULL, v, j);
    v = take_min ? (o < v ? o : v) : (o > v ? o : v);
}
// top (ascending, entry e in lane e) <- the 32 smallest of top U batch (batch in any order)
__device__ __forceinline__ void merge32(key_t &top, key_t batch, int lane) {
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
        for (int j = k2 >> 1; j > 0; j >>= 1) cmpx64(batch, j, ((lane & k2) == 0) == ((lane & j) == 0));
    }
    const key_t r = __shfl_sync(FULL, batch, 31 - lane);
    top = r < top ? r : top;                                  // the 32 smallest, as a bitonic sequence
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) cmpx64(top, j, (lane & j) == 0);
}

// 64-entry list (two slots per lane, entry e in slot e/32 of lane e%32) <- the 64 smallest of list U batch.
// The largest 32 of (upper half U batch) cannot be among the 64 smallest: each of them has 32 smaller
// elements there and, because one of those comes from the upper half, the whole lower half below it too.
__device__ __forceinline__ void merge64(key_t &lo, key_t &hi, key_t batch, int lane) {
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
        for (int j = k2 >> 1; j > 0; j >>= 1) cmpx64(batch, j, ((lane & k2) == 0) == ((lane & j) == 0));
    }
    key_t r = __shfl_sync(FULL, batch, 31 - lane);
    key_t m = r < hi ? r : hi;                                // smallest 32 of (upper half U batch), bitonic
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) cmpx64(m, j, (lane & j) == 0);
    r = __shfl_sync(FULL, m, 31 - lane);
    hi = r > lo ? r : lo;                                     // bitonic halves of (lower half U m)
    lo = r < lo ? r : lo;
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        cmpx64(lo, j, (lane & j) == 0);
        cmpx64(hi, j, (lane & j) == 0);
    }
}

constexpr int SMALL_K = 6;        // up to this k a crowded step is reduced by k warp-minimum rounds

// Offers the sources of up to 32 contiguous segments (lane l: spts[start, start+len)) to the top-k.
template <int S, bool RADIUS>
__device__ __forceinline__ void scan_segments(int len, uint32_t start, const float4 *__restrict__ spts, float qx,
                                              float qy, float qz, float r2, int k, int lane, TopK64<S> &top,
                                              key_t &thr, int &hits, unsigned long long &evals) {
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    const int excl = incl - len;
    evals += static_cast<unsigned long long>(total);
    for (int base = 0; base < total; base += 32) {
        const int t = base + lane;
        int sg = 0;   // number of segments that end at or before t
#pragma unroll
        for (int step = 16; step; step >>= 1) {
            const int v = __shfl_sync(FULL, incl, sg + step - 1);
            if (v <= t) sg += step;
        }
        const uint32_t sst = __shfl_sync(FULL, start, sg);
        const int sex = __shfl_sync(FULL, excl, sg);
        const bool valid = t < total;
        const float4 c = __ldg(spts + (valid ? sst + static_cast<uint32_t>(t - sex) : 0u));
        const float dx = __fsub_rn(c.x, qx), dy = __fsub_rn(c.y, qy), dz = __fsub_rn(c.z, qz);
        const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
        const uint32_t ji = __float_as_uint(c.w);
        key_t ck;
        bool in_range;
        if (RADIUS) {
            ck = ji;
            in_range = valid && d < r2;
            hits += __popc(__ballot_sync(FULL, in_range));
        } else {
            ck = (static_cast<key_t>(__float_as_uint(d)) << 32) | ji;
            in_range = valid && d < 1e10f;     // upstream never admits d >= its 1e10 initial entries
        }
        const bool hit = in_range && ck < thr;
        unsigned m = __ballot_sync(FULL, hit);
        if (!m) continue;
        if (S == 2 && __popc(m) > 2 * MERGE_MIN) {
            merge64(top.key[0], top.key[S - 1], hit ? ck : KEY_NONE, lane);
            thr = top.kth(k);
            continue;
        }
        if (S == 1 && __popc(m) > MERGE_MIN) {
            if (k <= SMALL_K) {
                bool live = hit;
                for (int round = 0; round < k; round++) {
                    const unsigned hi = live ? static_cast<unsigned>(ck >> 32) : 0xFFFFFFFFu;
                    const unsigned mh = __reduce_min_sync(FULL, hi);
                    const bool tie = live && hi == mh;
                    const unsigned lo = tie ? static_cast<unsigned>(ck) : 0xFFFFFFFFu;
                    const unsigned ml = __reduce_min_sync(FULL, lo);
                    const key_t best = (static_cast<key_t>(mh) << 32) | ml;
                    if (!(best < thr)) break;              // nothing left that can enter (also: no live lane)
                    top.insert(best, lane);
                    thr = top.kth(k);
                    if (tie && lo == ml) live = false;
                }
            } else {
                merge32(top.key[0], hit ? ck : KEY_NONE, lane);
                thr = top.kth(k);
            }
            continue;
        }
        while (m) {
            const int l = __ffs(m) - 1;
            m &= m - 1;
            const key_t cand = __shfl_sync(FULL, ck, l);
            if (cand < thr) {
                top.insert(cand, lane);
                thr = top.kth(k);
            }
        }
    }
}

// the 27 cells of the first ring ordered own cell, 6 face, 12 edge, 8 corner neighbours; 2 bits per entry
constexpr unsigned long long NB_DX = 0x2a802a95402551ull, NB_DY = 0x28282528251945ull, NB_DZ = 0x22221862185615ull;

// S: top-k slots per lane (k <= 32 S).  RADIUS: the k lowest indices with d < r2 instead of kNN.
template <int S, bool RADIUS>
__global__ void __launch_bounds__(256, 3) grid_query_kernel(const float4 *__restrict__ spts,
                                                         const uint32_t *__restrict__ cell_start,
                                                         const GridTile *__restrict__ grid,
                                                         const float *__restrict__ y,
                                                         const int64_t *__restrict__ ptr_x,
                                                         const int64_t *__restrict__ ptr_y, int T, int64_t ny,
                                                         int k, float r2, int32_t *__restrict__ nbr,
                                                         float *__restrict__ d2out, int32_t *__restrict__ cnt_out,
                                                         unsigned long long *__restrict__ pair_evals) {
    const int lane = threadIdx.x & 31;
    unsigned long long evals = 0;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int64_t nchunks = (ny + 31) >> 5;
    // a warp takes 32 consecutive queries at a time: coalesced loads, one tile lookup per lane
    for (int64_t chunk = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; chunk < nchunks;
         chunk += nwarps) {
        const int64_t qmine = chunk * 32 + lane;
        const bool have = qmine < ny;
        const int b_l = have ? find_tile(ptr_y, T, qmine) : 0;
        const float qx_l = have ? y[qmine * 3 + 0] : 0.f, qy_l = have ? y[qmine * 3 + 1] : 0.f,
                    qz_l = have ? y[qmine * 3 + 2] : 0.f;
        const int nq = static_cast<int>(ny - chunk * 32 < 32 ? ny - chunk * 32 : 32);
        for (int u = 0; u < nq; u++) {
            const int64_t q = chunk * 32 + u;
            const int b = __shfl_sync(FULL, b_l, u);
            const float qx = __shfl_sync(FULL, qx_l, u), qy = __shfl_sync(FULL, qy_l, u),
                        qz = __shfl_sync(FULL, qz_l, u);
            const int64_t s0 = ptr_x[b], s1 = ptr_x[b + 1];
            TopK64<S> top;
            top.init();
            key_t thr = key_sentinel();
            int hits = 0;
            if (s1 > s0) {
                const GridTile g = grid[b];
                const int cx = axis_cell(qx, g.ox, g.inv_h, g.nx);
                const int cy = axis_cell(qy, g.oy, g.inv_h, g.ny);
                const int cz = axis_cell(qz, g.oz, g.inv_h, g.nz);
                const int nmax = max(g.nx, max(g.ny, g.nz));
                const float margin = 1e-5f * g.h * static_cast<float>(nmax) +
                                     4e-7f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + fabsf(g.ox) + fabsf(g.oy) + fabsf(g.oz));
                // safe (shrunk) gaps from the query to the faces of its own cell: lower bounds for whole cells / rows
                const float fx = g.ox + static_cast<float>(cx) * g.h, fy = g.oy + static_cast<float>(cy) * g.h,
                            fz = g.oz + static_cast<float>(cz) * g.h;
                const float k1 = 1.f - 1e-4f;
                const float lox = fmaxf((qx - fx) * k1 - margin, 0.f), hix = fmaxf((fx + g.h - qx) * k1 - margin, 0.f);
                const float loy = fmaxf((qy - fy) * k1 - margin, 0.f), hiy = fmaxf((fy + g.h - qy) * k1 - margin, 0.f);
                const float loz = fmaxf((qz - fz) * k1 - margin, 0.f), hiz = fmaxf((fz + g.h - qz) * k1 - margin, 0.f);
                const float hs = g.h * k1;             // a further whole cell in between adds at least this much
                // ---- ring 1, cell by cell: own cell + face neighbours, then edge, then corner neighbours;
                //      a cell whose nearest point is provably beyond the current k-th distance is skipped
                {
                    const int ddx = static_cast<int>((NB_DX >> (2 * lane)) & 3ull) - 1;
                    const int ddy = static_cast<int>((NB_DY >> (2 * lane)) & 3ull) - 1;
                    const int ddz = static_cast<int>((NB_DZ >> (2 * lane)) & 3ull) - 1;
                    const int xx = cx + ddx, yy = cy + ddy, zz = cz + ddz;
                    const bool cell_ok = lane < 27 && xx >= 0 && xx < g.nx && yy >= 0 && yy < g.ny && zz >= 0 && zz < g.nz;
                    uint32_t a = 0;
                    int len = 0;
                    if (cell_ok) {
                        const int64_t c = g.base + xx + g.nx * (yy + g.ny * zz);
                        a = __ldg(cell_start + c);
                        len = static_cast<int>(__ldg(cell_start + c + 1) - a);
                    }
                    // squared distance from the query to the neighbour cell (safe lower bound)
                    const float sx = ddx < 0 ? lox : (ddx > 0 ? hix : 0.f);
                    const float sy = ddy < 0 ? loy : (ddy > 0 ? hiy : 0.f);
                    const float sz = ddz < 0 ? loz : (ddz > 0 ? hiz : 0.f);
                    const float b2 = sx * sx + sy * sy + sz * sz;
#pragma unroll 1
                    for (int stage = 0; stage < 3; stage++) {
                        const int lo = stage == 0 ? 0 : (stage == 1 ? 7 : 19), hi = stage == 0 ? 7 : (stage == 1 ? 19 : 27);
                        const float lim = RADIUS ? r2 : __uint_as_float(static_cast<unsigned>(thr >> 32));
                        const bool use = lane >= lo && lane < hi && (RADIUS ? b2 < lim : b2 <= lim);
                        scan_segments<S, RADIUS>(use ? len : 0, a, spts, qx, qy, qz, r2, k, lane, top, thr, hits, evals);
                    }
                }
                for (int r = 1;; r++) {
                    if (r > 1) {
                        const int side = 2 * r - 1;
                        const int nseg = 8 * r + 2 * side * side;
                        for (int sb = 0; sb < nseg; sb += 32) {
                            const int s = sb + lane;
                            uint32_t start = 0;
                            int len = 0;
                            if (s < nseg) {
                                int dy, dz, x0, x1;
                                if (s < 8 * r) {              // perimeter rows of the (2r+1)^2 square: full x extent
                                    const int sd = s / (2 * r), o = s % (2 * r);
                                    if (sd == 0) { dy = -r + o; dz = -r; }
                                    else if (sd == 1) { dy = r; dz = -r + o; }
                                    else if (sd == 2) { dy = r - o; dz = r; }
                                    else { dy = -r; dz = r - o; }
                                    x0 = cx - r; x1 = cx + r;
                                } else {                      // interior rows: only the two end cells are new
                                    const int t = s - 8 * r, w = t >> 1;
                                    dy = w % side - (r - 1);
                                    dz = w / side - (r - 1);
                                    x0 = x1 = (t & 1) ? cx + r : cx - r;
                                }
                                const int yy = cy + dy, zz = cz + dz;
                                // lower bound of the segment: |d| - 1 whole cells plus the gap inside the own cell
                                const float sy = dy < 0 ? loy + static_cast<float>(-dy - 1) * hs
                                                        : (dy > 0 ? hiy + static_cast<float>(dy - 1) * hs : 0.f);
                                const float sz = dz < 0 ? loz + static_cast<float>(-dz - 1) * hs
                                                        : (dz > 0 ? hiz + static_cast<float>(dz - 1) * hs : 0.f);
                                const float sx = x0 != x1 ? 0.f
                                                          : (x0 < cx ? lox + static_cast<float>(r - 1) * hs
                                                                     : hix + static_cast<float>(r - 1) * hs);
                                const float lim = RADIUS ? r2 : __uint_as_float(static_cast<unsigned>(thr >> 32));
                                const float b2 = sx * sx + sy * sy + sz * sz;
                                const bool reach = RADIUS ? b2 < lim : b2 <= lim;
                                if (reach && yy >= 0 && yy < g.ny && zz >= 0 && zz < g.nz) {
                                    x0 = max(x0, 0);
                                    x1 = min(x1, g.nx - 1);
                                    if (x0 <= x1) {
                                        const int64_t c0 = g.base + x0 + g.nx * (yy + g.ny * zz);
                                        start = __ldg(cell_start + c0);
                                        len = static_cast<int>(__ldg(cell_start + c0 + (x1 - x0) + 1) - start);
                                    }
                                }
                            }
                            scan_segments<S, RADIUS>(len, start, spts, qx, qy, qz, r2, k, lane, top, thr, hits, evals);
                        }
                    }
                    // every source of the tile seen?
                    const int bx0 = cx - r, bx1 = cx + r, by0 = cy - r, by1 = cy + r, bz0 = cz - r, bz1 = cz + r;
                    if (bx0 <= 0 && bx1 >= g.nx - 1 && by0 <= 0 && by1 >= g.ny - 1 && bz0 <= 0 && bz1 >= g.nz - 1) break;
                    // lower bound on the distance to anything outside the visited block
                    float bd = 3.0e38f;
                    if (bx0 > 0) bd = fminf(bd, qx - (g.ox + static_cast<float>(bx0) * g.h));
                    if (bx1 < g.nx - 1) bd = fminf(bd, (g.ox + static_cast<float>(bx1 + 1) * g.h) - qx);
                    if (by0 > 0) bd = fminf(bd, qy - (g.oy + static_cast<float>(by0) * g.h));
                    if (by1 < g.ny - 1) bd = fminf(bd, (g.oy + static_cast<float>(by1 + 1) * g.h) - qy);
                    if (bz0 > 0) bd = fminf(bd, qz - (g.oz + static_cast<float>(bz0) * g.h));
                    if (bz1 < g.nz - 1) bd = fminf(bd, (g.oz + static_cast<float>(bz1 + 1) * g.h) - qz);
                    const float bs = bd * (1.f - 1e-4f) - margin;
                    const float lim = RADIUS ? r2 : __uint_as_float(static_cast<unsigned>(thr >> 32));
                    if (bs > 0.f && lim < bs * bs) break;
                }
            }
#pragma unroll
            for (int s = 0; s < S; s++) {
                const int e = s * 32 + lane;
                if (e < k) {
                    nbr[q * k + e] = static_cast<int32_t>(static_cast<uint32_t>(top.key[s]));
                    if (!RADIUS && d2out) d2out[q * k + e] = __uint_as_float(static_cast<unsigned>(top.key[s] >> 32));
                }
            }
            if (RADIUS && lane == 0) cnt_out[q] = hits < k ? hits : k;
        }
    }
    if (pair_evals && lane == 0 && evals) atomicAdd(pair_evals, evals);
}

// ---- k <= 4 (the k = 2 search inside knn_interpolate, src/model.py:149): one THREAD per query.
// With ~one source per occupied cell the warp-per-query machinery (scan, flatten, ballots) costs far
// more than the handful of distances a query needs; here a thread walks the 27 cells of its first ring
// in the same nearest-first order, skips cells that cannot beat its current k-th distance, and keeps
// its k best keys in registers.  Same keys, same stopping bound, same results.

template <int K>
__device__ __forceinline__ void small_insert(key_t (&best)[K], key_t ck) {
    if (ck < best[K - 1]) {
        best[K - 1] = ck;
#pragma unroll
        for (int i = K - 1; i > 0; i--) {
            if (best[i] < best[i - 1]) {
                const key_t t = best[i];
                best[i] = best[i - 1];
                best[i - 1] = t;
            }
        }
    }
}

template <int K>
__device__ __forceinline__ void small_scan(const float4 *__restrict__ spts, uint32_t a, uint32_t e, float qx, float qy,
                                           float qz, key_t (&best)[K], unsigned &evals) {
    evals += e - a;
    for (uint32_t i = a; i < e; i++) {
        const float4 c = __ldg(spts + i);
        const float dx = __fsub_rn(c.x, qx), dy = __fsub_rn(c.y, qy), dz = __fsub_rn(c.z, qz);
        const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
        if (d < 1e10f) small_insert<K>(best, (static_cast<key_t>(__float_as_uint(d)) << 32) | __float_as_uint(c.w));
    }
}

template <int K>
__global__ void __launch_bounds__(256) grid_query_small_kernel(const float4 *__restrict__ spts,
                                                               const uint32_t *__restrict__ cell_start,
                                                               const GridTile *__restrict__ grid,
                                                               const float *__restrict__ y,
                                                               const int64_t *__restrict__ ptr_x,
                                                               const int64_t *__restrict__ ptr_y, int T, int64_t ny,
                                                               int32_t *__restrict__ nbr, float *__restrict__ d2out,
                                                               unsigned long long *__restrict__ pair_evals) {
    const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q >= ny) return;
    unsigned evals = 0;
    const int b = find_tile(ptr_y, T, q);
    const int64_t s0 = ptr_x[b], s1 = ptr_x[b + 1];
    const float qx = y[q * 3 + 0], qy = y[q * 3 + 1], qz = y[q * 3 + 2];
    key_t best[K];
#pragma unroll
    for (int i = 0; i < K; i++) best[i] = key_sentinel();
    if (s1 > s0) {
        const GridTile g = grid[b];
        const int cx = axis_cell(qx, g.ox, g.inv_h, g.nx);
        const int cy = axis_cell(qy, g.oy, g.inv_h, g.ny);
        const int cz = axis_cell(qz, g.oz, g.inv_h, g.nz);
        const int nmax = max(g.nx, max(g.ny, g.nz));
        const float margin = 1e-5f * g.h * static_cast<float>(nmax) +
                             4e-7f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + fabsf(g.ox) + fabsf(g.oy) + fabsf(g.oz));
        // safe gaps from the query to the six faces of its own cell
        const float fy = g.oy + static_cast<float>(cy) * g.h, fz = g.oz + static_cast<float>(cz) * g.h;
        const float k1 = 1.f - 1e-4f;
        const float loy = fmaxf((qy - fy) * k1 - margin, 0.f), hiy = fmaxf((fy + g.h - qy) * k1 - margin, 0.f);
        const float loz = fmaxf((qz - fz) * k1 - margin, 0.f), hiz = fmaxf((fz + g.h - qz) * k1 - margin, 0.f);
        // ring 1 as 9 rows of up to 3 cells (contiguous in the cell table), nearest rows first: the 18 table
        // entries are fetched up front (independent loads), then a row is scanned unless its distance to the
        // query already exceeds the current k-th distance
        uint32_t ra[9], re[9];
        float rb2[9];
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
#pragma unroll
        for (int s = 0; s < 9; s++) {
            const int ddy = static_cast<int>((0x22161u >> (2 * s)) & 3u) - 1;   // 0,-1,1,0,0,-1,1,-1,1
            const int ddz = static_cast<int>((0x28215u >> (2 * s)) & 3u) - 1;   // 0,0,0,-1,1,-1,-1,1,1
            const int yy = cy + ddy, zz = cz + ddz;
            ra[s] = re[s] = 0u;
            if (yy >= 0 && yy < g.ny && zz >= 0 && zz < g.nz) {
                const int64_t c0 = g.base + x0 + g.nx * (yy + g.ny * zz);
                ra[s] = __ldg(cell_start + c0);
                re[s] = __ldg(cell_start + c0 + (x1 - x0) + 1);
            }
            const float sy = ddy < 0 ? loy : (ddy > 0 ? hiy : 0.f);
            const float sz = ddz < 0 ? loz : (ddz > 0 ? hiz : 0.f);
            rb2[s] = sy * sy + sz * sz;
        }
#pragma unroll
        for (int s = 0; s < 9; s++) {
            if (rb2[s] > __uint_as_float(static_cast<unsigned>(best[K - 1] >> 32))) continue;
            small_scan<K>(spts, ra[s], re[s], qx, qy, qz, best, evals);
        }
        for (int r = 1;; r++) {
            if (r > 1) {                               // rare: shells beyond the first ring, row by row
                for (int dz = -r; dz <= r; dz++) {
                    const int zz = cz + dz;
                    if (zz < 0 || zz >= g.nz) continue;
                    for (int dy = -r; dy <= r; dy++) {
                        const int yy = cy + dy;
                        if (yy < 0 || yy >= g.ny) continue;
                        const bool rim = (dz == -r || dz == r || dy == -r || dy == r);
                        const int64_t row = g.base + g.nx * (yy + g.ny * zz);
                        if (rim) {
                            const int x0 = max(cx - r, 0), x1 = min(cx + r, g.nx - 1);
                            if (x0 <= x1)
                                small_scan<K>(spts, __ldg(cell_start + row + x0), __ldg(cell_start + row + x1 + 1), qx, qy,
                                              qz, best, evals);
                        } else {
                            if (cx - r >= 0)
                                small_scan<K>(spts, __ldg(cell_start + row + cx - r), __ldg(cell_start + row + cx - r + 1),
                                              qx, qy, qz, best, evals);
                            if (cx + r < g.nx)
                                small_scan<K>(spts, __ldg(cell_start + row + cx + r), __ldg(cell_start + row + cx + r + 1),
                                              qx, qy, qz, best, evals);
                        }
                    }
                }
            }
            const int bx0 = cx - r, bx1 = cx + r, by0 = cy - r, by1 = cy + r, bz0 = cz - r, bz1 = cz + r;
            if (bx0 <= 0 && bx1 >= g.nx - 1 && by0 <= 0 && by1 >= g.ny - 1 && bz0 <= 0 && bz1 >= g.nz - 1) break;
            float bd = 3.0e38f;
            if (bx0 > 0) bd = fminf(bd, qx - (g.ox + static_cast<float>(bx0) * g.h));
            if (bx1 < g.nx - 1) bd = fminf(bd, (g.ox + static_cast<float>(bx1 + 1) * g.h) - qx);
            if (by0 > 0) bd = fminf(bd, qy - (g.oy + static_cast<float>(by0) * g.h));
            if (by1 < g.ny - 1) bd = fminf(bd, (g.oy + static_cast<float>(by1 + 1) * g.h) - qy);
            if (bz0 > 0) bd = fminf(bd, qz - (g.oz + static_cast<float>(bz0) * g.h));
            if (bz1 < g.nz - 1) bd = fminf(bd, (g.oz + static_cast<float>(bz1 + 1) * g.h) - qz);
            const float bs = bd * (1.f - 1e-4f) - margin;
            if (bs > 0.f && __uint_as_float(static_cast<unsigned>(best[K - 1] >> 32)) < bs * bs) break;
        }
    }
#pragma unroll
    for (int i = 0; i < K; i++) {
        nbr[q * K + i] = static_cast<int32_t>(static_cast<uint32_t>(best[i]));
        if (d2out) d2out[q * K + i] = __uint_as_float(static_cast<unsigned>(best[i] >> 32));
    }
    if (pair_evals) {                       // the warp's active lanes add up, one atomic per warp
        const unsigned act = __activemask();
        const unsigned total = __reduce_add_sync(act, evals);
        if ((threadIdx.x & 31) == __ffs(act) - 1 && total) atomicAdd(pair_evals, static_cast<unsigned long long>(total));
    }
}

// ---- k >= 5: one THREAD per query, queries binned by cell, per-thread max-heap in shared memory.
// The warp-per-query kernel above spends its issue slots on selection: every candidate that beats the
// current k-th key costs a ballot / shuffle insertion that occupies all 32 lanes.  Here a thread owns a
// query and a binary max-heap of its k best keys (column `threadIdx.x` of a [k][HT + 1] shared array:
// conflict-free for the heap walk, and conflict-free again, transposed, when a warp writes the finished
// rows out coalesced).  A candidate costs its distance (3 FSUB, 1 FMUL, 2 FFMA) and one 64-bit compare with
// the heap root; survivors of a 32-candidate batch are remembered in a bit mask and inserted afterwards,
// so the warp walks the heap max-over-lanes(survivors) times per batch instead of once per surviving
// (lane, candidate) pair.  Queries are processed in cell order (a counting sort of the queries by the cell
// of the source grid they fall in, built with the sources' own histogram / scan), so the lanes of a warp
// read the same few rows of the cell list: L1 broadcasts instead of 32 scattered sectors.
// Same candidate sets, same keys, same stopping bound as above => bit-identical results.
constexpr int HT = 128;            // threads per CTA of the heap kernel

__global__ void __launch_bounds__(256) query_count_kernel(const float *__restrict__ y,
                                                          const int64_t *__restrict__ ptr_y, int T, int64_t ny,
                                                          const GridTile *__restrict__ grid,
                                                          uint32_t *__restrict__ qslot, uint32_t *__restrict__ qcount) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= ny) return;
    const int b = find_tile(ptr_y, T, i);
    const GridTile g = grid[b];
    const int cx = axis_cell(y[i * 3 + 0], g.ox, g.inv_h, g.nx);
    const int cy = axis_cell(y[i * 3 + 1], g.oy, g.inv_h, g.ny);
    const int cz = axis_cell(y[i * 3 + 2], g.oz, g.inv_h, g.nz);
    const int64_t c = g.base + cx + g.nx * (cy + g.ny * cz);
    qslot[i] = static_cast<uint32_t>(c);
    atomicAdd(&qcount[c], 1u);
}

// qstart = exclusive scan continued from the sources' table, i.e. offset by the number of sources
__global__ void __launch_bounds__(256) query_scatter_kernel(int64_t ny, const uint32_t *__restrict__ qslot,
                                                            const uint32_t *__restrict__ qstart, uint32_t nx_total,
                                                            uint32_t *__restrict__ qfill, uint32_t *__restrict__ qorder) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= ny) return;
    const uint32_t c = qslot[i];
    qorder[qstart[c] - nx_total + atomicAdd(&qfill[c], 1u)] = static_cast<uint32_t>(i);
}

// sift `nk` down from slot i of the max-heap h[0..n) (stride st)
__device__ __forceinline__ void heap_sift(key_t *__restrict__ h, int st, int n, int i, key_t nk) {
    for (;;) {
        int c = 2 * i + 1;
        if (c >= n) break;
        key_t kc = h[c * st];
        if (c + 1 < n) {
            const key_t kr = h[(c + 1) * st];
            if (kr > kc) { kc = kr; c++; }
        }
        if (!(kc > nk)) break;
        h[i * st] = kc;
        i = c;
    }
    h[i * st] = nk;
}

// Selection state of one query.  While fewer than k candidates have been seen they are appended unsorted
// (`fill` of them so far; the other slots hold sentinels, the largest key); the k-th one triggers Floyd's
// heap construction and from then on `root` is the k-th best key and a better candidate replaces the root.
struct HeapState {
    key_t root;
    int fill;
};

__device__ __forceinline__ void heap_offer(key_t *__restrict__ h, int st, int k, HeapState &hs, key_t ck) {
    if (hs.fill < k) {
        h[hs.fill * st] = ck;
        if (++hs.fill == k) {
            for (int i = k / 2 - 1; i >= 0; i--) heap_sift(h, st, k, i, h[i * st]);
            hs.root = h[0];
        }
    } else if (ck < hs.root) {
        heap_sift(h, st, k, 0, ck);
        hs.root = h[0];
    }
}

constexpr int HB = 4;      // candidates whose loads are in flight together

// ORDERED: rows ascending by key (the contract of p2w_knn / torch_cluster).  Otherwise the finished heap is
// written as it stands: entry 0 is the k-th (farthest) neighbour, the rest in no particular order (what the
// spatial vote needs: it sorts the probabilities itself) -- this skips k root extractions per query.
// The kernel is ONE loop nest (rings > rows > row parts > candidate batches) with a single copy of the scan
// body: the lanes of a warp diverge on trip counts, and a body replicated per ring-1 row thrashed the
// instruction cache (ncu: "no instruction" was the top stall).
template <bool RADIUS, bool ORDERED>
__global__ void __launch_bounds__(HT) grid_query_heap_kernel(const float4 *__restrict__ spts,
                                                             const uint32_t *__restrict__ cell_start,
                                                             const GridTile *__restrict__ grid,
                                                             const float *__restrict__ y,
                                                             const int64_t *__restrict__ ptr_x,
                                                             const int64_t *__restrict__ ptr_y, int T, int64_t ny,
                                                             int k, float r2, const uint32_t *__restrict__ qorder,
                                                             int32_t *__restrict__ nbr, float *__restrict__ d2out,
                                                             int32_t *__restrict__ cnt_out,
                                                             unsigned long long *__restrict__ pair_evals) {
    extern __shared__ __align__(16) unsigned char heap_raw[];
    key_t *heap_s = reinterpret_cast<key_t *>(heap_raw);
    constexpr int ST = HT + 1;
    key_t *h = heap_s + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int64_t p = static_cast<int64_t>(blockIdx.x) * HT + threadIdx.x;
    const bool active = p < ny;
    const int64_t q = active ? static_cast<int64_t>(qorder[p]) : 0;
    for (int e = 0; e < k; e++) h[e * ST] = key_sentinel();
    HeapState hs;
    hs.root = key_sentinel();
    hs.fill = 0;
    int hits = 0;
    unsigned evals = 0;
    if (active) {
        const int b = find_tile(ptr_y, T, q);
        const int64_t s0 = ptr_x[b], s1 = ptr_x[b + 1];
        const float qx = y[q * 3 + 0], qy = y[q * 3 + 1], qz = y[q * 3 + 2];
        if (s1 > s0) {
            const GridTile g = grid[b];
            const int cx = axis_cell(qx, g.ox, g.inv_h, g.nx);
            const int cy = axis_cell(qy, g.oy, g.inv_h, g.ny);
            const int cz = axis_cell(qz, g.oz, g.inv_h, g.nz);
            const int nmax = max(g.nx, max(g.ny, g.nz));
            const float margin = 1e-5f * g.h * static_cast<float>(nmax) +
                                 4e-7f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + fabsf(g.ox) + fabsf(g.oy) + fabsf(g.oz));
            const float fx = g.ox + static_cast<float>(cx) * g.h, fy = g.oy + static_cast<float>(cy) * g.h,
                        fz = g.oz + static_cast<float>(cz) * g.h;
            const float k1 = 1.f - 1e-4f;
            const float lox = fmaxf((qx - fx) * k1 - margin, 0.f), hix = fmaxf((fx + g.h - qx) * k1 - margin, 0.f);
            const float loy = fmaxf((qy - fy) * k1 - margin, 0.f), hiy = fmaxf((fy + g.h - qy) * k1 - margin, 0.f);
            const float loz = fmaxf((qz - fz) * k1 - margin, 0.f), hiz = fmaxf((fz + g.h - qz) * k1 - margin, 0.f);
            const float hs1 = g.h * k1;             // a further whole cell in between adds at least this much
#pragma unroll 1
            for (int r = 1;; r++) {
                // shell r as rows along x.  Ring 1: its 9 rows, nearest first, each the cells cx-1..cx+1.  Shell r > 1:
                // (2r+1)^2 rows; a rim row is new over cx-r..cx+r, an interior row only in its two end cells.
                const int side = 2 * r + 1, nrows = side * side;
#pragma unroll 1
                for (int s = 0; s < nrows; s++) {
                    int dy, dz;
                    if (r == 1) {
                        dy = static_cast<int>((0x22161u >> (2 * s)) & 3u) - 1;   // 0,-1,1,0,0,-1,1,-1,1
                        dz = static_cast<int>((0x28215u >> (2 * s)) & 3u) - 1;   // 0,0,0,-1,1,-1,-1,1,1
                    } else {
                        dy = s % side - r;
                        dz = s / side - r;
                    }
                    const int yy = cy + dy, zz = cz + dz;
                    if (yy < 0 || yy >= g.ny || zz < 0 || zz >= g.nz) continue;
                    const float sy = dy < 0 ? loy + static_cast<float>(-dy - 1) * hs1
                                            : (dy > 0 ? hiy + static_cast<float>(dy - 1) * hs1 : 0.f);
                    const float sz = dz < 0 ? loz + static_cast<float>(-dz - 1) * hs1
                                            : (dz > 0 ? hiz + static_cast<float>(dz - 1) * hs1 : 0.f);
                    const float rb2 = sy * sy + sz * sz;
                    const float lim = RADIUS ? r2 : __uint_as_float(static_cast<unsigned>(hs.root >> 32));
                    if (RADIUS ? !(rb2 < lim) : rb2 > lim) continue;
                    // the end cells are r - 1 whole cells plus the gap inside the own cell away along x
                    const float sxl = lox + static_cast<float>(r - 1) * hs1, sxh = hix + static_cast<float>(r - 1) * hs1;
                    const bool reach_lo = RADIUS ? rb2 + sxl * sxl < lim : rb2 + sxl * sxl <= lim;
                    const bool reach_hi = RADIUS ? rb2 + sxh * sxh < lim : rb2 + sxh * sxh <= lim;
                    const uint32_t *row = cell_start + (g.base + g.nx * (yy + g.ny * zz));
                    const bool rim = r == 1 || dz == -r || dz == r || dy == -r || dy == r;
                    uint32_t a0 = 0, e0 = 0, a1 = 0, e1 = 0;
                    if (rim) {
                        int x0 = max(cx - r, 0), x1 = min(cx + r, g.nx - 1);
                        if (r == 1) {                      // ring 1 drops an end cell that cannot hold a result
                            if (x0 < cx && !reach_lo) x0 = cx;
                            if (x1 > cx && !reach_hi) x1 = cx;
                        }
                        a0 = __ldg(row + x0);
                        e0 = __ldg(row + x1 + 1);
                    } else {
                        if (cx - r >= 0 && reach_lo) { a0 = __ldg(row + cx - r); e0 = __ldg(row + cx - r + 1); }
                        if (cx + r < g.nx && reach_hi) { a1 = __ldg(row + cx + r); e1 = __ldg(row + cx + r + 1); }
                    }
#pragma unroll 1
                    for (int part = 0; part < 2; part++) {
                        const uint32_t a = part ? a1 : a0, e = part ? e1 : e0;
#pragma unroll 1
                        for (uint32_t base = a; base < e; base += 32) {
                            const uint32_t n = min(e - base, 32u);
                            unsigned mask = 0;
#pragma unroll 1
                            for (uint32_t j0 = 0; j0 < n; j0 += HB) {
                                float4 c[HB];
#pragma unroll
                                for (int u = 0; u < HB; u++) c[u] = __ldg(spts + base + min(j0 + u, n - 1));
#pragma unroll
                                for (int u = 0; u < HB; u++) {
                                    const float dx = __fsub_rn(c[u].x, qx), dy2 = __fsub_rn(c[u].y, qy), dz2 = __fsub_rn(c[u].z, qz);
                                    const float d = __fmaf_rn(dz2, dz2, __fmaf_rn(dy2, dy2, __fmul_rn(dx, dx)));
                                    bool ok = j0 + u < n;
                                    key_t ck;
                                    if (RADIUS) {
                                        ok = ok && d < r2;
                                        hits += ok ? 1 : 0;
                                        ck = __float_as_uint(c[u].w);
                                    } else {
                                        ok = ok && d < 1e10f;
                                        ck = (static_cast<key_t>(__float_as_uint(d)) << 32) | __float_as_uint(c[u].w);
                                    }
                                    if (ok && ck < hs.root) mask |= 1u << (j0 + u);
                                }
                            }
                            evals += n;
#pragma unroll 1
                            while (mask) {
                                const uint32_t j = static_cast<uint32_t>(__ffs(mask) - 1);
                                mask &= mask - 1;
                                const float4 c = __ldg(spts + base + j);
                                key_t ck;
                                if (RADIUS) {
                                    ck = __float_as_uint(c.w);
                                } else {
                                    const float dx = __fsub_rn(c.x, qx), dy2 = __fsub_rn(c.y, qy), dz2 = __fsub_rn(c.z, qz);
                                    const float d = __fmaf_rn(dz2, dz2, __fmaf_rn(dy2, dy2, __fmul_rn(dx, dx)));
                                    ck = (static_cast<key_t>(__float_as_uint(d)) << 32) | __float_as_uint(c.w);
                                }
                                heap_offer(h, ST, k, hs, ck);
                            }
                        }
                    }
                }
                const int bx0 = cx - r, bx1 = cx + r, by0 = cy - r, by1 = cy + r, bz0 = cz - r, bz1 = cz + r;
                if (bx0 <= 0 && bx1 >= g.nx - 1 && by0 <= 0 && by1 >= g.ny - 1 && bz0 <= 0 && bz1 >= g.nz - 1) break;
                float bd = 3.0e38f;
                if (bx0 > 0) bd = fminf(bd, qx - (g.ox + static_cast<float>(bx0) * g.h));
                if (bx1 < g.nx - 1) bd = fminf(bd, (g.ox + static_cast<float>(bx1 + 1) * g.h) - qx);
                if (by0 > 0) bd = fminf(bd, qy - (g.oy + static_cast<float>(by0) * g.h));
                if (by1 < g.ny - 1) bd = fminf(bd, (g.oy + static_cast<float>(by1 + 1) * g.h) - qy);
                if (bz0 > 0) bd = fminf(bd, qz - (g.oz + static_cast<float>(bz0) * g.h));
                if (bz1 < g.nz - 1) bd = fminf(bd, (g.oz + static_cast<float>(bz1 + 1) * g.h) - qz);
                const float bs = bd * (1.f - 1e-4f) - margin;
                const float lim = RADIUS ? r2 : __uint_as_float(static_cast<unsigned>(hs.root >> 32));
                if (bs > 0.f && lim < bs * bs) break;
            }
        }
        // fewer than k candidates in reach: the unsorted prefix (+ sentinels) still has to become a heap
        if (hs.fill < k)
            for (int i = k / 2 - 1; i >= 0; i--) heap_sift(h, ST, k, i, h[i * ST]);
        if (ORDERED) {
            // heap sort in place: h[0..k) ascending (the unfilled entries are sentinels, the largest keys)
            for (int n = k - 1; n >= 1; n--) {
                const key_t last = h[n * ST];
                h[n * ST] = h[0];
                heap_sift(h, ST, n, 0, last);
            }
        }
    }
    // a warp writes its 32 rows coalesced: lane e reads entry e of query t's column (transposed, conflict-free)
    __syncwarp();
    const int wbase = threadIdx.x & ~31;
    for (int t = 0; t < 32; t++) {
        const long long qq = __shfl_sync(FULL, active ? static_cast<long long>(q) : -1ll, t);
        if (qq < 0) continue;
        for (int e = lane; e < k; e += 32) {
            const key_t v = heap_s[e * ST + wbase + t];
            nbr[qq * k + e] = static_cast<int32_t>(static_cast<uint32_t>(v));
            if (!RADIUS && d2out) d2out[qq * k + e] = __uint_as_float(static_cast<unsigned>(v >> 32));
        }
    }
    if (RADIUS && active) cnt_out[q] = hits < k ? hits : k;
    if (pair_evals) {
        unsigned long long ev = evals;
        for (int o = 16; o; o >>= 1) ev += __shfl_xor_sync(FULL, ev, o);
        if (lane == 0 && ev) atomicAdd(pair_evals, ev);
    }
}

struct GridWs {
    GridTile *grid;
    uint32_t *slot, *cell_start, *fill, *bsum, *qslot, *qorder;
    float *box;
    unsigned long long *pair_evals;
    float4 *spts;
    int64_t table;     // entries of one cell table (bound on the total number of cells, + 1); cell_start / fill hold
                       // TWO tables back to back: sources, then queries (scanned as one array)
    size_t total;
};

inline size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }
inline int64_t scan32_blocks(int64_t n) { return (n + SC_TILE - 1) / SC_TILE; }

inline GridWs grid_ws(void *base, int64_t nx, int64_t ny, int T) {
    GridWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void *p = base ? static_cast<unsigned char *>(base) + off : nullptr;
        off += align_up(bytes);
        return p;
    };
    w.table = 4 * nx + 64 * static_cast<int64_t>(T) + 1;       // sum over tiles of max(4 n_b, 64), + the end slot
    w.grid = static_cast<GridTile *>(take(sizeof(GridTile) * static_cast<size_t>(T)));
    w.box = static_cast<float *>(take(64));
    w.pair_evals = static_cast<unsigned long long *>(take(64));
    w.slot = static_cast<uint32_t *>(take(4 * static_cast<size_t>(nx)));
    w.qslot = static_cast<uint32_t *>(take(4 * static_cast<size_t>(ny)));
    w.qorder = static_cast<uint32_t *>(take(4 * static_cast<size_t>(ny)));
    w.spts = static_cast<float4 *>(take(16 * static_cast<size_t>(nx)));
    w.cell_start = static_cast<uint32_t *>(take(4 * static_cast<size_t>(2 * w.table)));
    w.fill = static_cast<uint32_t *>(take(4 * static_cast<size_t>(2 * w.table)));
    w.bsum = static_cast<uint32_t *>(take(4 * static_cast<size_t>(scan32_blocks(2 * w.table) + 2)));
    w.total = off;
    return w;
}

// P2W_KNN_WARP=1 keeps the warp-per-query kernel for k >= 5 (A/B measurements; results are identical)
inline bool use_warp_kernel() {
    static const bool v = [] { const char *e = getenv("P2W_KNN_WARP"); return e && e[0] == '1'; }();
    return v;
}

inline bool force_heap_kernel() {
    static const bool v = [] { const char *e = getenv("P2W_KNN_HEAP"); return e && e[0] == '1'; }();
    return v;
}

int grid_search(const float *x, const float *y, const int64_t *ptr_x, const int64_t *ptr_y, int T, int64_t nx,
                int64_t ny, int k, float r2, bool radius, float cell_hint, bool unordered, int32_t *nbr, float *d2,
                int32_t *cnt, void *ws, size_t ws_bytes, cudaStream_t st, const char *what) {
    P2W_REQUIRE(T >= 1 && nx >= 0 && ny >= 0, "%s: bad sizes", what);
    P2W_REQUIRE(nx < (int64_t(1) << 29) && ny < (int64_t(1) << 31) && nx + ny < (int64_t(1) << 32),
                "%s: nx must stay below 2^29 sources and nx + ny below 2^32 per call", what);
    P2W_REQUIRE(T < (1 << 24), "%s: too many tiles", what);
    if (ny == 0) return P2W_OK;
    const GridWs w = grid_ws(ws, nx, ny, T);
    P2W_REQUIRE(ws != nullptr && ws_bytes >= w.total && (reinterpret_cast<uintptr_t>(ws) & 15u) == 0,
                "%s: workspace too small or misaligned (%zu bytes needed)", what, w.total);
    const float tau = fmaxf(1.f, 0.45f * static_cast<float>(k));
    P2W_REQUIRE(cell_hint >= 0.f && cell_hint < 1e30f, "%s: bad cell size", what);
    const bool small = !radius && k <= 4;
    // Kernel choice, measured on B200 (profiles/r2_knn_ab.txt).  The per-thread heap wins the RADIUS search at the bench
    // shapes (SA1: 1.21 -> 0.89 ms) and every k <= 32 search on LARGE tiles (16 384 points, uniform or TLS-like: 1.4-2.1x,
    // BASELINE.json configs[2]); on the real SA2 / SA3 kNN (sub-sampled surfaces, ~750 sources per tile) its lanes
    // diverge on candidate counts and ring-2 walks and the warp-wide kernel stays ahead (0.83 vs 1.12 ms); at k = 64 (the
    // vote) two thirds of the ~200 candidates enter a 6-level heap and it loses as well.  So: the heap for the radius
    // search, and for kNN up to k = 32 when the tiles average 8 192 sources or more.  P2W_KNN_HEAP=1 / P2W_KNN_WARP=1
    // force one kernel for every k >= 5 (A/B runs; results are identical).
    const bool big_tiles = nx >= static_cast<int64_t>(8192) * T;
    const bool heap = !small && !use_warp_kernel() && ((k <= 32 && (radius || big_tiles)) || force_heap_kernel());
    if (!heap) unordered = false;                       // the other kernels always order (a valid answer to the flag)
    const float *pre_box = nullptr;
    if (T == 1 && cell_hint > 0.f && nx > 65536) {          // one plot-wide tile: the box is everybody's job
        P2W_LAUNCH(box_init_kernel, 1, 32, 0, st)(w.box);
        P2W_LAUNCH(box_kernel, kNumSMs * 8, 256, 0, st)(x, nx, w.box);
        pre_box = w.box;
    }
    P2W_LAUNCH(grid_plan_kernel, T, 512, 0, st)(x, ptr_x, tau, cell_hint, pre_box, w.grid);
    P2W_LAUNCH(grid_base_kernel, 1, 1024, 0, st)(w.grid, T);
    // cell_start and fill are adjacent: one memset clears both (each: source table, then query table)
    cudaMemsetAsync(w.cell_start, 0, reinterpret_cast<unsigned char *>(w.fill + 2 * w.table) -
                                         reinterpret_cast<unsigned char *>(w.cell_start), st);
    cudaMemsetAsync(w.pair_evals, 0, sizeof(unsigned long long), st);
    if (nx > 0) {
        const unsigned blocks = static_cast<unsigned>((nx + 255) / 256);
        P2W_LAUNCH(grid_count_kernel, blocks, 256, 0, st)(x, ptr_x, T, nx, w.grid, w.slot, w.cell_start);
        if (heap)
            P2W_LAUNCH(query_count_kernel, (unsigned)((ny + 255) / 256), 256, 0, st)(y, ptr_y, T, ny, w.grid, w.qslot,
                                                                                     w.cell_start + w.table);
        const int64_t entries = heap ? 2 * w.table : w.table;
        const int64_t nb = scan32_blocks(entries);
        P2W_LAUNCH(scan32_partial_kernel, (unsigned)nb, SC_T, 0, st)(w.cell_start, entries, w.bsum);
        P2W_LAUNCH(scan32_single_kernel, 1, 1024, 0, st)(w.bsum, nb);
        P2W_LAUNCH(scan32_apply_kernel, (unsigned)nb, SC_T, 0, st)(w.cell_start, entries, w.bsum, w.cell_start);
        P2W_LAUNCH(grid_scatter_kernel, blocks, 256, 0, st)(x, nx, w.slot, w.cell_start, w.fill, w.spts);
        if (heap)
            P2W_LAUNCH(query_scatter_kernel, (unsigned)((ny + 255) / 256), 256, 0, st)(
                ny, w.qslot, w.cell_start + w.table, static_cast<uint32_t>(nx), w.fill + w.table, w.qorder);
    }
    if (small) {                                          // thread per query, k best keys in registers
        const unsigned gs = static_cast<unsigned>((ny + 255) / 256);
#define P2W_GS(KK) P2W_LAUNCH((grid_query_small_kernel<KK>), gs, 256, 0, st)(w.spts, w.cell_start, w.grid, y, ptr_x, ptr_y, T, ny, nbr, d2, w.pair_evals)
        if (k == 1) P2W_GS(1); else if (k == 2) P2W_GS(2); else if (k == 3) P2W_GS(3); else P2W_GS(4);
#undef P2W_GS
        return check_launch(what);
    }
    if (heap && nx > 0) {                                 // thread per query, heap of k keys in shared memory
        const size_t smem = static_cast<size_t>(k) * (HT + 1) * sizeof(key_t);
        static bool attr_done = false;
        if (!attr_done) {
            cudaFuncSetAttribute(grid_query_heap_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(grid_query_heap_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(grid_query_heap_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            attr_done = true;
        }
        const unsigned gh = static_cast<unsigned>((ny + HT - 1) / HT);
#define P2W_GH(R, O) P2W_LAUNCH((grid_query_heap_kernel<R, O>), gh, HT, smem, st)(w.spts, w.cell_start, w.grid, y, ptr_x, ptr_y, T, ny, k, r2, w.qorder, nbr, d2, cnt, w.pair_evals)
        if (radius) P2W_GH(true, true); else if (unordered) P2W_GH(false, false); else P2W_GH(false, true);
#undef P2W_GH
        return check_launch(what);
    }
    int64_t blocks = ((ny + 31) / 32 + 7) / 8;          // a warp takes 32 queries at a time
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    const unsigned gb = static_cast<unsigned>(blocks);
#define P2W_GQ(S, R) P2W_LAUNCH((grid_query_kernel<S, R>), gb, 256, 0, st)(w.spts, w.cell_start, w.grid, y, ptr_x, ptr_y, T, ny, k, r2, nbr, d2, cnt, w.pair_evals)
    if (radius) {
        if (k <= 32) P2W_GQ(1, true); else if (k <= 64) P2W_GQ(2, true); else P2W_GQ(4, true);
    } else {
        if (k <= 32) P2W_GQ(1, false); else if (k <= 64) P2W_GQ(2, false); else P2W_GQ(4, false);
    }
#undef P2W_GQ
    return check_launch(what);
}

}  // namespace
}  // namespace p2w

using namespace p2w;

extern "C" size_t p2w_grid_search_ws_bytes(int64_t nx, int64_t ny, int32_t num_tiles) {
    return grid_ws(nullptr, nx < 0 ? 0 : nx, ny < 0 ? 0 : ny, num_tiles < 1 ? 1 : num_tiles).total;
}

extern "C" int p2w_grid_search_pair_evals(const void *ws, int64_t nx, int64_t ny, int32_t num_tiles,
                                          const unsigned long long **counter) {
    P2W_REQUIRE(ws != nullptr && counter != nullptr, "p2w_grid_search_pair_evals: null argument");
    *counter = grid_ws(const_cast<void *>(ws), nx, ny, num_tiles < 1 ? 1 : num_tiles).pair_evals;
    return P2W_OK;
}

extern "C" int p2w_knn_grid_ex(const float *x, const float *y, const int64_t *ptr_x, const int64_t *ptr_y,
                               int32_t num_tiles, int64_t nx, int64_t ny, int32_t k, float cell_size, int32_t flags,
                               int32_t *nbr, float *d2, void *ws, size_t ws_bytes, p2w_stream_t stream) {
    P2W_REQUIRE(k >= 1 && k <= P2W_MAX_K, "p2w_knn_grid: k=%d outside [1,%d]", k, P2W_MAX_K);
    const bool unordered = (flags & P2W_KNN_UNORDERED) != 0;
    return grid_search(x, y, ptr_x, ptr_y, num_tiles, nx, ny, k, 0.f, false, cell_size, unordered, nbr, d2, nullptr, ws,
                       ws_bytes, as_stream(stream), "p2w_knn_grid");
}

extern "C" int p2w_knn_grid(const float *x, const float *y, const int64_t *ptr_x, const int64_t *ptr_y,
                            int32_t num_tiles, int64_t nx, int64_t ny, int32_t k, int32_t *nbr, float *d2, void *ws,
                            size_t ws_bytes, p2w_stream_t stream) {
    return p2w_knn_grid_ex(x, y, ptr_x, ptr_y, num_tiles, nx, ny, k, 0.f, 0, nbr, d2, ws, ws_bytes, stream);
}

extern "C" int p2w_radius_grid(const float *x, const float *y, const int64_t *ptr_x, const int64_t *ptr_y,
                               int32_t num_tiles, int64_t nx, int64_t ny, double r, int32_t max_nbr, int32_t *nbr,
                               int32_t *cnt, void *ws, size_t ws_bytes, p2w_stream_t stream) {
    P2W_REQUIRE(max_nbr >= 1 && max_nbr <= P2W_MAX_K, "p2w_radius_grid: max_num_neighbors=%d outside [1,%d]", max_nbr,
                P2W_MAX_K);
    const float r2 = static_cast<float>(r * r);   // upstream passes r*r (double) into a float argument
    return grid_search(x, y, ptr_x, ptr_y, num_tiles, nx, ny, max_nbr, r2, true, 0.f, false, nbr, nullptr, cnt, ws, ws_bytes,
                       as_stream(stream), "p2w_radius_grid");
}
