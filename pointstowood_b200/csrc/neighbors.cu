// neighbors.cu -- K1 exact kNN, K2 radius search, K3 farthest point sampling.
//
// Semantics follow the CUDA kernels of torch_cluster that the reference calls through
// torch_geometric (call sites src/model.py:118,120,149):
// FP32 distance d = fma(dz,dz, fma(dy,dy, dx*dx)), neighbours ordered by (d, index),
// radius = the lowest-index sources with d < (float)(r*r).
//
// Design (B200): one warp carries QW queries through a sweep over its tile's sources.  The
// sources are staged through shared memory in 2048-point chunks with 1-D TMA bulk copies
// (cp.async.bulk -> UBLKCP) into a double buffer guarded by mbarriers, so every source is
// read from L2 once per CTA and the FP32 pipes never wait on global memory.  Each lane
// evaluates one candidate per step; the running top-k lives in registers, one (d, index)
// entry per lane and slot, sorted across the warp, and a candidate enters by a ballot +
// shuffle insertion only when it beats the current k-th entry lexicographically -- so the
// result is independent of the scan order and bit-identical to the serial insertion sort.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "topk.cuh"

namespace p2w {
namespace {

constexpr int CH = 2048;   // sources per staged chunk (multiple of 4 keeps 16-byte alignment)
constexpr int NW = 16;     // warps per CTA
constexpr unsigned FULL = 0xffffffffu;

struct SweepSmem {
    float buf[2][CH * 3];
    uint64_t full[2];
};

// One thread stages chunk g (global source rows [g*CH, g*CH+CH) clipped to nx) into buf.
__device__ __forceinline__ void issue_chunk(const float *__restrict__ x, int64_t nx, int64_t g, float *buf,
                                            uint64_t *bar) {
    const int64_t first = g * CH;
    int64_t npts = nx - first;
    if (npts > CH) npts = CH;
    const uint32_t bytes = static_cast<uint32_t>(npts) * 12u;
    const uint32_t bulk = bytes & ~15u;
    const float *src = x + first * 3;
    for (uint32_t t = bulk / 4; t < bytes / 4; t++) buf[t] = src[t];   // <= 3 trailing floats
    mbar_arrive_expect_tx(bar, bulk);
    if (bulk) bulk_g2s(buf, src, bulk, bar);
}

// S: top-k slots per lane (k <= 32*S).  RADIUS: first-k-by-index within r2 instead of kNN.
template <int S, int QW, bool RADIUS>
__global__ void __launch_bounds__(NW * 32)
    sweep_kernel(const float *__restrict__ x, const float *__restrict__ y, const int64_t *__restrict__ ptr_x,
                 const int64_t *__restrict__ ptr_y, int B, int64_t nx, int64_t ny, int k, float r2, int use_bulk,
                 int32_t *__restrict__ nbr, float *__restrict__ d2out, int32_t *__restrict__ cnt_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SweepSmem &sm = *reinterpret_cast<SweepSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int QB = NW * QW;
    const int64_t q_lo = static_cast<int64_t>(blockIdx.x) * QB;
    const int64_t q_hi = (q_lo + QB < ny) ? q_lo + QB : ny;

    if (threadIdx.x == 0) {
        mbar_init(&sm.full[0], 1);
        mbar_init(&sm.full[1], 1);
        fence_barrier_init();
    }
    __syncthreads();
    uint32_t phase[2] = {0, 0};

    // this warp's queries
    int64_t q[QW];
    int qtile[QW];
    float qx[QW], qy[QW], qz[QW];
    TopK<S> top[QW];
    float thr_d[QW];
    int thr_i[QW];
    int cnt[QW];
#pragma unroll
    for (int u = 0; u < QW; u++) {
        q[u] = q_lo + warp * QW + u;
        const bool ok = q[u] < q_hi;
        qtile[u] = ok ? find_tile(ptr_y, B, q[u]) : -1;
        qx[u] = ok ? y[q[u] * 3 + 0] : 0.f;
        qy[u] = ok ? y[q[u] * 3 + 1] : 0.f;
        qz[u] = ok ? y[q[u] * 3 + 2] : 0.f;
        top[u].init();
        thr_d[u] = 1e10f;
        thr_i[u] = -1;
        cnt[u] = 0;
    }

    const int b_lo = find_tile(ptr_y, B, q_lo);
    const int b_hi = find_tile(ptr_y, B, q_hi - 1);
    for (int b = b_lo; b <= b_hi; b++) {
        const int64_t x0 = ptr_x[b], x1 = ptr_x[b + 1];
        if (x1 <= x0) continue;
        bool mine[QW];
#pragma unroll
        for (int u = 0; u < QW; u++) mine[u] = (qtile[u] == b);
        const int64_t g0 = x0 / CH, g1 = (x1 - 1) / CH;
        const int nchunk = static_cast<int>(g1 - g0 + 1);
        if (use_bulk && threadIdx.x == 0) {
            issue_chunk(x, nx, g0, sm.buf[0], &sm.full[0]);
            if (nchunk > 1) issue_chunk(x, nx, g0 + 1, sm.buf[1], &sm.full[1]);
        }
        for (int c = 0; c < nchunk; c++) {
            const int64_t first = (g0 + c) * CH;
            const int bsel = c & 1;
            const float *buf = sm.buf[bsel];
            if (use_bulk) {
                mbar_wait(&sm.full[bsel], phase[bsel]);
                phase[bsel] ^= 1;
            } else {   // misaligned base pointer: cooperative copy, single buffer
                int64_t npts = nx - first;
                if (npts > CH) npts = CH;
                __syncthreads();
                for (int t = threadIdx.x; t < npts * 3; t += NW * 32) sm.buf[bsel][t] = x[first * 3 + t];
                __syncthreads();
            }
            // candidates of this chunk, as 32-bit offsets from the chunk start
            const int lo = static_cast<int>(((x0 > first) ? x0 : first) - first);
            const int hi = static_cast<int>(((x1 < first + CH) ? x1 : first + CH) - first);
            const int ibase = static_cast<int>(first);
            bool any_mine = false;
            float act_thr[QW];   // queries of other tiles (or already full) can never be hit
#pragma unroll
            for (int u = 0; u < QW; u++) {
                any_mine |= mine[u];
                act_thr[u] = mine[u] ? (RADIUS ? r2 : thr_d[u]) : -1.f;
            }
            if (any_mine) {
                for (int base = lo; base < hi; base += 32) {
                    const int jl = base + lane;
                    const bool valid = jl < hi;
                    const int o = (valid ? jl : lo) * 3;
                    const float cx = buf[o], cy = buf[o + 1], cz = buf[o + 2];
                    float d[QW];
                    bool anyhit = false;
#pragma unroll
                    for (int u = 0; u < QW; u++) {
                        const float dx = __fsub_rn(cx, qx[u]), dy = __fsub_rn(cy, qy[u]), dz = __fsub_rn(cz, qz[u]);
                        d[u] = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                        // superset of the exact test: radius d < r2, kNN (d, j) < (thr_d, thr_i)
                        anyhit |= RADIUS ? (d[u] < act_thr[u]) : (d[u] <= act_thr[u]);
                    }
                    if (!__any_sync(FULL, anyhit && valid)) continue;
                    const int ji = ibase + jl;
#pragma unroll
                    for (int u = 0; u < QW; u++) {
                        if (RADIUS) {
                            const bool hit = valid && d[u] < act_thr[u];
                            const unsigned m = __ballot_sync(FULL, hit);
                            if (m) {
                                const int slot = cnt[u] + __popc(m & ((1u << lane) - 1u));
                                if (hit && slot < k) nbr[q[u] * k + slot] = ji;
                                cnt[u] += __popc(m);
                                if (cnt[u] >= k) { cnt[u] = k; mine[u] = false; act_thr[u] = -1.f; }
                            }
                        } else {
                            unsigned m = __ballot_sync(FULL, valid && d[u] <= act_thr[u] &&
                                                                 key_less(d[u], ji, thr_d[u], thr_i[u]));
                            while (m) {
                                const int l = __ffs(m) - 1;
                                m &= m - 1;
                                const float cd = __shfl_sync(FULL, d[u], l);
                                const int ci = __shfl_sync(FULL, ji, l);
                                if (key_less(cd, ci, thr_d[u], thr_i[u])) {
                                    top[u].insert(cd, ci, lane);
                                    top[u].kth(k, thr_d[u], thr_i[u]);
                                    act_thr[u] = thr_d[u];
                                }
                            }
                        }
                    }
                }
            }
            if (RADIUS) {
                bool live = false;
#pragma unroll
                for (int u = 0; u < QW; u++) live |= mine[u];
                const int any_live = __syncthreads_or(live ? 1 : 0);
                if (!any_live) {      // every query of this CTA is full: drain the prefetch and stop
                    if (use_bulk && c + 1 < nchunk) {
                        mbar_wait(&sm.full[bsel ^ 1], phase[bsel ^ 1]);
                        phase[bsel ^ 1] ^= 1;
                    }
                    break;
                }
            } else {
                __syncthreads();
            }
            if (use_bulk && threadIdx.x == 0 && c + 2 < nchunk)
                issue_chunk(x, nx, g0 + c + 2, sm.buf[bsel], &sm.full[bsel]);
        }
        __syncthreads();   // tile boundary: both buffers idle before the next tile's prologue
    }

    // results
#pragma unroll
    for (int u = 0; u < QW; u++) {
        if (qtile[u] < 0) continue;
        if (RADIUS) {
            for (int e = cnt[u] + lane; e < k; e += 32) nbr[q[u] * k + e] = -1;
            if (lane == 0) cnt_out[q[u]] = cnt[u];
        } else {
#pragma unroll
            for (int s = 0; s < S; s++) {
                const int e = s * 32 + lane;
                if (e < k) {
                    nbr[q[u] * k + e] = top[u].i[s];
                    if (d2out) d2out[q[u] * k + e] = top[u].d[s];
                }
            }
        }
    }
}

template <int S, int QW, bool RADIUS>
int launch_sweep(const float *x, const float *y, const int64_t *ptr_x, const int64_t *ptr_y, int B, int64_t nx,
                 int64_t ny, int k, float r2, int32_t *nbr, float *d2, int32_t *cnt, cudaStream_t st) {
    auto kern = sweep_kernel<S, QW, RADIUS>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SweepSmem));
        attr_done = true;
    }
    const int use_bulk = ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) ? 1 : 0;
    const int64_t grid = (ny + NW * QW - 1) / (NW * QW);
    P2W_LAUNCH(kern, (unsigned)grid, NW * 32, sizeof(SweepSmem), st)(x, y, ptr_x, ptr_y, B, nx, ny, k, r2, use_bulk, nbr, d2, cnt);
    return check_launch(RADIUS ? "p2w_radius" : "p2w_knn");
}

// ------------------------------------------------------------------ table -> [2,E] edges
__global__ void table_count_kernel(const int32_t *__restrict__ nbr, int64_t ny, int k, int64_t *__restrict__ cnt) {
    const int64_t q = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= ny) return;
    int c = 0;
    for (int e = lane; e < k; e += 32) c += nbr[q * k + e] >= 0;
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
    if (lane == 0) cnt[q] = c;
}

// single-CTA exclusive scan (compatibility path only: ny <= a few 100k); out[n] = total
__global__ void __launch_bounds__(1024) scan_small_kernel(int64_t *__restrict__ a, int64_t n) {
    __shared__ int64_t wsum[32];
    __shared__ int64_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = i < n ? a[i] : 0;
        int64_t s = v;
        for (int o = 1; o < 32; o <<= 1) {
            int64_t t = __shfl_up_sync(FULL, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) wsum[warp] = s;
        __syncthreads();
        if (warp == 0) {
            int64_t w = wsum[lane];
            for (int o = 1; o < 32; o <<= 1) {
                int64_t t = __shfl_up_sync(FULL, w, o);
                if (lane >= o) w += t;
            }
            wsum[lane] = w;
        }
        __syncthreads();
        const int64_t carry = carry_s;
        const int64_t excl = carry + (warp ? wsum[warp - 1] : 0) + s - v;
        if (i < n) a[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wsum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) a[n] = carry_s;
}

__global__ void table_fill_kernel(const int32_t *__restrict__ nbr, int64_t ny, int k,
                                  const int64_t *__restrict__ off, int64_t E, int64_t *__restrict__ edges) {
    const int64_t q = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= ny) return;
    int64_t o = off[q];
    for (int e0 = 0; e0 < k; e0 += 32) {
        const int e = e0 + lane;
        const int v = e < k ? nbr[q * k + e] : -1;
        const unsigned m = __ballot_sync(FULL, v >= 0);
        if (v >= 0) {
            const int64_t p = o + __popc(m & ((1u << lane) - 1u));
            edges[p] = q;
            edges[E + p] = v;
        }
        o += __popc(m);
    }
}

// ------------------------------------------------------------------ K3 fps
// One 1024-thread CTA per tile.  Thread t owns points t, t+1024, ... ; up to PT of them are
// cached in registers together with their running min-distance (tiles <= 16384 points never
// touch memory after the prologue); longer tiles spill the remainder to dist_ws.
constexpr int FPS_T = 1024;
constexpr int FPS_PT = 16;

__global__ void __launch_bounds__(FPS_T) fps_kernel(const float *__restrict__ src, const int64_t *__restrict__ ptr,
                                                    const int64_t *__restrict__ out_ptr, float *__restrict__ dist_ws,
                                                    int64_t *__restrict__ out) {
    __shared__ float red_v[32];
    __shared__ int red_i[32];
    __shared__ float last[3];
    const int b = blockIdx.x;
    const int64_t s0 = ptr[b];
    const int n = static_cast<int>(ptr[b + 1] - s0);
    const int64_t o0 = out_ptr[b];
    const int m = static_cast<int>(out_ptr[b + 1] - o0);
    if (m <= 0 || n <= 0) return;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    float px[FPS_PT], py[FPS_PT], pz[FPS_PT], pd[FPS_PT];
#pragma unroll
    for (int u = 0; u < FPS_PT; u++) {
        const int i = t + u * FPS_T;
        const bool ok = i < n;
        px[u] = ok ? src[(s0 + i) * 3 + 0] : 0.f;
        py[u] = ok ? src[(s0 + i) * 3 + 1] : 0.f;
        pz[u] = ok ? src[(s0 + i) * 3 + 2] : 0.f;
        pd[u] = 5e4f;
    }
    for (int i = t + FPS_PT * FPS_T; i < n; i += FPS_T) dist_ws[s0 + i] = 5e4f;
    if (t == 0) {
        out[o0] = s0;
        last[0] = src[s0 * 3 + 0];
        last[1] = src[s0 * 3 + 1];
        last[2] = src[s0 * 3 + 2];
    }
    __syncthreads();
    for (int j = 1; j < m; j++) {
        const float lx = last[0], ly = last[1], lz = last[2];
        float bv = -1.f;
        int bi = 0x7fffffff;
#pragma unroll
        for (int u = 0; u < FPS_PT; u++) {
            const int i = t + u * FPS_T;
            if (i < n) {
                const float dx = __fsub_rn(px[u], lx), dy = __fsub_rn(py[u], ly), dz = __fsub_rn(pz[u], lz);
                const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                const float v = pd[u] < d ? pd[u] : d;
                pd[u] = v;
                if (v > bv) { bv = v; bi = i; }
            }
        }
        for (int i = t + FPS_PT * FPS_T; i < n; i += FPS_T) {
            const float dx = __fsub_rn(src[(s0 + i) * 3 + 0], lx), dy = __fsub_rn(src[(s0 + i) * 3 + 1], ly),
                        dz = __fsub_rn(src[(s0 + i) * 3 + 2], lz);
            const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
            const float old = dist_ws[s0 + i];
            const float v = old < d ? old : d;
            dist_ws[s0 + i] = v;
            if (v > bv) { bv = v; bi = i; }
        }
        // arg-max, lowest index on ties
        for (int o = 16; o; o >>= 1) {
            const float ov = __shfl_xor_sync(FULL, bv, o);
            const int oi = __shfl_xor_sync(FULL, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            bv = red_v[lane];
            bi = red_i[lane];
            for (int o = 16; o; o >>= 1) {
                const float ov = __shfl_xor_sync(FULL, bv, o);
                const int oi = __shfl_xor_sync(FULL, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) {
                if (bi == 0x7fffffff) bi = 0;
                out[o0 + j] = s0 + bi;
                last[0] = src[(s0 + bi) * 3 + 0];
                last[1] = src[(s0 + bi) * 3 + 1];
                last[2] = src[(s0 + bi) * 3 + 2];
            }
        }
        __syncthreads();
    }
}

// ---- K3 on a thread-block CLUSTER: eight CTAs (eight SMs) share one tile.  A round of farthest point sampling is
// a reduction over the whole tile followed by a broadcast of the winner, 4 096 times in a row for a 16 384-point tile
// at ratio 0.25, so the latency of ONE round is the whole cost.  With one 1 024-thread CTA per tile (above) a round is
// ~16 distance updates per thread, two block barriers and a dependent global load of the winner's coordinates
// (2.7 us); here a CTA of 256 threads keeps 8 points per thread in registers (2 048 points per CTA), reduces them to
// one candidate WITH its coordinates, and the eight candidates meet through distributed shared memory: every CTA
// reads the eight slots of the cluster (`map_shared_rank`), so all of them know the next reference point after ONE
// cluster barrier per round (slots are double-buffered by round parity).  Same arithmetic, same tie rule (lowest
// index), hence the same samples as the single-CTA kernel and the oracle.
constexpr int FPSC_T = 256, FPSC_PT = 8, FPSC_C = 8;
constexpr int FPSC_SPAN = FPSC_T * FPSC_C;             // points visited per register slot across the cluster

struct __align__(16) FpsSlot {
    float v;
    int i;
    float x, y, z;
    int pad[3];
};

__device__ __forceinline__ bool fps_better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__global__ void __cluster_dims__(FPSC_C, 1, 1) __launch_bounds__(FPSC_T)
    fps_cluster_kernel(const float *__restrict__ src, const int64_t *__restrict__ ptr, const int64_t *__restrict__ out_ptr,
                       float *__restrict__ dist_ws, int64_t *__restrict__ out) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ FpsSlot slot[2];
    __shared__ FpsSlot red[FPSC_T / 32];
    const int rank = static_cast<int>(cluster.block_rank());
    const int b = blockIdx.x / FPSC_C;
    const int64_t s0 = ptr[b];
    const int n = static_cast<int>(ptr[b + 1] - s0);
    const int64_t o0 = out_ptr[b];
    const int m = static_cast<int>(out_ptr[b + 1] - o0);
    if (m <= 0 || n <= 0) return;                      // the same for every CTA of the cluster
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int g = rank * FPSC_T + t;                   // 0 .. 2047: this thread's first point
    float px[FPSC_PT], py[FPSC_PT], pz[FPSC_PT], pd[FPSC_PT];
#pragma unroll
    for (int u = 0; u < FPSC_PT; u++) {
        const int i = g + u * FPSC_SPAN;
        const bool ok = i < n;
        px[u] = ok ? src[(s0 + i) * 3 + 0] : 0.f;
        py[u] = ok ? src[(s0 + i) * 3 + 1] : 0.f;
        pz[u] = ok ? src[(s0 + i) * 3 + 2] : 0.f;
        pd[u] = 5e4f;
    }
    for (int i = g + FPSC_PT * FPSC_SPAN; i < n; i += FPSC_SPAN) dist_ws[s0 + i] = 5e4f;
    float lx = src[s0 * 3 + 0], ly = src[s0 * 3 + 1], lz = src[s0 * 3 + 2];
    if (rank == 0 && t == 0) out[o0] = s0;
    for (int j = 1; j < m; j++) {
        float bv = -1.f, bx = 0.f, by = 0.f, bz = 0.f;
        int bi = 0x7fffffff;
#pragma unroll
        for (int u = 0; u < FPSC_PT; u++) {
            const int i = g + u * FPSC_SPAN;
            if (i < n) {
                const float dx = __fsub_rn(px[u], lx), dy = __fsub_rn(py[u], ly), dz = __fsub_rn(pz[u], lz);
                const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                const float v = pd[u] < d ? pd[u] : d;
                pd[u] = v;
                if (v > bv) { bv = v; bi = i; bx = px[u]; by = py[u]; bz = pz[u]; }      // i ascends with u: ties keep the lower
            }
        }
        for (int i = g + FPSC_PT * FPSC_SPAN; i < n; i += FPSC_SPAN) {               // tiles beyond 16 384 points
            const float qx = src[(s0 + i) * 3 + 0], qy = src[(s0 + i) * 3 + 1], qz = src[(s0 + i) * 3 + 2];
            const float dx = __fsub_rn(qx, lx), dy = __fsub_rn(qy, ly), dz = __fsub_rn(qz, lz);
            const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
            const float old = dist_ws[s0 + i];
            const float v = old < d ? old : d;
            dist_ws[s0 + i] = v;
            if (v > bv) { bv = v; bi = i; bx = qx; by = qy; bz = qz; }
        }
        for (int o = 16; o; o >>= 1) {
            const float ov = __shfl_xor_sync(FULL, bv, o);
            const int oi = __shfl_xor_sync(FULL, bi, o);
            const float ox = __shfl_xor_sync(FULL, bx, o), oy = __shfl_xor_sync(FULL, by, o), oz = __shfl_xor_sync(FULL, bz, o);
            if (fps_better(ov, oi, bv, bi)) { bv = ov; bi = oi; bx = ox; by = oy; bz = oz; }
        }
        if (lane == 0) { red[warp].v = bv; red[warp].i = bi; red[warp].x = bx; red[warp].y = by; red[warp].z = bz; }
        __syncthreads();
        if (warp == 0) {
            const FpsSlot c = red[lane & (FPSC_T / 32 - 1)];
            bv = c.v; bi = c.i; bx = c.x; by = c.y; bz = c.z;
            for (int o = FPSC_T / 64; o; o >>= 1) {
                const float ov = __shfl_xor_sync(FULL, bv, o);
                const int oi = __shfl_xor_sync(FULL, bi, o);
                const float ox = __shfl_xor_sync(FULL, bx, o), oy = __shfl_xor_sync(FULL, by, o), oz = __shfl_xor_sync(FULL, bz, o);
                if (fps_better(ov, oi, bv, bi)) { bv = ov; bi = oi; bx = ox; by = oy; bz = oz; }
            }
            if (lane == 0) {
                FpsSlot &d = slot[j & 1];
                d.v = bv; d.i = bi; d.x = bx; d.y = by; d.z = bz;
            }
        }
        cluster.sync();                                  // every CTA's candidate of this round is in its slot
        bv = -1.f;
        bi = 0x7fffffff;
#pragma unroll
        for (int r = 0; r < FPSC_C; r++) {
            const FpsSlot *p = cluster.map_shared_rank(&slot[j & 1], r);
            const float4 a = *reinterpret_cast<const float4 *>(p);               // v, i, x, y
            const float cz = p->z;
            const int ci = __float_as_int(a.y);
            if (fps_better(a.x, ci, bv, bi)) { bv = a.x; bi = ci; lx = a.z; ly = a.w; lz = cz; }
        }
        if (bi == 0x7fffffff) { bi = 0; lx = src[s0 * 3 + 0]; ly = src[s0 * 3 + 1]; lz = src[s0 * 3 + 2]; }
        if (rank == 0 && t == 0) out[o0 + j] = s0 + bi;
    }
    cluster.sync();                                      // no CTA leaves while its slots may still be read
}

}  // namespace
}  // namespace p2w

using namespace p2w;

extern "C" int p2w_knn(const float *x, const float *y, const int64_t *ptr_x, const int64_t *ptr_y, int32_t num_tiles,
                       int64_t nx, int64_t ny, int32_t k, int32_t *nbr, float *d2, p2w_stream_t stream) {
    P2W_REQUIRE(k >= 1 && k <= P2W_MAX_K, "p2w_knn: k=%d outside [1,%d]", k, P2W_MAX_K);
    P2W_REQUIRE(num_tiles >= 1 && nx >= 0 && ny >= 0, "p2w_knn: bad sizes");
    P2W_REQUIRE(nx < (int64_t(1) << 31), "p2w_knn: nx must fit int32 indices");
    if (ny == 0) return P2W_OK;
    cudaStream_t st = as_stream(stream);
    if (k <= 32) return launch_sweep<1, 4, false>(x, y, ptr_x, ptr_y, num_tiles, nx, ny, k, 0.f, nbr, d2, nullptr, st);
    if (k <= 64) return launch_sweep<2, 2, false>(x, y, ptr_x, ptr_y, num_tiles, nx, ny, k, 0.f, nbr, d2, nullptr, st);
    return launch_sweep<4, 2, false>(x, y, ptr_x, ptr_y, num_tiles, nx, ny, k, 0.f, nbr, d2, nullptr, st);
}

extern "C" int p2w_radius(const float *x, const float *y, const int64_t *ptr_x, const int64_t *ptr_y,
                          int32_t num_tiles, int64_t nx, int64_t ny, double r, int32_t max_nbr, int32_t *nbr,
                          int32_t *cnt, p2w_stream_t stream) {
    P2W_REQUIRE(max_nbr >= 1, "p2w_radius: max_num_neighbors=%d must be positive", max_nbr);
    P2W_REQUIRE(num_tiles >= 1 && nx >= 0 && ny >= 0, "p2w_radius: bad sizes");
    P2W_REQUIRE(nx < (int64_t(1) << 31), "p2w_radius: nx must fit int32 indices");
    if (ny == 0) return P2W_OK;
    const float r2 = static_cast<float>(r * r);   // upstream passes r*r (double) into a float argument
    return launch_sweep<1, 4, true>(x, y, ptr_x, ptr_y, num_tiles, nx, ny, max_nbr, r2, nbr, nullptr, cnt,
                                    as_stream(stream));
}

extern "C" int p2w_table_count(const int32_t *nbr, int64_t ny, int32_t k, int64_t *edge_offset, p2w_stream_t stream) {
    cudaStream_t st = as_stream(stream);
    if (ny > 0) P2W_LAUNCH(table_count_kernel, (unsigned)((ny * 32 + 255) / 256), 256, 0, st)(nbr, ny, k, edge_offset);
    P2W_LAUNCH(scan_small_kernel, 1, 1024, 0, st)(edge_offset, ny);
    return check_launch("p2w_table_count");
}

extern "C" int p2w_table_to_edges(const int32_t *nbr, int64_t ny, int32_t k, const int64_t *edge_offset,
                                  int64_t num_edges, int64_t *edges, p2w_stream_t stream) {
    if (ny == 0 || num_edges == 0) return P2W_OK;
    P2W_LAUNCH(table_fill_kernel, (unsigned)((ny * 32 + 255) / 256), 256, 0, as_stream(stream))(nbr, ny, k, edge_offset, num_edges, edges);
    return check_launch("p2w_table_to_edges");
}

extern "C" int p2w_fps(const float *src, const int64_t *ptr, const int64_t *out_ptr, int32_t num_tiles, int64_t n,
                       float *dist_ws, int64_t *out, p2w_stream_t stream) {
    P2W_REQUIRE(num_tiles >= 1, "p2w_fps: num_tiles must be positive");
    if (n == 0) return P2W_OK;
    // Few tiles: a cluster of eight CTAs per tile (B = 8 x 16 384 points, ratio 0.25: 11.3 -> 8.0 ms on B200).  Many
    // tiles fill the SMs with one CTA each, and eight CTAs per tile would queue in waves (B = 64: 11.3 vs 17.6 ms).
    // P2W_FPS_CLUSTER=0 / 1 forces one of them (A/B runs; the samples are identical).
    static const int forced = [] { const char *e = getenv("P2W_FPS_CLUSTER"); return e ? atoi(e) : -1; }();
    const bool single = forced >= 0 ? forced == 0 : num_tiles * FPSC_C > kNumSMs * 2;
    if (single)
        P2W_LAUNCH(fps_kernel, num_tiles, FPS_T, 0, as_stream(stream))(src, ptr, out_ptr, dist_ws, out);
    else
        P2W_LAUNCH(fps_cluster_kernel, num_tiles * FPSC_C, FPSC_T, 0, as_stream(stream))(src, ptr, out_ptr, dist_ws, out);
    return check_launch("p2w_fps");
}
