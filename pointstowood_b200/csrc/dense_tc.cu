// dense_tc.cu -- the EXPAND convolution of InvertedResidualBlock (src/model.py:46-85) with everything that follows it up to
// the next GEMM fused into its epilogue, on tcgen05 / TMEM:
//
//     out[n, co] = relu( relu( x[n, :] . W[co, :] + b[co] ) * a[co] + c[co] )            bf16 in, bf16 out, FP32 accumulate
//
// (k=1 conv + BatchNorm folded into W, b; ReLU; then the depthwise k=1 conv + BatchNorm of the following DepthwiseSeparable
// block folded into (a, c); ReLU).  The library path is a cuBLASLt GEMM with a bias + ReLU epilogue followed by a streaming
// pass over the [N, 4C] result (p2w_affine_relu): this GEMM is HBM-bound by its OUTPUT (K = C is small), so the extra pass
// costs as much as the GEMM itself (tools/bench_dense.py: 0.22 + 0.31 ms at N = 870 k, C = 128).  Here the [N, 4C] tensor
// is written once.
//
// Same machinery as conv_tc.cu, minus the gather: the contraction runs transposed (D^T[co, n] = sum_k W[co, k] x[n, k]):
// A = weights, pre-packed bf16 in the canonical no-swizzle K-major layout and streamed through a ring of 16 KB slices by 1-D
// TMA bulk copies (resident when all of them fit), B = a 128-row activation tile re-laid K-major by four loader warps with
// 16-byte loads / stores.  D^T lands in TMEM with one CHANNEL per lane, so bias, a, c are three registers of an epilogue
// thread.  The epilogue is the critical role (output-bound GEMM: 1 KB of results per row for 128 x 512 MACs at C = 128), hence
// SIXTEEN epilogue warps, four per TMEM lane quarter, each taking 32 of the 128 rows in two passes of 16: accumulator -> both
// stages -> bf16, exchanged pairwise with the neighbouring lane so 32-bit words go into the warp's own staging block
// [row][32 channels] -> read back with eight lanes per row, so a store instruction writes 8 rows x 64 contiguous bytes.
// History of the epilogue on the SA1 block (870 k rows, C = 128; library GEMM 0.22 ms + affine pass 0.31 ms): 2-byte stores
// from a channel-per-lane thread 0.71 ms; row-per-lane with 16-byte pieces stored straight from registers (32 lines per
// instruction) and constants by __ldg 0.85 ms, constants prefetched 0.51 ms; constants in shared memory + staged rows 0.32 ms
// with the LSU data pipe at 89 %; channel-per-lane + staging 0.32 ms (16-bit stores of two lanes into one word conflict);
// pairwise exchange 0.30 ms; sixteen epilogue warps 0.28 ms.  Roles per persistent CTA: warps 0-15 epilogue, 16-19 tile
// loader, 20 MMA issue (one elected lane), 21 weight producer.  Two 128-column accumulators alternate.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace p2w {
namespace {

constexpr int NT = 128;                        // rows per tile (MMA N)
constexpr int SLICE_K = 64;                    // k extent of one ring slice: four K = 16 MMAs
constexpr int SLICE_BYTES = 128 * SLICE_K * 2; // 16 KB: [8 k-chunks][128 rows][8 bf16]
constexpr int MAX_STAGES = 12;
constexpr int EPI_WARPS = 16, EPI_THREADS = EPI_WARPS * 32, LOAD_WARPS = 4, LOAD_THREADS = LOAD_WARPS * 32;
constexpr int THREADS = EPI_THREADS + LOAD_THREADS + 64;
constexpr int MAX_TILE_BUFS = 2;
constexpr int LU = 16;                         // 16-byte loads a loader thread keeps in flight
constexpr int LBO_B = NT * 16 + 16;            // k-chunk stride of the activation tile, padded against bank conflicts
constexpr int TMEM_COLS = 2 * NT;
constexpr int ROWS_PER_WARP = NT / (EPI_WARPS / 4);       // epilogue warps per TMEM lane quarter split the tile's rows
constexpr int MMA_WARP = EPI_WARPS + LOAD_WARPS, PRODUCER_WARP = MMA_WARP + 1;
constexpr unsigned FULL = 0xffffffffu;

struct DenseParams {
    const __nv_bfloat16 *x;        // [n, K] row-major
    __nv_bfloat16 *out;            // [n, Co] row-major
    const unsigned char *wpack;    // [NB][K/8][128][8] bf16
    const float *bias, *a, *c;     // [NB * 128] (padded); a == NULL: no second stage
    int64_t n;
    int K, Co, NB, num_tiles, stages, resident, tile_bufs;
};

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;       // descriptor version (Blackwell); layout type 0 = no swizzle
    return d;
}
// kind::f16: D = F32, A = B = BF16, both K-major, M = 128, N = NT
__host__ __device__ constexpr uint32_t instr_desc() {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(NT >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                     uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\n"
        "mov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

constexpr int STAGE_COLS = 32;                              // columns an epilogue warp transposes per pass
constexpr int STAGE_ROW = STAGE_COLS * 2;                   // bytes per staged row: 64, conflict-free both ways (see the epilogue)
constexpr int PASS_ROWS = 16;                               // rows an epilogue warp takes through its staging block at a time
constexpr int STAGE_BYTES = PASS_ROWS * STAGE_ROW;
struct SmemLayout {
    uint32_t ring, tiles, stage, consts, bars, tmem, total;
};
__host__ __device__ inline uint32_t tile_bytes(int K) { return ((K / 8) * LBO_B + 127u) & ~127u; }
__host__ __device__ inline SmemLayout smem_layout(int K, int NB, int stages, int tile_bufs) {
    SmemLayout L;
    L.ring = 0;
    L.tiles = L.ring + stages * SLICE_BYTES;
    L.stage = L.tiles + tile_bufs * tile_bytes(K);
    L.consts = L.stage + EPI_WARPS * STAGE_BYTES;
    L.bars = (L.consts + 3u * NB * 128u * sizeof(float) + 7u) & ~7u;
    L.tmem = L.bars + 8 * (2 * MAX_STAGES + 4 + 2 * MAX_TILE_BUFS);
    L.total = L.tmem + 16;
    return L;
}

__global__ void __launch_bounds__(THREADS, 1) dense_tc_kernel(const DenseParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const SmemLayout L = smem_layout(p.K, p.NB, p.stages, p.tile_bufs);
    const int TILE_BUFS = p.tile_bufs;
    const uint32_t tbytes = tile_bytes(p.K);
    const int STAGES = p.stages;
    unsigned char *ring = smem + L.ring;
    unsigned char *tiles = smem + L.tiles;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars);
    uint64_t *ring_full = bars, *ring_empty = bars + MAX_STAGES;
    uint64_t *acc_full = bars + 2 * MAX_STAGES, *acc_empty = acc_full + 2;
    uint64_t *b_full = acc_empty + 2, *b_empty = b_full + MAX_TILE_BUFS;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + L.tmem);
    const int warp = __shfl_sync(FULL, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&ring_full[s], 1); mbar_init(&ring_empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], EPI_THREADS); }
        for (int m = 0; m < MAX_TILE_BUFS; m++) { mbar_init(&b_full[m], LOAD_THREADS); mbar_init(&b_empty[m], 1); }
        fence_barrier_init();
    }
    {   // the per-channel constants of both stages: [bias | a | c], read by every epilogue thread for every tile
        float *sc = reinterpret_cast<float *>(smem + L.consts);
        const int nc = p.NB * 128;
        for (int i = threadIdx.x; i < nc; i += THREADS) {
            sc[i] = p.bias[i];
            sc[nc + i] = p.a ? p.a[i] : 1.f;
            sc[2 * nc + i] = p.c ? p.c[i] : 0.f;
        }
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const int my_tiles = (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int ns = p.K / SLICE_K;                      // ring slices per accumulator block

    if (warp == PRODUCER_WARP) {
        // ------------------------------------------------ weight producer (all slices once when they fit the ring)
        int slot = 0;
        uint32_t ph = 0;
        const int passes = p.resident ? (my_tiles > 0 ? 1 : 0) : my_tiles;
        for (int it = 0; it < passes; it++) {
            const unsigned char *src = p.wpack;
            for (int i = 0; i < p.NB * ns; i++) {
                mbar_wait(&ring_empty[slot], ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&ring_full[slot], SLICE_BYTES);
                    bulk_g2s(ring + slot * SLICE_BYTES, src, SLICE_BYTES, &ring_full[slot]);
                }
                src += SLICE_BYTES;
                if (++slot == STAGES) { slot = 0; ph ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == MMA_WARP) {
        // ------------------------------------------------ MMA issuer (A = weight slice, B = activation tile: D^T[co, n])
        int slot = 0, acc = 0, mbuf = 0;
        uint32_t ph = 0, mph = 0, use0 = 0, use1 = 0;
        const uint64_t a_desc0 = smem_desc(smem_u32(ring), 2048, 128);
        const uint64_t b_desc0 = smem_desc(smem_u32(tiles), LBO_B, 128);
        const uint32_t a_hi = static_cast<uint32_t>(a_desc0 >> 32), a_lo0 = static_cast<uint32_t>(a_desc0);
        const uint32_t b_hi = static_cast<uint32_t>(b_desc0 >> 32);
        constexpr uint32_t ID = instr_desc();
        constexpr uint32_t bq = 2u * (LBO_B >> 4);        // B descriptor step per K = 16
        const bool stream = !p.resident;
        for (int it = 0; it < my_tiles; it++) {
            mbar_wait(&b_full[mbuf], mph);
            tc_fence_after();
            const uint32_t b_lo0 = static_cast<uint32_t>(b_desc0) + static_cast<uint32_t>(mbuf) * (tbytes >> 4);
            for (int blk = 0; blk < p.NB; blk++) {
                mbar_wait(&acc_empty[acc], ((acc ? use1 : use0) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_addr = tmem_base + acc * NT;
                uint32_t bd = b_lo0;
                for (int s = 0; s < ns; s++) {
                    if (stream || it == 0) mbar_wait(&ring_full[slot], ph);
                    const uint32_t ad = a_lo0 + static_cast<uint32_t>(slot) * (SLICE_BYTES >> 4);
                    if (elect_one()) {
                        umma(d_addr, ad, a_hi, bd, b_hi, ID, s ? 1u : 0u);
                        umma(d_addr, ad + 256u, a_hi, bd + bq, b_hi, ID, 1u);
                        umma(d_addr, ad + 512u, a_hi, bd + 2u * bq, b_hi, ID, 1u);
                        umma(d_addr, ad + 768u, a_hi, bd + 3u * bq, b_hi, ID, 1u);
                        if (stream) umma_commit(&ring_empty[slot]);
                    }
                    bd += 4u * bq;
                    if (++slot == STAGES) { slot = 0; ph ^= 1; }
                }
                if (elect_one()) umma_commit(&acc_full[acc]);
                if (acc) use1++; else use0++;
                acc ^= 1;
            }
            if (elect_one()) umma_commit(&b_empty[mbuf]);
            if (++mbuf == TILE_BUFS) { mbuf = 0; mph ^= 1; }
        }
        __syncwarp();
    } else if (warp >= EPI_WARPS) {
        // ------------------------------------------------ tile loader: 128 rows x K bf16 -> K-major [K/8][128][8]
        uint32_t mph = 0;
        int mbuf = 0;
        const int lw = warp - EPI_WARPS;
        const int CPR = p.K >> 3;                          // 16-byte chunks per row
        const int cpr_c = CPR < 32 ? CPR : 32;
        const int rpw = 32 / cpr_c;                        // rows a warp covers per pass (CPR <= 32)
        for (int it = 0; it < my_tiles; it++) {
            const int64_t t0 = (static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(it) * gridDim.x) * NT;
            unsigned char *tile = tiles + mbuf * tbytes;
            mbar_wait(&b_empty[mbuf], mph ^ 1);
            for (int c0 = 0; c0 < CPR; c0 += 32) {         // K > 256: the row in several 512-byte pieces
                const int ck = c0 + lane % cpr_c, ri = lane / cpr_c;
                for (int r0 = lw * rpw; r0 < NT; r0 += LOAD_WARPS * rpw * LU) {
                    uint4 v[LU];
#pragma unroll
                    for (int u = 0; u < LU; u++) {
                        const int n = r0 + u * LOAD_WARPS * rpw + ri;
                        v[u] = make_uint4(0, 0, 0, 0);
                        if (n < NT && ck < CPR && t0 + n < p.n)
                            v[u] = __ldg(reinterpret_cast<const uint4 *>(p.x + (t0 + n) * p.K + ck * 8));
                    }
#pragma unroll
                    for (int u = 0; u < LU; u++) {
                        const int n = r0 + u * LOAD_WARPS * rpw + ri;
                        if (n < NT && ck < CPR) *reinterpret_cast<uint4 *>(tile + ck * LBO_B + n * 16) = v[u];
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(&b_full[mbuf]);
            if (++mbuf == TILE_BUFS) { mbuf = 0; mph ^= 1; }
        }
    } else {
        // ------------------------------------------------ epilogue warps: TMEM lane quarter q (channels), row half h
        int acc = 0;
        uint32_t use0 = 0, use1 = 0;
        const int q = warp & 3, h = warp >> 2;
        const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16);
        const bool two = p.a != nullptr;
        unsigned char *stage = smem + L.stage + warp * STAGE_BYTES;
        const float *sconst = reinterpret_cast<const float *>(smem + L.consts);
        const int nconst = p.NB * 128;
        for (int it = 0; it < my_tiles; it++) {
            const int64_t t0 = (static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(it) * gridDim.x) * NT;
            for (int blk = 0; blk < p.NB; blk++) {
                const int cw = blk * 128 + 32 * q;                     // first channel of this warp
                const float bias = sconst[cw + lane], a = sconst[nconst + cw + lane], c = sconst[2 * nconst + cw + lane];
                mbar_wait(&acc_full[acc], (acc ? use1 : use0) & 1);
                tc_fence_after();
                // 32 rows per warp in two passes of 16: a lane owns ONE channel (its constants are registers); the bf16
                // results go through the warp's own staging block [row][32 channels] and come back with eight lanes per
                // row, so a store instruction writes 8 rows x 64 contiguous bytes (whole sectors)
#pragma unroll 1
                for (int j = 0; j < ROWS_PER_WARP / PASS_ROWS; j++) {
                    const int r0 = ROWS_PER_WARP * h + PASS_ROWS * j;
                    uint32_t r[PASS_ROWS];
                    tmem_ld16(lane_taddr + acc * NT + r0, r);
                    // 2 x 2 exchange with the neighbouring lane: an even lane ends up with rows e, e + 2 of channels (L, L + 1), an
                    // odd one with rows e + 1, e + 3 of (L - 1, L): 32-bit stores, sixteen lanes per row and two adjacent rows
                    // per instruction = all 32 banks once (16-bit stores of two lanes into one word cost a second wavefront)
                    const bool odd = lane & 1;
                    unsigned char *sw = stage + (odd ? STAGE_ROW : 0) + (lane & ~1) * 2;
#pragma unroll
                    for (int e = 0; e < PASS_ROWS; e += 4) {
                        float v[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            v[u] = fmaxf(__uint_as_float(r[e + u]) + bias, 0.f);
                            if (two) v[u] = fmaxf(fmaf(v[u], a, c), 0.f);
                        }
                        __nv_bfloat162 t = __floats2bfloat162_rn(v[0], v[2]);
                        const uint32_t pa = *reinterpret_cast<uint32_t *>(&t);       // rows e, e + 2
                        t = __floats2bfloat162_rn(v[1], v[3]);
                        const uint32_t pb = *reinterpret_cast<uint32_t *>(&t);       // rows e + 1, e + 3
                        const uint32_t got = __shfl_xor_sync(FULL, odd ? pa : pb, 1);
                        const uint32_t x = odd ? got : pa, y = odd ? pb : got;       // (channel L&~1, channel L|1)
                        *reinterpret_cast<uint32_t *>(sw + e * STAGE_ROW) = __byte_perm(x, y, 0x5410);
                        *reinterpret_cast<uint32_t *>(sw + (e + 2) * STAGE_ROW) = __byte_perm(x, y, 0x7632);
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < PASS_ROWS / 8; i++) {
                        const int rr = i * 8 + (lane >> 2), c8 = (lane & 3) * 8;
                        const uint4 v = *reinterpret_cast<const uint4 *>(stage + rr * STAGE_ROW + c8 * 2);
                        const int64_t row = t0 + r0 + rr;
                        if (row < p.n && cw + c8 < p.Co)                             // Co % 8 == 0: a group of 8 is all in or all out
                            *reinterpret_cast<uint4 *>(p.out + row * p.Co + cw + c8) = v;
                    }
                    __syncwarp();
                }
                tc_fence_before();
                mbar_arrive(&acc_empty[acc]);
                if (acc) use1++; else use0++;
                acc ^= 1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// w [Co, K] fp32 row-major -> bf16 [NB][K/8][128][8] (rows beyond Co zero): every ring slice contiguous
__global__ void dense_pack_kernel(const float *__restrict__ w, int Co, int K, int NB, __nv_bfloat16 *__restrict__ out) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t total = static_cast<int64_t>(NB) * K * 128;
    if (idx >= total) return;
    const int e = static_cast<int>(idx & 7);
    const int r = static_cast<int>((idx >> 3) & 127);
    const int kc = static_cast<int>((idx >> 10) % (K >> 3));
    const int blk = static_cast<int>((idx >> 10) / (K >> 3));
    const int row = blk * 128 + r, k = kc * 8 + e;
    out[idx] = __float2bfloat16(row < Co ? w[static_cast<int64_t>(row) * K + k] : 0.f);
}
__global__ void dense_padvec_kernel(const float *__restrict__ v, int n, int np, float fill, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) out[i] = (v && i < n) ? v[i] : fill;
}

struct DensePlan {
    int NB;
    size_t w_bytes, off_bias, off_a, off_c, total;
};
inline DensePlan dense_plan(int K, int Co) {
    DensePlan t;
    t.NB = (Co + 127) / 128;
    t.w_bytes = static_cast<size_t>(t.NB) * K * 128 * 2;
    t.off_bias = t.w_bytes;
    t.off_a = t.off_bias + sizeof(float) * t.NB * 128;
    t.off_c = t.off_a + sizeof(float) * t.NB * 128;
    t.total = t.off_c + sizeof(float) * t.NB * 128 + 256;
    return t;
}

}  // namespace
}  // namespace p2w

using namespace p2w;

extern "C" size_t p2w_dense_expand_ws_bytes(int32_t k, int32_t c_out) { return dense_plan(k, c_out).total; }

extern "C" int p2w_dense_expand(const void *x, int64_t n, int32_t k, int32_t c_out, const float *w, const float *bias,
                                const float *a, const float *c, void *out, void *ws, size_t ws_bytes, int32_t flags,
                                p2w_stream_t stream) {
    P2W_REQUIRE(k >= 64 && k <= 512 && k % 64 == 0, "p2w_dense_expand: k=%d must be a multiple of 64 in [64, 512]", k);
    P2W_REQUIRE(c_out >= 8 && c_out % 8 == 0, "p2w_dense_expand: c_out=%d must be a positive multiple of 8", c_out);
    P2W_REQUIRE((a == nullptr) == (c == nullptr), "p2w_dense_expand: a and c go together");
    P2W_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0 &&
                    (reinterpret_cast<uintptr_t>(ws) & 127u) == 0,
                "p2w_dense_expand: rows must be 16-byte aligned, the workspace 128-byte aligned");
    const DensePlan t = dense_plan(k, c_out);
    P2W_REQUIRE(ws_bytes >= t.total, "p2w_dense_expand: workspace too small");
    if (n == 0) return P2W_OK;
    cudaStream_t st = as_stream(stream);
    unsigned char *base = static_cast<unsigned char *>(ws);
    float *bp = reinterpret_cast<float *>(base + t.off_bias), *ap = reinterpret_cast<float *>(base + t.off_a),
          *cp = reinterpret_cast<float *>(base + t.off_c);
    if (!(flags & 1)) {                                                    // bit 0: `ws` already holds the packed weights
        const int64_t nw = static_cast<int64_t>(t.NB) * k * 128;
        P2W_LAUNCH(dense_pack_kernel, (unsigned)((nw + 255) / 256), 256, 0, st)(w, c_out, k, t.NB, reinterpret_cast<__nv_bfloat16 *>(base));
        P2W_LAUNCH(dense_padvec_kernel, (t.NB * 128 + 255) / 256, 256, 0, st)(bias, c_out, t.NB * 128, 0.f, bp);
        P2W_LAUNCH(dense_padvec_kernel, (t.NB * 128 + 255) / 256, 256, 0, st)(a, c_out, t.NB * 128, 1.f, ap);
        P2W_LAUNCH(dense_padvec_kernel, (t.NB * 128 + 255) / 256, 256, 0, st)(c, c_out, t.NB * 128, 0.f, cp);
    }
    const int per_tile = t.NB * (k / SLICE_K);
    // Two activation tiles (the next one loads while this one multiplies) whenever a ring of three slices still fits beside
    // them: trading the second tile for a deeper ring measured slower (K = 256: 0.35 ms against 0.29 ms)
    int tile_bufs = 0, stages = 0;
    for (int tb = MAX_TILE_BUFS; tb >= 1 && !tile_bufs; tb--) {
        const unsigned fixed = smem_layout(k, t.NB, 0, tb).total;
        if (fixed + 3u * SLICE_BYTES > 227u * 1024u) continue;
        tile_bufs = tb;
        stages = static_cast<int>((227u * 1024u - fixed) / SLICE_BYTES);
    }
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    P2W_REQUIRE(tile_bufs >= 1 && stages >= 3, "p2w_dense_expand: k=%d leaves no room for the weight ring", k);
    const int resident = per_tile <= stages ? 1 : 0;
    if (resident) stages = per_tile;
    const SmemLayout L = smem_layout(k, t.NB, stages, tile_bufs);
    DenseParams p;
    p.x = static_cast<const __nv_bfloat16 *>(x);
    p.out = static_cast<__nv_bfloat16 *>(out);
    p.wpack = base;
    p.bias = bp;
    p.a = a ? ap : nullptr;
    p.c = a ? cp : nullptr;
    p.n = n;
    p.K = k; p.Co = c_out; p.NB = t.NB;
    p.num_tiles = static_cast<int>((n + NT - 1) / NT);
    p.stages = stages; p.resident = resident; p.tile_bufs = tile_bufs;
    static int sm_count = 0;
    static unsigned smem_set = 0;
    if (!sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (sm_count <= 0) sm_count = kNumSMs;
    }
    if (L.total > smem_set) {
        cudaFuncSetAttribute(dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
        smem_set = L.total;
    }
    int grid = sm_count;
    if (grid > p.num_tiles) grid = p.num_tiles;
    P2W_LAUNCH(dense_tc_kernel, grid, THREADS, L.total, st)(p);
    return check_launch("p2w_dense_expand");
}
