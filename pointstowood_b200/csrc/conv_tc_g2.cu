// conv_tc_g2.cu -- the H <= 64 (SA1) instantiation of conv_tc.cu: TWO gather groups, register reallocation and
// pipelined TMEM loads in epilogue 2 (see the comments at the role split); everything else as in
// conv_tc.cu -- K5 on the 5th-generation tensor cores: fused gather -> per-edge MLP -> max with
// BF16 operands, FP32 accumulation in TMEM (tcgen05.mma, cta_group::1, M=128 x N=128 x K=16).
//
// Replaces MessagePassing.propagate(aggr='max') around PointNetConv.message
// (src/pointnet.py:108,116-132).  The contraction is run TRANSPOSED so that the max over a
// target's 32 edges never crosses threads:
//
//     D1^T[h, e] = sum_k W1[h, k] * msg[e, k]        A = W1 (K-major), B = msg tile (K-major)
//     D2^T[c, e] = sum_h W2[c, h] * hid[e, h]        A = W2 (K-major), B = hid tile (MN-major)
//
// An accumulator block is 128 channels (TMEM lanes) x 128 edges (TMEM columns = 4 targets x 32
// edges).  An epilogue thread owns one channel: bias / ReLU / BatchNorm are per-thread constants
// and the segment max is a 32-long register reduction.  Per CTA (persistent), four roles run as a
// pipeline over the edge tiles:
//   warps 0-3  epilogue: epilogue 1 (TMEM -> ReLU -> bf16 hid tile in shared memory) and
//              epilogue 2 (TMEM -> ReLU -> BN -> max -> out[t, c], coalesced over c);
//   warps 4-7  gather: neighbour rows + geometry of the NEXT tile into the msg tile (bf16, canonical
//              no-swizzle K-major layout) as soon as layer 1 of the current tile has consumed it;
//   warp 8     one thread issues tcgen05.mma and tcgen05.commit;
//   warp 9     one thread streams the pre-packed bf16 weights through a 4-stage ring of 8 KB
//              slices with 1-D TMA bulk copies (the weights stay L2 resident).
// Two 128-column accumulators alternate, so the MMAs of block j+1 overlap the epilogue of block j,
// and the gather of tile i+1 overlaps layer 2 and both epilogues of tile i.
// Epilogue algebra (the epilogue warps share the SM's issue slots with everything else, so every
// per-edge instruction counts): b1 rides in the contraction as a constant-one message column, so
// epilogue 1 is one cvt.rn.relu.bf16x2 per two values; y -> BN(ReLU(y + b2)) is monotone (increasing
// for a non-negative BN scale, decreasing otherwise), so rows of W2 whose BN scale is negative are
// negated when packed and epilogue 2 is a plain running max per edge, with bias / ReLU / BN applied
// once per (channel, target) to the extremal value -- bit-identical to applying them per edge.
// The [E, C+4], [E, H] and [E, C'] edge tensors of the reference never exist in HBM.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace p2w {
namespace {

constexpr int NT = 128;                        // edges per tile (MMA N)
constexpr int TPT = NT / 32;                   // targets per tile
constexpr int PAD_K = 32;                      // K extents (C + 5, H) are padded / required to multiples of this
constexpr int SLICE_K = 64;                    // k extent of one ring slice: four K=16 MMAs per wait / elect / commit
constexpr int SLICE_BYTES = 128 * SLICE_K * 2; // 16 KB: [8 k-chunks][128 rows][8 bf16]; the last slice of a block
constexpr int TAIL_BYTES = SLICE_BYTES / 2;    // may be a half one (K % 64 == 32: two MMAs, 8 KB copied)
constexpr int MAX_STAGES = 14;                  // weight-ring depth is chosen per layer shape
constexpr int EPI_THREADS = 128;                // warps 0-3 (one TMEM lane quarter each)
constexpr int GATHER_WARPS = 4;                 // warps 4-7
constexpr int GATHER_THREADS = GATHER_WARPS * 32;
// Four warpgroups: 0 epilogue (warps 0-3), 1 and 2 the two gather groups (warps 4-7, 8-11), 3 = MMA issue (warp 12),
// weight producer (warp 13) and two idle warps.  The kernel is launched at 64 registers per thread (two CTAs per
// SM); setmaxnreg moves eight registers per thread from each of the other three warpgroups (-> 56) to the epilogue (-> 88).
constexpr int W_GATHER = 4, W_MMA = 12, W_PROD = 13;
constexpr int THREADS = 512;
constexpr int REGS_EPI = 88, REGS_MMA = 56, REGS_GATHER = 56;
constexpr int VALID_SLOTS = 8;                  // per-target validity flags of the tiles in flight
constexpr int MAX_MSG_BUFS = 3;                 // msg tiles the gather warps may run ahead of layer 1
constexpr int LBO1 = NT * 16 + 16;             // k-chunk stride of the msg tile, padded against bank conflicts
constexpr int TMEM_COLS = 2 * NT;
constexpr unsigned FULL = 0xffffffffu;

// Timing experiments (tools/conv_timeline.py, tools/prof_conv.py): compile with -DP2W_CONV_INSTRUMENT and set
// the environment variable P2W_CONV_DEBUG to a bit mask -- 1 no weight streaming, 2 no feature gather, 4 empty
// epilogues, 8 no MMAs, 16 clock64 timeline of every role of CTA 0.  The default build has none of it.
#ifdef P2W_CONV_INSTRUMENT
__device__ long long g_timeline[16 * 512];       // one region per warp of CTA 0: the stores never stall on an atomic
// event = clock64 << 16 | tile << 8 | role << 4 | tag, tiles 20..27 of CTA 0 (tools/conv_timeline.py decodes)
#define P2W_TSR(p, it, role, tag)                                                                              \
    do {                                                                                                       \
        if (((p).debug & 16) && blockIdx.x == 0 && (it) >= 20 && (it) < 28 && (threadIdx.x & 31) == 0 &&        \
            ts_n < 512)                                                                                        \
            g_timeline[(threadIdx.x >> 5) * 512 + ts_n++] =                                                    \
                (clock64() << 16) | ((long long)(it) << 8) | ((role) << 4) | (tag);                            \
    } while (0)
#define P2W_DBG(p, bit) ((p).debug & (bit))
#else
#define P2W_DBG(p, bit) 0
#define P2W_TSR(p, it, role, tag) do { } while (0)
#endif

struct ConvTcParams {
    const void *x;                 // [n_src, C] FP32 or BF16 (x_bf16)
    const float *pos_src, *pos_tgt;
    const int32_t *nbr;
    const int64_t *tgt_index;       // optional: target t sits at pos_tgt[tgt_index[t]]
    int64_t n_tgt;
    int K, C, H, Co, K1p, NB1, NB2, num_tiles, x_bf16, out_bf16;
    int msg_bufs;                  // msg tiles in shared memory (the gather runs msg_bufs - 1 tiles ahead)
    int stages, resident;          // ring depth; resident: the ring holds ALL weight slices, loaded once
    int dup;                       // rows of W1 packed twice (H <= 64): epilogue 1 splits the columns four ways
    int debug;                     // P2W_CONV_INSTRUMENT builds only
    const unsigned char *wpack;
    const float *b1p, *b2p, *scale, *shift;
    void *out;                     // [n_tgt, Co] FP32 or BF16 (out_bf16)
};

// ---- tcgen05 / TMEM wrappers
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
// kind::f16 instruction descriptor: D=F32, A=B=BF16, A K-major, B K- or MN-major, M=128, N=NT
__host__ __device__ constexpr uint32_t instr_desc(bool b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn_major ? 1u : 0u) << 16) |
           (static_cast<uint32_t>(NT >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}
// Descriptors travel as (low word, high word): only the 14-bit start-address field in the low word moves inside
// the issue loop, so stepping a descriptor is ONE 32-bit uniform add instead of a 64-bit add on a register pair.
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// The same load WITHOUT the wait, and the wait as a separate statement that names the registers (so the compiler
// cannot move their first use in front of it): lets the load of the next 32 columns fly while the current 32
// are reduced -- an exposed TMEM round trip (~150 cycles) per chunk otherwise.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
// one lane of a converged warp (the warp-uniform issue pattern ptxas keeps in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void gather_bar(int grp) {          // named barriers 1, 2: one per gather group
    asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(GATHER_THREADS) : "memory");
}
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);   // .x (low half) = a
    return *reinterpret_cast<const uint32_t *>(&v);
}

__device__ __forceinline__ uint32_t pack_bf16_relu(float a, float b) {   // low half = relu(a), high half = relu(b)
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}

struct SmemLayout {
    uint32_t ring, b1, b2, sj, svalid, bars, tmem, total;
};
__host__ __device__ inline uint32_t msg_tile_bytes(int K1p) { return ((K1p / 8) * LBO1 + 127u) & ~127u; }
__host__ __device__ inline SmemLayout smem_layout(int K1p, int H, int stages, int msg_bufs) {
    SmemLayout L;
    L.ring = 0;
    L.b1 = L.ring + stages * SLICE_BYTES;
    L.b2 = L.b1 + msg_bufs * msg_tile_bytes(K1p);
    L.sj = L.b2 + (NT / 8) * (H * 16);
    L.svalid = L.sj + 4 * NT * 4;          // two rows of source indices (alternating) per gather group
    L.bars = (L.svalid + VALID_SLOTS * TPT * 4 + 7u) & ~7u;
    L.tmem = L.bars + 8 * (2 * MAX_STAGES + 6 + 2 * MAX_MSG_BUFS);
    L.total = L.tmem + 16;
    return L;
}

__global__ void __launch_bounds__(THREADS, 2) conv_tc2_kernel(const ConvTcParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const SmemLayout L = smem_layout(p.K1p, p.H, p.stages, p.msg_bufs);
    const int MB = p.msg_bufs;
    const uint32_t msg_bytes = msg_tile_bytes(p.K1p);
    const int STAGES = p.stages;
    unsigned char *ring = smem + L.ring;
    unsigned char *b1 = smem + L.b1;
    unsigned char *b2 = smem + L.b2;
    int *s_j = reinterpret_cast<int *>(smem + L.sj);
    int *s_valid = reinterpret_cast<int *>(smem + L.svalid);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars);
    uint64_t *ring_full = bars, *ring_empty = bars + MAX_STAGES;
    uint64_t *acc_full = bars + 2 * MAX_STAGES, *acc_empty = acc_full + 2;
    uint64_t *b2_full = acc_empty + 2, *b2_empty = b2_full + 1;
    uint64_t *b1_full = b2_empty + 1, *b1_empty = b1_full + MAX_MSG_BUFS;      // one pair per msg buffer
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + L.tmem);

    const int warp = __shfl_sync(FULL, static_cast<int>(threadIdx.x >> 5), 0);   // provably warp-uniform
#ifdef P2W_CONV_INSTRUMENT
    int ts_n = 0;
#endif
    const int lane = threadIdx.x & 31;
    const int kchunks = p.K1p >> 3;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&ring_full[s], 1); mbar_init(&ring_empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], EPI_THREADS); }
        for (int m = 0; m < MAX_MSG_BUFS; m++) { mbar_init(&b1_full[m], GATHER_THREADS); mbar_init(&b1_empty[m], 1); }
        mbar_init(b2_full, EPI_THREADS);
        mbar_init(b2_empty, 1);
        fence_barrier_init();
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // K padding of the msg tile (chunks after the geometry chunk) is zero for the whole kernel
    for (int m = 0; m < MB; m++) {
        for (int i = threadIdx.x; i < (kchunks - (p.C >> 3) - 1) * NT; i += THREADS) {
            const int kc = (p.C >> 3) + 1 + i / NT, n = i % NT;
            *reinterpret_cast<uint4 *>(b1 + m * msg_bytes + kc * LBO1 + n * 16) = make_uint4(0, 0, 0, 0);
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    const int my_tiles = (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                         static_cast<int>(gridDim.x);
    const int n1 = (p.K1p + SLICE_K - 1) / SLICE_K, n2 = (p.H + SLICE_K - 1) / SLICE_K;   // ring slices per accumulator block
    const bool tail1 = (p.K1p % SLICE_K) != 0, tail2 = (p.H % SLICE_K) != 0;               // last slice is a half one

    if (warp == W_PROD) {
        // ------------------------------------------------ weight producer
        reg_dec<REGS_MMA>();
        // The whole warp walks the (warp-uniform) loop, one elected lane issues: addresses and counters stay
        // in uniform registers, so the loop body is a wait, an expect_tx and a bulk copy.
        int slot = 0;
        uint32_t ph = 0;
        const int passes = P2W_DBG(p, 1) ? 0 : (p.resident ? (my_tiles > 0 ? 1 : 0) : my_tiles);
        for (int it = 0; it < passes; it++) {
            const unsigned char *src = p.wpack;
#pragma unroll 1
            for (int layer = 0; layer < 2; layer++) {
                const int ns = layer == 0 ? n1 : n2, total = (layer == 0 ? p.NB1 : p.NB2) * ns;
                const uint32_t last_bytes = (layer == 0 ? tail1 : tail2) ? TAIL_BYTES : SLICE_BYTES;
                for (int i = 0, s = 0; i < total; i++) {
                    const uint32_t bytes = s == ns - 1 ? last_bytes : static_cast<uint32_t>(SLICE_BYTES);
                    mbar_wait(&ring_empty[slot], ph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&ring_full[slot], bytes);
                        bulk_g2s(ring + slot * SLICE_BYTES, src, bytes, &ring_full[slot]);
                    }
                    src += bytes;
                    if (++s == ns) s = 0;
                    if (++slot == STAGES) { slot = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == W_MMA) {
        reg_dec<REGS_MMA>();
        // ------------------------------------------------ MMA issuer: warp-uniform loop, one elected lane
        // issues; descriptors advance by increments (a lone issuing warp is latency-bound on every
        // dependent instruction, so the per-slice body must stay a few instructions long).
        int slot = 0, acc = 0, mbuf = 0;
        uint32_t ph = 0, tph = 0, mph = 0, use0 = 0, use1 = 0;
        const uint64_t a_desc0 = smem_desc(smem_u32(ring), 2048, 128);            // + slot * 1024 (+ 256 per K = 16 step)
        const uint64_t b1_desc0 = smem_desc(smem_u32(b1), LBO1, 128);             // + 2 LBO1/16 per K = 16 step
        const uint64_t b2_desc0 = smem_desc(smem_u32(b2), 128, p.H * 16);         // + 16 per K = 16 step
        constexpr uint32_t ID1 = instr_desc(false), ID2 = instr_desc(true);
        const bool stream = !(p.resident || P2W_DBG(p, 1));
        for (int it = 0; it < my_tiles; it++) {
#pragma unroll 1
            for (int layer = 0; layer < 2; layer++) {
                P2W_TSR(p, it, 1 + layer, 0);
                mbar_wait(layer == 0 ? &b1_full[mbuf] : b2_full, layer == 0 ? mph : tph);
                tc_fence_after();
                P2W_TSR(p, it, 1 + layer, 1);
                const int nb = layer == 0 ? p.NB1 : p.NB2, ns = layer == 0 ? n1 : n2;
                const uint64_t b_desc0 = layer == 0 ? b1_desc0 + static_cast<uint32_t>(mbuf) * (msg_bytes >> 4) : b2_desc0;
                const uint32_t b_hi = static_cast<uint32_t>(b_desc0 >> 32), a_hi = static_cast<uint32_t>(a_desc0 >> 32);
                const uint32_t a_lo0 = static_cast<uint32_t>(a_desc0);
                const uint32_t bq = layer == 0 ? 2u * (LBO1 >> 4) : 16u;      // B descriptor step per K = 16
                const bool tail = layer == 0 ? tail1 : tail2;
                const uint32_t idesc = layer == 0 ? ID1 : ID2;
                for (int blk = 0; blk < nb; blk++) {
                    mbar_wait(&acc_empty[acc], ((acc ? use1 : use0) & 1) ^ 1);
                    tc_fence_after();
                    P2W_TSR(p, it, 1 + layer, 2);
                    const uint32_t d_addr = tmem_base + acc * NT;
                    uint32_t bd = static_cast<uint32_t>(b_desc0);
                    for (int s = 0; s < ns; s++) {
                        if (stream || (it == 0 && !P2W_DBG(p, 1))) mbar_wait(&ring_full[slot], ph);
                        const uint32_t ad = a_lo0 + static_cast<uint32_t>(slot) * (SLICE_BYTES >> 4);
                        const bool half = tail && s == ns - 1;
                        if (elect_one()) {
                            if (!P2W_DBG(p, 8)) {
                                umma(d_addr, ad, a_hi, bd, b_hi, idesc, s ? 1u : 0u);
                                umma(d_addr, ad + 256u, a_hi, bd + bq, b_hi, idesc, 1u);
                                if (!half) {
                                    umma(d_addr, ad + 512u, a_hi, bd + 2u * bq, b_hi, idesc, 1u);
                                    umma(d_addr, ad + 768u, a_hi, bd + 3u * bq, b_hi, idesc, 1u);
                                }
                            }
                            if (stream) umma_commit(&ring_empty[slot]);
                        }
                        bd += 4u * bq;
                        if (++slot == STAGES) { slot = 0; ph ^= 1; }
                    }
                    P2W_TSR(p, it, 1 + layer, 3);
                    if (elect_one()) umma_commit(&acc_full[acc]);
                    if (acc) use1++; else use0++;
                    acc ^= 1;
                }
                if (elect_one()) umma_commit(layer == 0 ? &b1_empty[mbuf] : b2_empty);
            }
            tph ^= 1;
            if (++mbuf == MB) { mbuf = 0; mph ^= 1; }
        }
        __syncwarp();
    } else if (warp > W_PROD) {
        reg_dec<REGS_MMA>();          // idle warps of warpgroup 3: setmaxnreg is a warpgroup-wide instruction
    } else if (warp >= W_GATHER) {
        // ------------------------------------------------ gather warps: two groups of four, group g takes the tiles
        // g, g+2, ... of this CTA.  A gather warp's work per tile is one long serial instruction stream (~350
        // instructions with their issue latencies, plus one exposed global-load latency for the feature rows): the
        // per-role timeline showed it pacing SA1 and SA2, so two groups keep two tiles in the making.  (With a
        // single msg tile both groups would write the same buffer and a parity wait cannot tell its second-next
        // use from the next one: one group works alone then.)
        reg_dec<REGS_GATHER>();
        const int gw = (warp - W_GATHER) & 3, grp = (warp - W_GATHER) >> 2;
        const int ngrp = MB >= 2 ? 2 : 1;
        int mbuf = grp % MB;
        uint32_t mph = static_cast<uint32_t>(grp / MB) & 1u;
        const int CPR = p.C >> 3;
        const int cpr_c = CPR < 32 ? CPR : 32;
        const int rpw = 32 / cpr_c;
        const int ck = lane % cpr_c, ri = lane / cpr_c;
        // Software pipeline over tiles (this warp owns target gw of every tile): the neighbour row and the
        // target index of tile it+2 and the positions of tile it+1 are in flight while the features of tile it
        // are fetched; NOTHING is consumed in the step that loads it (a ballot on a freshly loaded row would
        // expose a DRAM latency per tile), so a tile costs one global-load latency instead of a chain.
        static_assert(TPT == GATHER_WARPS, "one gather warp per target of a tile");
        auto load_raw = [&](int it, int &j, int64_t &ti) {
            j = -1;
            ti = -1;
            if (it < my_tiles) {
                const int64_t t = (static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(it) * gridDim.x) * TPT + gw;
                if (t < p.n_tgt) {
                    if (lane < p.K) j = p.nbr[t * p.K + lane];
                    ti = p.tgt_index ? p.tgt_index[t] : t;
                }
            }
        };
        auto resolve = [&](int &j, int64_t ti, unsigned &m, float4 &ps, float4 &pt) {
            m = __ballot_sync(FULL, j >= 0);
            const int jf = m ? __shfl_sync(FULL, j, __ffs(m) - 1) : 0;
            if (j < 0) j = jf;                     // padded slot: duplicate a valid edge (max unchanged)
            ps = __ldg(reinterpret_cast<const float4 *>(p.pos_src) + j);
            pt = ti >= 0 ? __ldg(reinterpret_cast<const float4 *>(p.pos_tgt) + ti) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        int j0, j1, j2;
        int64_t ti0, ti1, ti2;
        unsigned m0, m1;
        float4 pt0, pt1, ps0, ps1;
        load_raw(grp, j0, ti0);
        load_raw(grp + ngrp, j1, ti1);
        resolve(j0, ti0, m0, ps0, pt0);
        for (int it = grp; it < (grp < ngrp ? my_tiles : 0); it += ngrp) {
            const int tile = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
            const int64_t t0 = static_cast<int64_t>(tile) * TPT;
            (void)t0;
            unsigned char *msg = b1 + mbuf * msg_bytes;
            int *sj = s_j + (2 * grp + ((it / ngrp) & 1)) * NT;
            load_raw(it + 2 * ngrp, j2, ti2);                                   // two steps ahead: stays in flight
            resolve(j1, ti1, m1, ps1, pt1);                                     // tile it+1: its row was loaded a step ago
            const int n = gw * 32 + lane;
            sj[n] = j0;
            if (lane == 0) s_valid[(it & (VALID_SLOTS - 1)) * TPT + gw] = m0 ? 1 : 0;
            P2W_TSR(p, it, 3 + grp, 0);
            mbar_wait(&b1_empty[mbuf], mph ^ 1);   // layer 1 of the tile that last used this buffer is done
            P2W_TSR(p, it, 3 + grp, 1);
            {
                const float dx = ps0.x - pt0.x, dy = ps0.y - pt0.y, dz = ps0.z - pt0.z;
                float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
                for (int o = 16; o; o >>= 1) nrm = fmaxf(nrm, __shfl_xor_sync(FULL, nrm, o));
                const float den = nrm + 1e-8f;
                uint4 g;
                const float inv = 1.f / den;        // one division; the quotients are rounded to bf16 right below
                g.x = pack_bf16(dx * inv, dy * inv);
                g.y = pack_bf16(dz * inv, ps0.w);
                g.z = 0x00003F80u;                 // column C+4 = 1.0: carries b1 through the contraction
                g.w = 0;
                *reinterpret_cast<uint4 *>(msg + CPR * LBO1 + n * 16) = g;
            }
            j0 = j1; m0 = m1; pt0 = pt1; ps0 = ps1;
            j1 = j2; ti1 = ti2;
            gather_bar(grp);
            P2W_TSR(p, it, 3 + grp, 3);
            // feature rows: lanes run along a row (coalesced), 8 channels -> one 16-byte smem store
            if (P2W_DBG(p, 2)) {
            } else if (p.x_bf16) {
                const __nv_bfloat16 *xb = static_cast<const __nv_bfloat16 *>(p.x);
                for (int r0 = gw * rpw; r0 < NT; r0 += GATHER_WARPS * rpw * 4) {
                    uint4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int n = r0 + u * GATHER_WARPS * rpw + ri;
                        if (n < NT && ck < CPR)
                            v[u] = __ldg(reinterpret_cast<const uint4 *>(xb + static_cast<int64_t>(sj[n]) * p.C + ck * 8));
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int n = r0 + u * GATHER_WARPS * rpw + ri;
                        if (n < NT && ck < CPR) *reinterpret_cast<uint4 *>(msg + ck * LBO1 + n * 16) = v[u];
                    }
                }
            } else {
                const float *xf = static_cast<const float *>(p.x);
                for (int r0 = gw * rpw; r0 < NT; r0 += GATHER_WARPS * rpw * 2) {
                    float4 lo[2], hi[2];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int n = r0 + u * GATHER_WARPS * rpw + ri;
                        if (n < NT && ck < CPR) {
                            const float4 *src =
                                reinterpret_cast<const float4 *>(xf + static_cast<int64_t>(sj[n]) * p.C + ck * 8);
                            lo[u] = __ldg(src);
                            hi[u] = __ldg(src + 1);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int n = r0 + u * GATHER_WARPS * rpw + ri;
                        if (n < NT && ck < CPR) {
                            uint4 v;
                            v.x = pack_bf16(lo[u].x, lo[u].y);
                            v.y = pack_bf16(lo[u].z, lo[u].w);
                            v.z = pack_bf16(hi[u].x, hi[u].y);
                            v.w = pack_bf16(hi[u].z, hi[u].w);
                            *reinterpret_cast<uint4 *>(msg + ck * LBO1 + n * 16) = v;
                        }
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(&b1_full[mbuf]);
            P2W_TSR(p, it, 3 + grp, 2);
            mbuf += ngrp;
            while (mbuf >= MB) { mbuf -= MB; mph ^= 1; }
        }
    } else {
        reg_inc<REGS_EPI>();
        // ------------------------------------------------ epilogue warps (TMEM lane quarter q = warp)
        int acc = 0;
        uint32_t tph = 0, use0 = 0, use1 = 0;
        const int q = warp;
        const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16);
        for (int it = 0; it < my_tiles; it++) {
            const int tile = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
            const int64_t t0 = static_cast<int64_t>(tile) * TPT;
            // ---- epilogue 1: hid[e, h] = relu(D1^T[h, e] + b1[h]) as the MN-major B operand of layer 2
            for (int blk = 0; blk < p.NB1; blk++) {
                P2W_TSR(p, it, 5, 0);
                mbar_wait(&acc_full[acc], (acc ? use1 : use0) & 1);
                tc_fence_after();
                P2W_TSR(p, it, 5, 1);
                if (blk == 0) mbar_wait(b2_empty, tph ^ 1);   // layer 2 of the previous tile is done with hid
                P2W_TSR(p, it, 5, 2);
                // H <= 64: the rows of W1 were packed twice, lanes 64-127 repeat lanes 0-63, and the four warps
                // convert a quarter of the tile each; otherwise a warp whose channels are all padding skips the block
                const int h = p.dup ? ((32 * q + lane) & 63) : blk * 128 + 32 * q + lane;
                const int c_lo = p.dup ? (q >> 1) * 2 : 0;
                const int c_hi = P2W_DBG(p, 4) ? 0 : (p.dup ? c_lo + 2 : (blk * 128 + 32 * q < p.H ? NT / 32 : 0));
                for (int c = c_lo; c < c_hi; c++) {
                    const int n0 = c * 32;
                    uint32_t r[32];
                    tmem_ld32(lane_taddr + acc * NT + n0, r);
                    if (h < p.H) {
                        unsigned char *dst = b2 + (n0 >> 3) * (p.H * 16) + (h >> 3) * 128 + (h & 7) * 16;
#pragma unroll
                        for (int g = 0; g < 4; g++) {
                            uint4 v;
                            v.x = pack_bf16_relu(__uint_as_float(r[8 * g + 0]), __uint_as_float(r[8 * g + 1]));
                            v.y = pack_bf16_relu(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3]));
                            v.z = pack_bf16_relu(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
                            v.w = pack_bf16_relu(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]));
                            *reinterpret_cast<uint4 *>(dst + g * (p.H * 16)) = v;
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&acc_empty[acc]);
                if (acc) use1++; else use0++;
                acc ^= 1;
            }
            fence_proxy_async();
            mbar_arrive(b2_full);
            P2W_TSR(p, it, 5, 3);

            // ---- epilogue 2: out[t, c] = max_e BN(relu(D2^T[c, e] + b2[c]))
            const int *valid = s_valid + (it & (VALID_SLOTS - 1)) * TPT;
            for (int blk = 0; blk < p.NB2; blk++) {
                P2W_TSR(p, it, 6, 0);
                mbar_wait(&acc_full[acc], (acc ? use1 : use0) & 1);
                tc_fence_after();
                P2W_TSR(p, it, 6, 1);
                const int co = blk * 128 + 32 * q + lane;
                const float bias = p.b2p[co], sc = p.scale[co], sh = p.shift[co];
                const float sgn = sc < 0.f ? -1.f : 1.f;      // rows with a negative BN scale were packed negated
                uint32_t ra[32], rb[32];
                if (!P2W_DBG(p, 4)) {
                    tmem_ld32_issue(lane_taddr + acc * NT, ra);
                    tmem_ld_wait(ra);
                }
#pragma unroll
                for (int tt = 0; tt < (P2W_DBG(p, 4) ? 0 : TPT); tt++) {
                    uint32_t (&r)[32] = (tt & 1) ? rb : ra;          // this target's columns: already here
                    uint32_t (&rn)[32] = (tt & 1) ? ra : rb;         // the next target's: requested now, awaited below
                    if (tt + 1 < TPT) tmem_ld32_issue(lane_taddr + acc * NT + (tt + 1) * 32, rn);
                    float m0 = __uint_as_float(r[0]), m1 = __uint_as_float(r[1]), m2 = __uint_as_float(r[2]),
                          m3 = __uint_as_float(r[3]);
#pragma unroll
                    for (int e = 4; e < 32; e += 4) {
                        m0 = fmaxf(m0, __uint_as_float(r[e]));
                        m1 = fmaxf(m1, __uint_as_float(r[e + 1]));
                        m2 = fmaxf(m2, __uint_as_float(r[e + 2]));
                        m3 = fmaxf(m3, __uint_as_float(r[e + 3]));
                    }
                    const float ext = sgn * fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));   // the edge value that maximises BN(ReLU(.))
                    const float m = fmaf(fmaxf(ext + bias, 0.f), sc, sh);
                    const int64_t t = t0 + tt;
                    if (t < p.n_tgt && co < p.Co) {
                        const float v = valid[tt] ? m : 0.f;
                        if (p.out_bf16) static_cast<__nv_bfloat16 *>(p.out)[t * p.Co + co] = __float2bfloat16(v);
                        else static_cast<float *>(p.out)[t * p.Co + co] = v;
                    }
                    if (tt + 1 < TPT) tmem_ld_wait(rn);
                }
                tc_fence_before();
                mbar_arrive(&acc_empty[acc]);
                P2W_TSR(p, it, 6, 2);
                if (acc) use1++; else use0++;
                acc ^= 1;
            }
            tph ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                     : "memory");
    }
}

// w [R, Kreal] fp32 row-major -> bf16 [NB][Kp/8][128][8] (zero padded): every ring slice contiguous
// bias_col (may be NULL): an extra column Kreal holding bias[row] (b1 rides in the contraction);
// neg_if (may be NULL): rows with neg_if[row] < 0 are stored negated (W2 rows of negative BN scale);
// dup64: packed row r holds source row r % 64 (W1 with H <= 64: TMEM lanes 64-127 repeat lanes 0-63).
__global__ void prepack_kernel(const float *__restrict__ w, int R, int Kreal, int NB, int Kp,
                               const float *__restrict__ bias_col, const float *__restrict__ neg_if, int dup64,
                               __nv_bfloat16 *__restrict__ out) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t total = static_cast<int64_t>(NB) * Kp * 128;
    if (idx >= total) return;
    const int e = static_cast<int>(idx & 7);
    const int r = static_cast<int>((idx >> 3) & 127);
    const int kc = static_cast<int>((idx >> 10) % (Kp >> 3));
    const int blk = static_cast<int>((idx >> 10) / (Kp >> 3));
    const int row = dup64 ? ((blk * 128 + r) & 63) : blk * 128 + r, k = kc * 8 + e;
    float v = 0.f;
    if (row < R) {
        if (k < Kreal) v = w[static_cast<int64_t>(row) * Kreal + k];
        else if (k == Kreal && bias_col) v = bias_col[row];
        if (neg_if && neg_if[row] < 0.f) v = -v;
    }
    out[idx] = __float2bfloat16(v);
}

__global__ void padvec_kernel(const float *__restrict__ v, int n, int np, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) out[i] = i < n ? v[i] : 0.f;
}

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

struct TcPlan {
    int K1p, NB1, NB2;
    size_t w1_bytes, w2_bytes, off_b1, off_b2, off_scale, off_shift, total;
};
inline TcPlan tc_plan(int c_in, int hidden, int c_out) {
    TcPlan t;
    t.K1p = round_up(c_in + 5, PAD_K);         // + the constant-one column that carries b1
    t.NB1 = (hidden + 127) / 128;
    t.NB2 = (c_out + 127) / 128;
    t.w1_bytes = static_cast<size_t>(t.NB1) * t.K1p * 128 * 2;
    t.w2_bytes = static_cast<size_t>(t.NB2) * hidden * 128 * 2;
    t.off_b1 = t.w1_bytes + t.w2_bytes;
    t.off_b2 = t.off_b1 + sizeof(float) * t.NB1 * 128;
    t.off_scale = t.off_b2 + sizeof(float) * t.NB2 * 128;
    t.off_shift = t.off_scale + sizeof(float) * t.NB2 * 128;
    t.total = t.off_shift + sizeof(float) * t.NB2 * 128 + 256;
    return t;
}

}  // namespace
}  // namespace p2w

using namespace p2w;

#ifdef P2W_CONV_INSTRUMENT
extern "C" int p2wdbg_conv2_timeline(long long *dst_host, int n) {
    static long long host[16 * 512];
    cudaMemcpyFromSymbol(host, g_timeline, sizeof(host));
    int cnt = 0;
    for (int i = 0; i < 16 * 512 && cnt < n; i++)
        if (host[i]) dst_host[cnt++] = host[i];
    void *sym = nullptr;
    cudaGetSymbolAddress(&sym, g_timeline);
    cudaMemset(sym, 0, sizeof(host));
    return cnt;
}
#endif

size_t p2w_conv_tc2_ws_bytes(int32_t c_in, int32_t hidden, int32_t c_out) { return tc_plan(c_in, hidden, c_out).total; }

int p2w_conv_tc2_launch(const void *x, int x_bf16, const float *pos_src, const float *pos_tgt, const int32_t *nbr, int64_t n_src,
                       int64_t n_tgt, int32_t k, int32_t c_in, int32_t hidden, int32_t c_out, const float *w1,
                       const float *b1, const float *w2, const float *b2, const float *bn_scale,
                       const float *bn_shift, void *out, int out_bf16, void *ws, size_t ws_bytes, cudaStream_t st,
                       bool packed, const int64_t *tgt_index) {
    (void)n_src;
    P2W_REQUIRE(c_in % 8 == 0 && c_in >= 8 && c_in <= 256,
                "p2w_pointnet_conv_max(bf16): c_in=%d must be a multiple of 8 in [8, 256]", c_in);
    P2W_REQUIRE(hidden % PAD_K == 0, "p2w_pointnet_conv_max(bf16): hidden=%d must be a multiple of 32", hidden);
    P2W_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(pos_src) & 15u) == 0 &&
                    (reinterpret_cast<uintptr_t>(pos_tgt) & 15u) == 0 && (reinterpret_cast<uintptr_t>(ws) & 127u) == 0,
                "p2w_pointnet_conv_max(bf16): x / pos / workspace must be 16-byte aligned");
    const TcPlan t = tc_plan(c_in, hidden, c_out);
    P2W_REQUIRE(ws_bytes >= t.total, "p2w_pointnet_conv_max(bf16): workspace too small");
    // Weight-ring depth.  The ring must keep (L2 latency x consumption rate) bytes in flight, so it takes
    // whatever shared memory the tiles leave: all of it when one CTA per SM is the only option, up to
    // the two-CTAs-per-SM budget otherwise.  A layer whose slices all fit (SA1) keeps them resident.
    const int per_tile = t.NB1 * ((t.K1p + SLICE_K - 1) / SLICE_K) + t.NB2 * ((hidden + SLICE_K - 1) / SLICE_K);
    const unsigned budget2 = 113u * 1024u, budget1 = 227u * 1024u;
    // msg buffers: the gather (three dependent global loads per tile) runs ahead of layer 1 when the tiles
    // leave room for a second / third msg tile next to a weight ring of useful depth.
    int msg_bufs = 1, stages = 0, resident = 0;
    for (int mb = MAX_MSG_BUFS; mb >= 1; mb--) {
        const unsigned fixed = smem_layout(t.K1p, hidden, 0, mb).total;
        int st;
        if (fixed + 2u * SLICE_BYTES <= budget2) st = static_cast<int>((budget2 - fixed) / SLICE_BYTES);
        else st = fixed + 2u * SLICE_BYTES <= budget1 ? static_cast<int>((budget1 - fixed) / SLICE_BYTES) : 0;
        if (st > MAX_STAGES) st = MAX_STAGES;
        const int res = per_tile <= st ? 1 : 0;
        if (res || st >= 4 || mb == 1) {
            msg_bufs = mb;
            stages = res ? per_tile : st;
            resident = res;
            break;
        }
    }
    P2W_REQUIRE(stages >= 2, "p2w_pointnet_conv_max(bf16): layer too wide for one CTA");
    const SmemLayout L = smem_layout(t.K1p, hidden, stages, msg_bufs);
    P2W_REQUIRE(L.total <= 227 * 1024, "p2w_pointnet_conv_max(bf16): layer too wide for one CTA (%u bytes smem)",
                L.total);
    unsigned char *base = static_cast<unsigned char *>(ws);
    __nv_bfloat16 *w1p = reinterpret_cast<__nv_bfloat16 *>(base);
    __nv_bfloat16 *w2p = reinterpret_cast<__nv_bfloat16 *>(base + t.w1_bytes);
    float *b1p = reinterpret_cast<float *>(base + t.off_b1);
    float *b2p = reinterpret_cast<float *>(base + t.off_b2);
    float *scp = reinterpret_cast<float *>(base + t.off_scale);
    float *shp = reinterpret_cast<float *>(base + t.off_shift);
    const int dup = hidden <= 64 ? 1 : 0;
    if (!packed) {
        const int64_t n1 = static_cast<int64_t>(t.NB1) * t.K1p * 128, n2 = static_cast<int64_t>(t.NB2) * hidden * 128;
        P2W_LAUNCH(prepack_kernel, (unsigned)((n1 + 255) / 256), 256, 0, st)(w1, hidden, c_in + 4, t.NB1, t.K1p, b1, nullptr, dup, w1p);
        P2W_LAUNCH(prepack_kernel, (unsigned)((n2 + 255) / 256), 256, 0, st)(w2, c_out, hidden, t.NB2, hidden, nullptr, bn_scale, 0, w2p);
        P2W_LAUNCH(padvec_kernel, (t.NB1 * 128 + 255) / 256, 256, 0, st)(b1, hidden, t.NB1 * 128, b1p);
        P2W_LAUNCH(padvec_kernel, (t.NB2 * 128 + 255) / 256, 256, 0, st)(b2, c_out, t.NB2 * 128, b2p);
        P2W_LAUNCH(padvec_kernel, (t.NB2 * 128 + 255) / 256, 256, 0, st)(bn_scale, c_out, t.NB2 * 128, scp);
        P2W_LAUNCH(padvec_kernel, (t.NB2 * 128 + 255) / 256, 256, 0, st)(bn_shift, c_out, t.NB2 * 128, shp);
    }
    ConvTcParams p;
    p.x_bf16 = x_bf16; p.out_bf16 = out_bf16;
    p.stages = stages; p.resident = resident; p.msg_bufs = msg_bufs;
    p.dup = dup;
    p.debug = 0;
#ifdef P2W_CONV_INSTRUMENT
    {
        static int dbg = -1;
        if (dbg < 0) { const char *e = getenv("P2W_CONV_DEBUG"); dbg = e ? atoi(e) : 0; }
        p.debug = dbg;
    }
#endif
    p.tgt_index = tgt_index;
    p.x = x; p.pos_src = pos_src; p.pos_tgt = pos_tgt; p.nbr = nbr;
    p.n_tgt = n_tgt; p.K = k; p.C = c_in; p.H = hidden; p.Co = c_out;
    p.K1p = t.K1p; p.NB1 = t.NB1; p.NB2 = t.NB2;
    p.num_tiles = static_cast<int>((n_tgt + TPT - 1) / TPT);
    p.wpack = base; p.b1p = b1p; p.b2p = b2p; p.scale = scp; p.shift = shp; p.out = out;
    static int sm_count = 0;
    static unsigned smem_set = 0;
    if (!sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (sm_count <= 0) sm_count = kNumSMs;
    }
    if (L.total > smem_set) {
        cudaFuncSetAttribute(conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
        smem_set = L.total;
    }
    const int per_sm = (L.total <= 113 * 1024) ? 2 : 1;     // TMEM: 2 x 256 columns fit one SM
    int grid = sm_count * per_sm;
    if (grid > p.num_tiles) grid = p.num_tiles;
    P2W_LAUNCH(conv_tc2_kernel, grid, THREADS, L.total, st)(p);
    return check_launch("p2w_pointnet_conv_max(bf16)");
}
