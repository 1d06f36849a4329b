// vote.cu -- spatial vote over the classified points (src/predicter.py:113-142,
// PointCloudClassifier.compute_labels): for every ORIGINAL point and its k nearest classified points
//     pwood = median of the neighbours' wood probabilities (np.median: mean of the two middle
//             order statistics when k is even, in float64);
//     label = any_wood == 1:  argmax over classes c of sum_{pred == c} prob   (1 iff the wood-predicted
//                             neighbours carry strictly more probability mass than the others),
//             otherwise:      1 iff any neighbour's prediction exceeds any_wood.
// One warp per original point; a lane holds up to four neighbours (k <= 128); order statistics from a
// warp-wide bitonic sort of (probability, class) keys, sums in float64 like numpy -- taken over the SORTED
// sequence, so the result is a function of the neighbour SET: the order of a row of `nbr` does not matter
// (the search may return its heap unsorted, and a plot sharded over GPUs votes exactly like one GPU).
#include "common.cuh"

namespace p2w {
namespace {

constexpr unsigned FULL = 0xffffffffu;
// Ascending bitonic sort of 32 R values held R per lane; element e = r * 32 + lane.
template <int R>
__device__ __forceinline__ void warp_sort(uint32_t (&v)[R], int lane) {
#pragma unroll
    for (int k2 = 2; k2 <= 32 * R; k2 <<= 1) {
#pragma unroll
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            if (j >= 32) {                      // partner lives in the same lane
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int rp = r ^ (j >> 5);
                    if (rp > r) {
                        const bool up = (((r * 32 + lane) & k2) == 0);
                        const uint32_t a = v[r], b = v[rp];
                        const bool sw = up ? (b < a) : (a < b);
                        v[r] = sw ? b : a;
                        v[rp] = sw ? a : b;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const uint32_t o = __shfl_xor_sync(FULL, v[r], j);
                    const bool up = (((r * 32 + lane) & k2) == 0);
                    const bool lower = (lane & j) == 0;
                    v[r] = (up == lower) ? min(v[r], o) : max(v[r], o);
                }
            }
        }
    }
}

template <int R>
__global__ void __launch_bounds__(256) vote_kernel(const int32_t *__restrict__ nbr, int64_t n, int k,
                                                   const float *__restrict__ prob, const uint8_t *__restrict__ pred,
                                                   float any_wood, uint8_t *__restrict__ label,
                                                   double *__restrict__ pwood) {
    const int64_t q = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= n) return;
    // key = (bits of prob) << 1 | class: prob in [0, 1] has a non-negative bit pattern that orders like the value
    uint32_t v[R];
    int cnt = 0, any = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int e = r * 32 + lane;
        const int j = e < k ? nbr[q * k + e] : -1;
        v[r] = 0xFFFFFFFFu;                                          // missing neighbours sort last
        if (j >= 0) {
            const int cls = pred[j];
            v[r] = (__float_as_uint(prob[j]) << 1) | static_cast<uint32_t>(cls & 1);
            cnt++;
            if (static_cast<float>(cls) > any_wood) any = 1;
        }
    }
    warp_sort<R>(v, lane);
    double w0 = 0.0, w1 = 0.0;       // class votes (float64 sums, as numpy), over the sorted sequence
#pragma unroll
    for (int r = 0; r < R; r++) {
        if (v[r] != 0xFFFFFFFFu) {
            const double pr = static_cast<double>(__uint_as_float(v[r] >> 1));
            if (v[r] & 1u) w1 += pr; else w0 += pr;
        }
    }
    for (int o = 16; o; o >>= 1) {
        cnt += __shfl_xor_sync(FULL, cnt, o);
        w0 += __shfl_xor_sync(FULL, w0, o);
        w1 += __shfl_xor_sync(FULL, w1, o);
        any |= __shfl_xor_sync(FULL, any, o);
    }
    // np.median: mean of the order statistics cnt/2 - 1 and cnt/2 (the same one when cnt is odd)
    const int hi = cnt >> 1, lo = (cnt & 1) ? hi : hi - 1;
    uint32_t a = 0, b = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const uint32_t ar = __shfl_sync(FULL, v[r], lo & 31), br = __shfl_sync(FULL, v[r], hi & 31);
        if ((lo >> 5) == r) a = ar;
        if ((hi >> 5) == r) b = br;
    }
    if (lane == 0) {
        pwood[q] = cnt ? (static_cast<double>(__uint_as_float(a >> 1)) + static_cast<double>(__uint_as_float(b >> 1))) * 0.5 : 0.0;
        label[q] = static_cast<uint8_t>(any_wood == 1.0f ? (w1 > w0 ? 1 : 0) : any);
    }
}

}  // namespace
}  // namespace p2w

using namespace p2w;

extern "C" int p2w_spatial_vote(const int32_t *nbr, int64_t n, int32_t k, const float *prob, const uint8_t *pred,
                                float any_wood, uint8_t *label, double *pwood, p2w_stream_t stream) {
    P2W_REQUIRE(k >= 1 && k <= 128, "p2w_spatial_vote: k=%d outside [1,128]", k);
    if (n == 0) return P2W_OK;
    const unsigned blocks = (unsigned)((n * 32 + 255) / 256);
    cudaStream_t st = as_stream(stream);
    if (k <= 32) P2W_LAUNCH(vote_kernel<1>, blocks, 256, 0, st)(nbr, n, k, prob, pred, any_wood, label, pwood);
    else if (k <= 64) P2W_LAUNCH(vote_kernel<2>, blocks, 256, 0, st)(nbr, n, k, prob, pred, any_wood, label, pwood);
    else P2W_LAUNCH(vote_kernel<4>, blocks, 256, 0, st)(nbr, n, k, prob, pred, any_wood, label, pwood);
    return check_launch("p2w_spatial_vote");
}
