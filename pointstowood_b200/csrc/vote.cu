// vote.cu -- spatial vote over the classified points (src/predicter.py:113-142,
// PointCloudClassifier.compute_labels): for every ORIGINAL point and its k nearest classified points
//     pwood = median of the neighbours' wood probabilities (np.median: mean of the two middle
//             order statistics when k is even, in float64);
//     label = any_wood == 1:  argmax over classes c of sum_{pred == c} prob   (1 iff the wood-predicted
//                             neighbours carry strictly more probability mass than the others),
//             otherwise:      1 iff any neighbour's prediction exceeds any_wood.
// One warp per original point; a lane holds up to four neighbours (k <= 128); order statistics from a
// warp-wide bitonic sort of the probabilities, sums in float64 like numpy.
#include "common.cuh"

namespace p2w {
namespace {

constexpr unsigned FULL = 0xffffffffu;
// Ascending bitonic sort of 32 R values held R per lane; element e = r * 32 + lane.
template <int R>
__device__ __forceinline__ void warp_sort(float (&v)[R], int lane) {
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
                        const float a = v[r], b = v[rp];
                        const bool sw = up ? (b < a) : (a < b);
                        v[r] = sw ? b : a;
                        v[rp] = sw ? a : b;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const float o = __shfl_xor_sync(FULL, v[r], j);
                    const bool up = (((r * 32 + lane) & k2) == 0);
                    const bool lower = (lane & j) == 0;
                    v[r] = (up == lower) ? fminf(v[r], o) : fmaxf(v[r], o);
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
    float v[R];
    int cnt = 0, any = 0;
    double w0 = 0.0, w1 = 0.0;       // class votes (float64 sums, as numpy)
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int e = r * 32 + lane;
        const int j = e < k ? nbr[q * k + e] : -1;
        v[r] = j >= 0 ? prob[j] : __int_as_float(0x7f800000);     // missing neighbours sort last
        const int cls = j >= 0 ? pred[j] : -1;
        cnt += j >= 0;
        if (cls == 0) w0 += static_cast<double>(v[r]);
        if (cls == 1) w1 += static_cast<double>(v[r]);
        if (cls >= 0 && static_cast<float>(cls) > any_wood) any = 1;
    }
    for (int o = 16; o; o >>= 1) {
        cnt += __shfl_xor_sync(FULL, cnt, o);
        w0 += __shfl_xor_sync(FULL, w0, o);
        w1 += __shfl_xor_sync(FULL, w1, o);
        any |= __shfl_xor_sync(FULL, any, o);
    }
    // np.median: mean of the order statistics cnt/2 - 1 and cnt/2 (the same one when cnt is odd)
    warp_sort<R>(v, lane);
    const int hi = cnt >> 1, lo = (cnt & 1) ? hi : hi - 1;
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const float ar = __shfl_sync(FULL, v[r], lo & 31), br = __shfl_sync(FULL, v[r], hi & 31);
        if ((lo >> 5) == r) a = ar;
        if ((hi >> 5) == r) b = br;
    }
    if (lane == 0) {
        pwood[q] = cnt ? (static_cast<double>(a) + static_cast<double>(b)) * 0.5 : 0.0;
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
