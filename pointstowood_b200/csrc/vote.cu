// vote.cu -- spatial vote over the classified points (src/predicter.py:113-142,
// PointCloudClassifier.compute_labels): for every ORIGINAL point and its k nearest classified points
//     pwood = median of the neighbours' wood probabilities (np.median: mean of the two middle
//             order statistics when k is even, in float64);
//     label = any_wood == 1:  argmax over classes c of sum_{pred == c} prob   (1 iff the wood-predicted
//                             neighbours carry strictly more probability mass than the others),
//             otherwise:      1 iff any neighbour's prediction exceeds any_wood.
// One warp per original point; a lane holds up to four neighbours (k <= 128); order statistics by rank
// counting over shuffles (ties broken by neighbour slot), sums in float64 like numpy.
#include "common.cuh"

namespace p2w {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int VOTE_PER_LANE = 4;

__global__ void __launch_bounds__(256) vote_kernel(const int32_t *__restrict__ nbr, int64_t n, int k,
                                                   const float *__restrict__ prob, const uint8_t *__restrict__ pred,
                                                   float any_wood, uint8_t *__restrict__ label,
                                                   double *__restrict__ pwood) {
    const int64_t q = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= n) return;
    float v[VOTE_PER_LANE];
    int cls[VOTE_PER_LANE];
    int cnt = 0;
#pragma unroll
    for (int u = 0; u < VOTE_PER_LANE; u++) {
        const int e = u * 32 + lane;
        const int j = e < k ? nbr[q * k + e] : -1;
        v[u] = j >= 0 ? prob[j] : __int_as_float(0x7f800000);     // missing neighbours sort last
        cls[u] = j >= 0 ? pred[j] : -1;
        cnt += j >= 0;
    }
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
    // class votes (float64 sums, as numpy)
    double w0 = 0.0, w1 = 0.0;
    int any = 0;
#pragma unroll
    for (int u = 0; u < VOTE_PER_LANE; u++) {
        if (cls[u] == 0) w0 += static_cast<double>(v[u]);
        if (cls[u] == 1) w1 += static_cast<double>(v[u]);
        if (cls[u] >= 0 && static_cast<float>(cls[u]) > any_wood) any = 1;
    }
    for (int o = 16; o; o >>= 1) {
        w0 += __shfl_xor_sync(FULL, w0, o);
        w1 += __shfl_xor_sync(FULL, w1, o);
        any |= __shfl_xor_sync(FULL, any, o);
    }
    // order statistics cnt/2 - 1 (when cnt is even) and cnt/2 by rank counting
    int rank[VOTE_PER_LANE] = {0, 0, 0, 0};
    const int slots = (k + 31) >> 5;
    for (int u2 = 0; u2 < slots; u2++) {
        for (int l = 0; l < 32; l++) {
            float o = v[0];
#pragma unroll
            for (int u = 1; u < VOTE_PER_LANE; u++)
                if (u == u2) o = v[u];
            o = __shfl_sync(FULL, o, l);
            const int oe = u2 * 32 + l;
#pragma unroll
            for (int u = 0; u < VOTE_PER_LANE; u++) {
                const int e = u * 32 + lane;
                rank[u] += (o < v[u] || (o == v[u] && oe < e)) ? 1 : 0;
            }
        }
    }
    const int hi = cnt >> 1, lo = (cnt & 1) ? hi : hi - 1;
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int u = 0; u < VOTE_PER_LANE; u++) {
        if (u * 32 + lane < k) {
            if (rank[u] == lo) a = static_cast<double>(v[u]);
            if (rank[u] == hi) b = static_cast<double>(v[u]);
        }
    }
    for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(FULL, a, o);      // exactly one lane holds each of the two (others add 0.0)
        b += __shfl_xor_sync(FULL, b, o);
    }
    if (lane == 0) {
        pwood[q] = cnt ? (a + b) * 0.5 : 0.0;
        label[q] = static_cast<uint8_t>(any_wood == 1.0f ? (w1 > w0 ? 1 : 0) : any);
    }
}

}  // namespace
}  // namespace p2w

using namespace p2w;

extern "C" int p2w_spatial_vote(const int32_t *nbr, int64_t n, int32_t k, const float *prob, const uint8_t *pred,
                                float any_wood, uint8_t *label, double *pwood, p2w_stream_t stream) {
    P2W_REQUIRE(k >= 1 && k <= 32 * VOTE_PER_LANE, "p2w_spatial_vote: k=%d outside [1,%d]", k, 32 * VOTE_PER_LANE);
    if (n == 0) return P2W_OK;
    P2W_LAUNCH(vote_kernel, (unsigned)((n * 32 + 255) / 256), 256, 0, as_stream(stream))(nbr, n, k, prob, pred, any_wood, label, pwood);
    return check_launch("p2w_spatial_vote");
}
