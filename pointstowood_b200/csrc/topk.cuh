// topk.cuh -- warp-distributed sorted top-k shared by the brute-force sweep (neighbors.cu) and the
// cell-list search (grid_knn.cu).  Entry e of the ascending (d, index) list lives in slot e/32 of
// lane e%32; a candidate enters by ballot + shuffle insertion, so the final list does not depend on
// the order in which candidates are offered (ties on d break on the lower index, exactly like the
// serial insertion sort of torch_cluster's kernel).
#pragma once
#include "common.cuh"

namespace p2w {

constexpr unsigned TOPK_FULL = 0xffffffffu;

__device__ __forceinline__ bool key_less(float d, int i, float td, int ti) {
    return d < td || (d == td && i < ti);
}

template <int S>
struct TopK {
    float d[S];
    int i[S];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < S; s++) { d[s] = 1e10f; i[s] = -1; }
    }
    // insert (cd, ci) keeping ascending (d, i) order; entry e lives in slot e/32, lane e%32
    __device__ __forceinline__ void insert(float cd, int ci, int lane) {
        int pos = 0;
#pragma unroll
        for (int s = 0; s < S; s++) pos += __popc(__ballot_sync(TOPK_FULL, key_less(d[s], i[s], cd, ci)));
#pragma unroll
        for (int s = S - 1; s >= 0; s--) {
            float ud = __shfl_up_sync(TOPK_FULL, d[s], 1);
            int ui = __shfl_up_sync(TOPK_FULL, i[s], 1);
            if (s > 0) {
                float wd = __shfl_sync(TOPK_FULL, d[s - 1], 31);
                int wi = __shfl_sync(TOPK_FULL, i[s - 1], 31);
                if (lane == 0) { ud = wd; ui = wi; }
            }
            const int e = s * 32 + lane;
            if (e == pos) { d[s] = cd; i[s] = ci; }
            else if (e > pos) { d[s] = ud; i[s] = ui; }
        }
    }
    __device__ __forceinline__ void kth(int k, float &td, int &ti) const {
        const int e = k - 1;
        float vd = d[0];
        int vi = i[0];
#pragma unroll
        for (int s = 1; s < S; s++)
            if ((e >> 5) == s) { vd = d[s]; vi = i[s]; }
        td = __shfl_sync(TOPK_FULL, vd, e & 31);
        ti = __shfl_sync(TOPK_FULL, vi, e & 31);
    }
};


}  // namespace p2w
