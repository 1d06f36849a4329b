// conv_tc.cu -- K5 on tcgen05 tensor cores (placeholder until the TMEM kernel lands).
#include "common.cuh"
using namespace p2w;
size_t p2w_conv_tc_ws_bytes(int32_t, int32_t, int32_t) { return 256; }
int p2w_conv_tc_launch(const float *, const float *, const float *, const int32_t *, int64_t, int64_t, int32_t,
                       int32_t, int32_t, int32_t, const float *, const float *, const float *, const float *,
                       const float *, const float *, float *, void *, size_t, cudaStream_t) {
    set_error("p2w_pointnet_conv_max: BF16 tensor-core mode not built");
    return P2W_EINVAL;
}
