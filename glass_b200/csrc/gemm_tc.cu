// tcgen05 / TMEM path of the label-mixed Linear pair -- placeholder until the kernel lands.
#include "common.cuh"

namespace glass {
bool pair_tc_supported(int, int, int, int64_t, int64_t, const void*, const void*) { return false; }
int pair_fwd_tc(const float*, int64_t, int, const float*, int64_t, int, const float*, const float*, const float*,
                const float*, const uint8_t*, float, int, float*, int64_t, float*, int64_t, int, cudaStream_t) {
    set_error("tcgen05 path not built");
    return GLASS_ERR_UNSUPPORTED;
}
int pair_bwd_dx_tc(const float*, int64_t, const float*, const float*, const float*, const uint8_t*, float, int, float*,
                   int64_t, int, float*, int64_t, int, int64_t, int, cudaStream_t) {
    set_error("tcgen05 path not built");
    return GLASS_ERR_UNSUPPORTED;
}
}  // namespace glass
