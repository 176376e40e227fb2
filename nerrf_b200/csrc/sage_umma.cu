// tcgen05 fused GraphSAGE-T layer (placeholder until the UMMA kernel lands).
#include "common.cuh"
namespace nerrf {
bool sage_umma_available() { return false; }
int sage_layer_umma(const float*, const void*, int, const int32_t*, const float*, const float*, const float*, float*,
                    int64_t, int64_t, int64_t, int, int, cudaStream_t) {
    set_error("UMMA layer kernel not built");
    return NERRF_ERR_INVALID;
}
}  // namespace nerrf
