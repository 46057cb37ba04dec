// dist.cu — multi-GPU entry points (placeholder until the single-GPU paths are green).
#include "ljmd_internal.cuh"
extern "C" {
int ljmd_get_unique_id(void*) { ljmd::set_error("multi-GPU not built yet"); return LJMD_E_UNSUPPORTED; }
int ljmd_create_dist(ljmd_t** out, const ljmd_params*, const void*, int32_t, int32_t) {
    if (out) *out = nullptr;
    ljmd::set_error("multi-GPU not built yet");
    return LJMD_E_UNSUPPORTED;
}
}
