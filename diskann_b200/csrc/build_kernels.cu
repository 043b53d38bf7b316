// build_kernels.cu — device-side Vamana construction and exhaustive scan (placeholders until
// the batched build lands; they fail loudly rather than fall back to the CPU).
#include "dab_common.cuh"

using namespace dab;

extern "C" {

int dab_build(dab_index* idx, uint32_t, uint32_t, float, uint32_t) {
    (void)idx;
    return fail(DAB_ERR_NOT_READY, "dab_build: not implemented yet");
}

int dab_flat_knn(dab_index* idx, const void*, uint32_t, uint32_t, uint32_t*, float*) {
    (void)idx;
    return fail(DAB_ERR_NOT_READY, "dab_flat_knn: not implemented yet");
}

}  // extern "C"
