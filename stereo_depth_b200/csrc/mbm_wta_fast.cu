// Kernel B, specialised variant (placeholder until the register-blocked kernel lands).
#include "common.cuh"

namespace sd {

bool mbm_wta_fast_supported(const Geom &) { return false; }

cudaError_t launch_mbm_wta_fast(const Geom &, int, const Scratch &, float *, float *, cudaStream_t) {
    return cudaErrorNotSupported;
}

}  // namespace sd
