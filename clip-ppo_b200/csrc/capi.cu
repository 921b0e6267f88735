// Library-level entry points of the C ABI (include/clipppo_b200.h): version, error strings.
#include "common.cuh"

namespace clipppo {
int& last_cuda_error_ref() {
    static thread_local int err = 0;
    return err;
}
}  // namespace clipppo

extern "C" int clipppo_abi_version(void) { return CLIPPPO_ABI_VERSION; }

extern "C" int clipppo_last_cuda_error(void) { return clipppo::last_cuda_error_ref(); }

extern "C" const char* clipppo_strerror(int status) {
    switch (status) {
        case CLIPPPO_OK: return "ok";
        case CLIPPPO_ERR_BAD_SHAPE: return "bad shape or size argument";
        case CLIPPPO_ERR_BAD_CHANNELS: return "contrast needs 1 or 3 channels";
        case CLIPPPO_ERR_BAD_PAD: return "reflect padding needs k/2 < min(H, W)";
        case CLIPPPO_ERR_NULL: return "required pointer is NULL";
        case CLIPPPO_ERR_WORKSPACE: return "workspace too small";
        case CLIPPPO_ERR_UNSUPPORTED: return "configuration not supported by the sm_100a kernels";
        case CLIPPPO_ERR_ALIGN: return "pointer or stride alignment requirement not met";
        case CLIPPPO_ERR_CUDA: return "CUDA call failed (see clipppo_last_cuda_error)";
        case CLIPPPO_ERR_DIM_MISMATCH: return "latent and embedding widths differ";
    }
    return "unknown status";
}
