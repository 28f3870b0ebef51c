// capi.cu -- library-level entry points of the C ABI (version, error strings).
#include "common.cuh"

namespace ssak {
static thread_local int g_last_cuda_error = 0;
void set_last_cuda_error(cudaError_t e) { g_last_cuda_error = (int)e; }
}  // namespace ssak

extern "C" int ssak_b200_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char *ssak_b200_strerror(int status) {
    switch (status) {
        case SSAK_OK: return "ok";
        case SSAK_ERR_INVALID_ARGUMENT: return "invalid argument";
        case SSAK_ERR_UNSUPPORTED: return "shape not supported by the sm_100a kernels";
        case SSAK_ERR_WORKSPACE: return "workspace too small";
        case SSAK_ERR_CUDA: return "CUDA runtime error (see ssak_b200_last_cuda_error)";
        default: return "unknown status";
    }
}

extern "C" int ssak_b200_last_cuda_error(void) { return ssak::g_last_cuda_error; }
