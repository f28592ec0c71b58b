// p24_api.cu — ABI version and error strings of libp24_b200.
#include <cuda_runtime.h>

#include "../../include/p24.h"

extern "C" int p24_abi_version(void) { return P24_ABI_VERSION; }

extern "C" const char* p24_error_string(int code) {
    switch (code) {
        case 0: return "success";
        case P24_E_BADARG: return "p24: bad argument (null pointer, non-positive size or misaligned workspace)";
        case P24_E_WORKSPACE: return "p24: workspace too small (see p24_workspace_bytes)";
        case P24_E_UNSUPPORTED: return "p24: unsupported configuration";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "p24: unknown error";
}
