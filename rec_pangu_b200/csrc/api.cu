// Library-level entry points.
#include "common.cuh"

RPB_API int rpb_version(void) { return 1; }

RPB_API const char* rpb_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case RPB_ERR_UNSUPPORTED: return "shape outside what the sm_100a kernels support";
        case RPB_ERR_BAD_ARG: return "bad argument";
        case RPB_ERR_NO_DRIVER: return "cuTensorMapEncodeTiled not available from the CUDA driver";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}
