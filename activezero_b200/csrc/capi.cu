// Library identification and error text for the C ABI (include/az_stereo.h).
#include "common.cuh"

extern "C" const char* az_version(void) { return "az_stereo 0.1.0 sm_100a"; }

extern "C" const char* az_error_string(int code) {
    if (code == 0) return "success";
    if (code == AZ_ERR_BAD_ARG) return "az_stereo: bad argument (null pointer, non-positive or unsupported size)";
    if (code == AZ_ERR_UNALIGNED) return "az_stereo: pointer alignment";
    return cudaGetErrorString((cudaError_t)code);
}
