// Error state, version and device queries of liborbit_b200.
#include "oa_common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void oa_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* oa_last_error(void) { return g_err; }

extern "C" int oa_abi_version(void) { return OA_ABI_VERSION; }

extern "C" int oa_device_info(int* sm_count, int* cc_major, int* cc_minor,
                              int64_t* l2_bytes, int64_t* hbm_bytes) {
    int dev = 0;
    OA_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    OA_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (l2_bytes) *l2_bytes = (int64_t)prop.l2CacheSize;
    if (hbm_bytes) *hbm_bytes = (int64_t)prop.totalGlobalMem;
    return OA_OK;
}

extern "C" size_t oa_track_args_size(void) { return sizeof(oa_track_args); }
extern "C" size_t oa_synth_params_size(void) { return sizeof(oa_synth_params); }
