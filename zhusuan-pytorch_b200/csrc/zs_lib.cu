// Library plumbing: version, error strings, device info.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "zs_common.cuh"

namespace zs {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* where, cudaError_t e) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
}
void set_last_error_msg(const char* msg) { snprintf(g_last_error, sizeof(g_last_error), "%s", msg); }

int sm_count() {
    // per device: a process may drive several GPUs (the value only sizes grids, so a benign race on the cache is fine)
    static int cached[64] = {0};
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        (void)cudaGetLastError();
        return 148;  // B200
    }
    if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        return 148;
    }
    if (dev >= 0 && dev < 64) cached[dev] = n;
    return n;
}

int pdl_mask() {
    static const int mask = [] {
        const char* e = getenv("ZS_PDL");
        return e ? atoi(e) : (PDL_LATENT_FWD | PDL_FUSED | PDL_SCALE);
    }();
    return mask;
}

}  // namespace zs

extern "C" {

int zs_abi_version(void) { return ZS_ABI_VERSION; }

const char* zs_strerror(int code) {
    switch (code) {
        case ZS_OK: return "ok";
        case ZS_ERR_ARG: return "invalid argument";
        case ZS_ERR_DTYPE: return "unsupported dtype (float32 / float64 only)";
        case ZS_ERR_CUDA: return "CUDA error";
        case ZS_ERR_NO_DEVICE: return "no sm_100 CUDA device";
        case ZS_ERR_WORKSPACE: return "workspace too small";
        case ZS_ERR_UNSUPPORTED: return "shape not supported by this entry point";
        case ZS_ERR_ALIGN: return "pointer alignment";
        default: return "unknown error";
    }
}

const char* zs_last_error(void) { return zs::g_last_error; }

int zs_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        zs::set_last_error("cudaGetDevice", e);
        (void)cudaGetLastError();
        return ZS_ERR_NO_DEVICE;
    }
    int n = 0, ma = 0, mi = 0;
    ZS_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    ZS_CUDA_TRY(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev));
    ZS_CUDA_TRY(cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = n;
    if (cc_major) *cc_major = ma;
    if (cc_minor) *cc_minor = mi;
    return ZS_OK;
}

}  // extern "C"
