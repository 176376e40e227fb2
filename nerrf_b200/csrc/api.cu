// ABI version, error string, device info.
#include "common.cuh"

namespace nerrf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int& c = cached[dev & 63];
    if (c == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) c = n;
        else return 148;
    }
    return c;
}

}  // namespace nerrf

extern "C" int nerrf_abi_version(void) { return NERRF_ABI_VERSION; }

extern "C" const char* nerrf_last_error(void) { return nerrf::g_err; }

extern "C" int nerrf_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        nerrf::set_error("no CUDA device");
        return NERRF_ERR_NODEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        cudaGetLastError();
        nerrf::set_error("cudaGetDeviceProperties failed");
        return NERRF_ERR_NODEVICE;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (prop.major != 10) {
        nerrf::set_error("device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
        return NERRF_ERR_NODEVICE;
    }
    return NERRF_OK;
}
