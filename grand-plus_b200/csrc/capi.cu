// capi.cu -- error plumbing and version/device queries of the C ABI (include/grandplus_b200.h).
#include "gp_common.cuh"

#include <string>

namespace {
thread_local std::string g_last_error;
}

void gp_set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

extern "C" {

const char *gp_last_error(void) { return g_last_error.c_str(); }

int gp_abi_version(void) { return GP_ABI_VERSION; }

int gp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();  // clear the sticky "no device" state so later calls report their own errors
        return 0;
    }
    return n;
}

}  // extern "C"
