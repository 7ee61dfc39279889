#include <stdarg.h>

#include "tds_common.cuh"

namespace tds {
std::string& last_error() {
    static thread_local std::string e;
    return e;
}
int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}
int sm_count() {
    // per device, cached (cudaGetDeviceProperties is slow)
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
    }
    return cached > 0 ? cached : 148;
}
}  // namespace tds

extern "C" int tds_version(void) { return 100; }
extern "C" const char* tds_last_error(void) { return tds::last_error().c_str(); }
