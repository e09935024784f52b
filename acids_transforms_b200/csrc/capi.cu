// capi.cu — ABI version, thread-local error message, device attribute cache.
#include <stdarg.h>
#include "common.cuh"

namespace acids {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
        if (cached <= 0) cached = 148;
    }
    return cached;
}

}  // namespace acids

extern "C" ACIDS_API int acids_abi_version(void) { return ACIDS_ABI_VERSION; }
extern "C" ACIDS_API const char* acids_last_error(void) { return acids::g_err; }
