// Error reporting and device queries shared by the C-ABI entry points.
#include <stdarg.h>

#include "common.cuh"

namespace dgtta {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace dgtta

extern "C" int dgtta_abi_version(void) { return DGTTA_ABI_VERSION; }
extern "C" const char *dgtta_last_error(void) { return dgtta::g_err; }
