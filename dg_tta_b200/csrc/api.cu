// Error reporting and device queries shared by the C-ABI entry points.
#include <stdarg.h>

#include <atomic>

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

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
unsigned long long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int sm_count()
{
    static std::atomic<int> cached[64];   // zero-initialised; concurrent first calls all store the same value
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

}  // namespace dgtta

namespace dgtta {
unsigned long long launches();
void preload_mind_fast();
void preload_mind_general();
void preload_gin();
void preload_gin_fused();
void preload_sampler();
void preload_philox();
void preload_consistency();
void preload_resize();
}

// CUDA loads kernels lazily, on their first launch (a few ms each).  GIN alone has 22 instantiations selected by the
// random kernel sizes of a call, so without this a "warm" process still hits cold kernels for many steps.
extern "C" int dgtta_preload_kernels(void)
{
    dgtta::preload_mind_fast();
    dgtta::preload_mind_general();
    dgtta::preload_gin();
    dgtta::preload_gin_fused();
    dgtta::preload_sampler();
    dgtta::preload_philox();
    dgtta::preload_consistency();
    dgtta::preload_resize();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}
extern "C" uint64_t dgtta_launch_count(void) { return dgtta::launches(); }
extern "C" int dgtta_abi_version(void) { return DGTTA_ABI_VERSION; }
extern "C" const char *dgtta_last_error(void) { return dgtta::g_err; }
