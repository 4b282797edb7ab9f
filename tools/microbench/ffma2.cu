// Microbenchmark: FFMA vs FFMA2 (packed f32x2) issue/throughput on sm_100a, alone and mixed with LDS.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int MODE>  // 0: FFMA scalar, 1: FFMA2, 2: FFMA + LDS mix (4:1), 3: FFMA2 + LDS mix (2:1)
__global__ void __launch_bounds__(512) bench(float *out, int iters, float g)
{
    __shared__ float sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i * 0.001f;
    __syncthreads();
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.01f + i;
    unsigned long long p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = ((unsigned long long)__float_as_uint(a[2 * i + 1]) << 32) | __float_as_uint(a[2 * i]);
    const unsigned long long gg = ((unsigned long long)__float_as_uint(g) << 32) | __float_as_uint(g);
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], g, 0.5f);
            if (MODE == 2) {
#pragma unroll
                for (int r = 0; r < 32; ++r) acc += sm[(threadIdx.x + r * 33 + it) & 2047];
            }
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], gg, gg);
            if (MODE == 3) {
#pragma unroll
                for (int r = 0; r < 32; ++r) acc += sm[(threadIdx.x + r * 33 + it) & 2047];
            }
        }
    }
    float s = acc;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, float *d, int sms)
{
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<MODE><<<sms, 512>>>(d, 10, 1.0001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench<MODE><<<sms, 512>>>(d, iters, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma_lane = (double)sms * 512 * iters * 128.0;  // lane-FMAs (128 per thread per iter in every mode)
    printf("%-28s %8.3f ms  %7.1f lane-FMA/ns  = %6.1f FMA/clk/SM @1.9GHz\n", name, ms, fma_lane / ms / 1e6,
           fma_lane / ms / 1e6 / sms / 1.9);
}

int main()
{
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *d;
    cudaMalloc(&d, sizeof(float) * sms * 512);
    run<0>("FFMA  (128/thread/iter)", d, sms);
    run<1>("FFMA2 (64/thread/iter)", d, sms);
    run<2>("FFMA  + 32 LDS", d, sms);
    run<3>("FFMA2 + 32 LDS", d, sms);
    run<0>("FFMA  again", d, sms);
    run<1>("FFMA2 again", d, sms);
    return 0;
}
