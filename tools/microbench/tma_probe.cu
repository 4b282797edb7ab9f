// Developer probe: minimal TMA tensor loads in several flavours (not product code).
//   tma_probe <variant>   0: 4-D map as __grid_constant__ struct member   1: 2-D map   2: map in global memory
//                         3: encoder from dlopen(libcuda)                 4: plain cp.async.bulk (no tensor map)
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct Pm { alignas(64) CUtensorMap m; const CUtensorMap *gm; float *out; const float *src; int bytes; int variant; int cx, cy; };
__global__ void k(const __grid_constant__ Pm p)
{
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bar;
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
    const unsigned d = (unsigned)__cvta_generic_to_shared(sm) + (p.variant == 8 ? 16u : 0u);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(p.bytes) : "memory");
        if (p.variant == 1)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(d), "l"(reinterpret_cast<uint64_t>(&p.m)), "r"(b), "r"(0), "r"(3) : "memory");
        else if (p.variant == 2)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(d), "l"(reinterpret_cast<uint64_t>(p.gm)), "r"(b), "r"(0), "r"(0), "r"(3), "r"(0) : "memory");
        else if (p.variant == 4)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(d), "l"(p.src), "r"(p.bytes), "r"(b) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(d), "l"(reinterpret_cast<uint64_t>(&p.m)), "r"(b), "r"(p.cx), "r"(p.cy), "r"(3), "r"(0) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(b) : "memory");
    for (int i = threadIdx.x; i < p.bytes / 4; i += blockDim.x) p.out[i] = sm[i + (p.variant == 8 ? 4 : 0)];
}
int main(int argc, char **argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    cudaFree(0);
    EncodeTiledFn enc = nullptr;
    if (variant == 3) {
        void *h = dlopen("libcuda.so.1", RTLD_NOW);
        enc = h ? (EncodeTiledFn)dlsym(h, "cuTensorMapEncodeTiled") : nullptr;
    } else {
        void *fp = nullptr; cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
        printf("entry point: err %d status %d ptr %p\n", (int)e, (int)q, fp);
        enc = (EncodeTiledFn)fp;
    }
    if (!enc) { printf("no encoder\n"); return 2; }
    const int W = 64, H = 32, D = 24, C = 12;
    float *g, *out; size_t n = (size_t)W * H * D * C;
    cudaMalloc(&g, n * 4); cudaMalloc(&out, 1 << 20);
    float *h = (float *)malloc(n * 4); for (size_t i = 0; i < n; ++i) h[i] = (float)i;
    cudaMemcpy(g, h, n * 4, cudaMemcpyHostToDevice);
    Pm p; p.out = out; p.variant = variant; p.src = g + 3 * H * W;
    CUresult r;
    if (variant == 1) {
        const cuuint64_t dims[2] = {W, (cuuint64_t)H * D * C}; const cuuint64_t str[1] = {W * 4};
        const cuuint32_t box[2] = {32, 16}; const cuuint32_t es[2] = {1, 1};
        r = enc(&p.m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t dims[4] = {W, H, D, C}; const cuuint64_t str[3] = {W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)D * H * W * 4};
        const cuuint32_t box[4] = {32, 16, 1, 1}; const cuuint32_t es[4] = {1, 1, 1, 1};
        r = enc(&p.m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                (variant == 6 || variant == 7) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    p.cx = (variant == 5 || variant == 7) ? -2 : 0; p.cy = p.cx;
    p.bytes = 32 * 16 * 4;
    CUtensorMap *gm; cudaMalloc(&gm, sizeof(CUtensorMap)); cudaMemcpy(gm, &p.m, sizeof(CUtensorMap), cudaMemcpyHostToDevice); p.gm = gm;
    const unsigned long long *w = (const unsigned long long *)&p.m;
    printf("variant %d encode %d desc:", variant, (int)r);
    for (int i = 0; i < 16; ++i) printf(" %016llx", w[i]);
    printf("\n");
    k<<<1, 128, p.bytes + 128>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    float v[2] = {0, 0};
    if (e == cudaSuccess) cudaMemcpy(v, out, 8, cudaMemcpyDeviceToHost);
    printf("variant %d: run %s, values %.0f %.0f (expect %d %d)\n", variant, cudaGetErrorString(e), v[0], v[1], 3 * H * W, 3 * H * W + 1);
    return e != cudaSuccess;
}
