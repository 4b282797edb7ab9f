// Consistency-loss reductions of the TTA step.
//
// Replaces the elementwise chain of dg_tta/tta/tta.py:263-269 + dg_tta/tta/torch_utils.py:90-104 over the two
// inverse-warped logit tensors target_a, target_b [B,C,D,H,W]:
//     m      = (sum_c a > 0) * (sum_c b > 0)                       common-content mask        (tta.py:264-266)
//     pa, pb = softmax_c(a) * m, softmax_c(b) * m                                              (tta.py:267-268)
//     N[b,c] = sum_v 2 pa pb          D[b,c] = sum_v (pa + pb)^2                               (torch_utils.py:94-95)
// (the reference then forms dice = (N/V) / (0.5 D/V) per (b,c) and loss = 1 - mean dice[:, 1:]; that last step is a
// handful of scalars and stays in torch, which also gives the gradient w.r.t. N and D for free).
// forward : one pass over both tensors -> sums[B,C,2] = {N, D}  (reference: ~10 passes over C*V floats)
// backward: one pass -> d loss / d a, given g = d loss / d sums:
//     q_c = m (2 gN_c pb_c + 2 gD_c (pa_c + pb_c)),   grad_a_c = pa_c (q_c - sum_j pa_j q_j)   (softmax Jacobian; the
//     mask is piecewise constant, its gradient is zero as in torch)
// A thread owns whole voxels (all C channels in registers for C <= 32), so softmax needs no cross-thread traffic;
// channel sums are reduced per block and added to the global sums with double-precision atomics.
#include "common.cuh"

namespace dgtta {
namespace closs {

constexpr int THREADS = 256;

template <int CMAX>
__device__ __forceinline__ bool load_softmax(const float *a, const float *b, size_t V, size_t v, int C, float (&pa)[CMAX],
                                             float (&pb)[CMAX])
{
    float sa = 0.f, sb = 0.f, ma = -__int_as_float(0x7f800000), mb = ma;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
            pa[c] = __ldg(a + (size_t)c * V + v);
            pb[c] = __ldg(b + (size_t)c * V + v);
            sa += pa[c]; sb += pb[c];
            ma = fmaxf(ma, pa[c]); mb = fmaxf(mb, pb[c]);
        }
    }
    if (!(sa > 0.f && sb > 0.f)) return false;   // outside the common content: contributes nothing, gradient zero
    float ea = 0.f, eb = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
            pa[c] = expf(pa[c] - ma); pb[c] = expf(pb[c] - mb);
            ea += pa[c]; eb += pb[c];
        }
    }
    const float ia = 1.f / ea, ib = 1.f / eb;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < C) { pa[c] *= ia; pb[c] *= ib; }
    }
    return true;
}

// Warp reduce-scatter of 32 values per lane: after the five exchange steps lane l holds the warp-wide sum of entry l
// (31 shuffles instead of 5 x 32 for a butterfly all-reduce of every entry).
__device__ __forceinline__ float reduce_scatter32(float (&x)[32], int lane)
{
#pragma unroll
    for (int n = 16; n >= 1; n >>= 1) {
        const bool up = lane & n;
#pragma unroll
        for (int j = 0; j < n; ++j) {
            const float send = up ? x[j] : x[j + n];
            const float keep = up ? x[j + n] : x[j];
            x[j] = keep + __shfl_xor_sync(0xffffffffu, send, n);
        }
    }
    return x[0];
}

// C <= 16: the 2C per-voxel contributions of a warp's 32 voxels are reduce-scattered every iteration, so a thread
// carries ONE accumulator instead of 2C (64 -> ~50 registers, twice the resident warps; the pass is latency-bound).
template <int CMAX>
__global__ void __launch_bounds__(THREADS, CMAX <= 16 ? 3 : 1) sums_kernel(const float *__restrict__ ta, const float *__restrict__ tb,
                                                                          double *__restrict__ sums, int C, size_t V)
{
    __shared__ float red[THREADS / 32][2 * CMAX];
    const int b = blockIdx.y;
    const float *a = ta + (size_t)b * C * V, *bb = tb + (size_t)b * C * V;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (CMAX <= 16) {
        float acc = 0.f;   // lane l: entry l of {N_0, D_0, N_1, D_1, ...}
        // all lanes of a warp iterate together (the shuffles need them): round the trip count up per warp
        const size_t stride = (size_t)gridDim.x * THREADS;
        for (size_t v0 = (size_t)blockIdx.x * THREADS + (threadIdx.x & ~31); v0 < V; v0 += stride) {
            const size_t v = v0 + lane;
            float pa[CMAX], pb[CMAX], x[32];
            const bool live = v < V && load_softmax<CMAX>(a, bb, V, v, C, pa, pb);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                if (c < CMAX && c < C && live) {
                    const float s = pa[c < CMAX ? c : 0] + pb[c < CMAX ? c : 0];
                    x[2 * c] = 2.f * pa[c < CMAX ? c : 0] * pb[c < CMAX ? c : 0];
                    x[2 * c + 1] = s * s;
                } else {
                    x[2 * c] = 0.f; x[2 * c + 1] = 0.f;
                }
            }
            acc += reduce_scatter32(x, lane);
        }
        if (lane < 2 * CMAX) red[warp][lane] = acc;
    } else {
        float accN[CMAX], accD[CMAX];
#pragma unroll
        for (int c = 0; c < CMAX; ++c) accN[c] = accD[c] = 0.f;
        for (size_t v = (size_t)blockIdx.x * THREADS + threadIdx.x; v < V; v += (size_t)gridDim.x * THREADS) {
            float pa[CMAX], pb[CMAX];
            if (!load_softmax<CMAX>(a, bb, V, v, C, pa, pb)) continue;
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                if (c < C) {
                    accN[c] = fmaf(2.f * pa[c], pb[c], accN[c]);
                    const float s = pa[c] + pb[c];
                    accD[c] = fmaf(s, s, accD[c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            if (c < C) {
                const float n = warp_sum(accN[c]), d = warp_sum(accD[c]);
                if (lane == 0) { red[warp][2 * c] = n; red[warp][2 * c + 1] = d; }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += THREADS) {
        double t = 0.0;
        for (int w = 0; w < THREADS / 32; ++w) t += (double)red[w][i];
        atomicAdd(&sums[(size_t)b * 2 * C + i], t);
    }
}

template <int CMAX>
__global__ void __launch_bounds__(THREADS, CMAX <= 16 ? 4 : 1) grad_kernel(const float *__restrict__ ta, const float *__restrict__ tb,
                                                       const float *__restrict__ gsums, float *__restrict__ grad_a, int C,
                                                       size_t V)
{
    __shared__ float g[2 * CMAX];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < 2 * C; i += THREADS) g[i] = gsums[(size_t)b * 2 * C + i];
    __syncthreads();
    const float *a = ta + (size_t)b * C * V, *bb = tb + (size_t)b * C * V;
    float *ga = grad_a + (size_t)b * C * V;
    for (size_t v = (size_t)blockIdx.x * THREADS + threadIdx.x; v < V; v += (size_t)gridDim.x * THREADS) {
        float pa[CMAX], pb[CMAX];
        if (!load_softmax<CMAX>(a, bb, V, v, C, pa, pb)) {
#pragma unroll
            for (int c = 0; c < CMAX; ++c)
                if (c < C) ga[(size_t)c * V + v] = 0.f;
            continue;
        }
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            if (c < C) {
                pb[c] = 2.f * (g[2 * c] * pb[c] + g[2 * c + 1] * (pa[c] + pb[c]));   // q_c (pb is dead afterwards)
                dot = fmaf(pa[c], pb[c], dot);
            }
        }
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) ga[(size_t)c * V + v] = pa[c] * (pb[c] - dot);
    }
}

static unsigned grid_x(size_t V)
{
    const size_t want = (V + THREADS - 1) / THREADS;
    const size_t cap = (size_t)sm_count() * 8;
    return (unsigned)(want < cap ? want : cap);
}

}  // namespace closs

void preload_consistency()
{
    DGTTA_TOUCH(closs::sums_kernel<8>); DGTTA_TOUCH(closs::sums_kernel<16>); DGTTA_TOUCH(closs::sums_kernel<32>);
    DGTTA_TOUCH(closs::grad_kernel<8>); DGTTA_TOUCH(closs::grad_kernel<16>); DGTTA_TOUCH(closs::grad_kernel<32>);
}

}  // namespace dgtta

using namespace dgtta;

static int closs_check(const void *a, const void *b, const void *o, int B, int C, long long V)
{
    if (!a || !b || !o) { set_error("dgtta_consistency: null pointer"); return DGTTA_ENULL; }
    if (B <= 0 || B > 65535 || C <= 0 || V <= 0) { set_error("dgtta_consistency: bad shape"); return DGTTA_EINVAL; }
    if (C > 128) { set_error("dgtta_consistency: more than 128 channels are not supported"); return DGTTA_EUNSUPPORTED; }
    return 0;
}

extern "C" int dgtta_consistency_sums_fwd(const float *target_a_dev, const float *target_b_dev, double *sums_dev, int B, int C,
                                          long long V, dgtta_stream_t stream_)
{
    int rc = closs_check(target_a_dev, target_b_dev, sums_dev, B, C, V);
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    cudaError_t e = cudaMemsetAsync(sums_dev, 0, (size_t)B * C * 2 * sizeof(double), stream);
    if (e != cudaSuccess) { set_error("dgtta_consistency_sums_fwd: memset: %s", cudaGetErrorString(e)); return (int)e; }
    const dim3 grid(closs::grid_x((size_t)V), (unsigned)B);
    if (C <= 8) closs::sums_kernel<8><<<grid, closs::THREADS, 0, stream>>>(target_a_dev, target_b_dev, sums_dev, C, (size_t)V);
    else if (C <= 16) closs::sums_kernel<16><<<grid, closs::THREADS, 0, stream>>>(target_a_dev, target_b_dev, sums_dev, C, (size_t)V);
    else if (C <= 32) closs::sums_kernel<32><<<grid, closs::THREADS, 0, stream>>>(target_a_dev, target_b_dev, sums_dev, C, (size_t)V);
    else closs::sums_kernel<128><<<grid, closs::THREADS, 0, stream>>>(target_a_dev, target_b_dev, sums_dev, C, (size_t)V);
    return check_launch("consistency_sums_kernel");
}

extern "C" int dgtta_consistency_sums_bwd(const float *target_a_dev, const float *target_b_dev, const float *grad_sums_dev,
                                          float *grad_a_dev, int B, int C, long long V, dgtta_stream_t stream_)
{
    int rc = closs_check(target_a_dev, target_b_dev, grad_a_dev, B, C, V);
    if (rc) return rc;
    if (!grad_sums_dev) { set_error("dgtta_consistency_sums_bwd: null pointer"); return DGTTA_ENULL; }
    cudaStream_t stream = (cudaStream_t)stream_;
    const dim3 grid(closs::grid_x((size_t)V), (unsigned)B);
    if (C <= 8) closs::grad_kernel<8><<<grid, closs::THREADS, 0, stream>>>(target_a_dev, target_b_dev, grad_sums_dev, grad_a_dev, C, (size_t)V);
    else if (C <= 16) closs::grad_kernel<16><<<grid, closs::THREADS, 0, stream>>>(target_a_dev, target_b_dev, grad_sums_dev, grad_a_dev, C, (size_t)V);
    else if (C <= 32) closs::grad_kernel<32><<<grid, closs::THREADS, 0, stream>>>(target_a_dev, target_b_dev, grad_sums_dev, grad_a_dev, C, (size_t)V);
    else closs::grad_kernel<128><<<grid, closs::THREADS, 0, stream>>>(target_a_dev, target_b_dev, grad_sums_dev, grad_a_dev, C, (size_t)V);
    return check_launch("consistency_grad_kernel");
}
