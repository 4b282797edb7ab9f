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
#include "sampler.cuh"

namespace dgtta {
namespace closs {

constexpr int THREADS = 256;

// exp(x) for x <= 0 as ex2.approx(x * log2 e): two instructions instead of expf's ~10; relative error <= 2^-22, far inside
// the loss tolerance (softmax probabilities; results below 2^-126 flush to 0)
__device__ __forceinline__ float exp_neg(float x) { return ex2_approx(x * 1.4426950408889634f); }

// EXACT: the channel count is the template parameter itself (no per-channel `c < C` predicates); otherwise CMAX is an
// upper bound and C the run-time count.
template <int CMAX, bool EXACT = false>
__device__ __forceinline__ bool load_softmax(const float *a, const float *b, size_t V, size_t v, int C_, float (&pa)[CMAX],
                                             float (&pb)[CMAX])
{
    const int C = EXACT ? CMAX : C_;
    float sa = 0.f, sb = 0.f, ma = -__int_as_float(0x7f800000), mb = ma;
    const float *qa = a + v, *qb = b + v;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
            pa[c] = __ldg(qa);
            pb[c] = __ldg(qb);
            qa += V; qb += V;
        }
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
            sa += pa[c]; sb += pb[c];
            ma = fmaxf(ma, pa[c]); mb = fmaxf(mb, pb[c]);
        }
    }
    if (!(sa > 0.f && sb > 0.f)) return false;   // outside the common content: contributes nothing, gradient zero
    float ea = 0.f, eb = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
            pa[c] = exp_neg(pa[c] - ma); pb[c] = exp_neg(pb[c] - mb);
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

// ---- fused inverse warp (the part round 1 left out): the two tensors above are
//     target_x = grid_sample(logits_x, affine_grid(R_x^-1), zeros)                                   (tta.py:571-575)
// and only exist to be consumed here.  The *_warp kernels take the un-warped logits and the two inverse affines, gather
// the C channels of both branches per output voxel (coordinates and corner weights once per voxel and branch) and feed the
// same mask / softmax / sums arithmetic — the warped logits are never materialised, forward or backward.
struct WarpGeo {
    SampleParams A, B;     // .in = logits of the branch, .theta = its inverse affine; sizes shared
};

// gathers the C channels of one branch at output voxel (w, h, d) of sample b; q (corner offsets / weights, out-of-volume
// corners with weight 0 at offset 0) is returned for the backward scatter
template <int CMAX>
__device__ __forceinline__ void gather_branch(const SampleParams &P, int b, int w, int h, int d, int C, float (&t)[CMAX], Corners &q)
{
    const Coords c = source_coords<DGTTA_PAD_ZEROS>(P, b, w, h, d);
    q = trilinear_corners(P, c);
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (q.off[k] < 0) { q.off[k] = 0; q.wgt[k] = 0.f; }
    const size_t Vi = (size_t)P.Di * P.Hi * P.Wi;
    const float *src = P.in + (size_t)b * C * Vi;
#pragma unroll
    for (int ch = 0; ch < CMAX; ++ch) {
        if (ch < C) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc = fmaf(__ldg(src + (size_t)ch * Vi + q.off[k]), q.wgt[k], acc);
            t[ch] = acc;
        }
    }
}

// mask + the two softmaxes on values already in registers; false: outside the common content
template <int CMAX>
__device__ __forceinline__ bool softmax_pair(int C, float (&pa)[CMAX], float (&pb)[CMAX])
{
    float sa = 0.f, sb = 0.f, ma = -__int_as_float(0x7f800000), mb = ma;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < C) { sa += pa[c]; sb += pb[c]; ma = fmaxf(ma, pa[c]); mb = fmaxf(mb, pb[c]); }
    }
    if (!(sa > 0.f && sb > 0.f)) return false;
    float ea = 0.f, eb = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < C) { pa[c] = exp_neg(pa[c] - ma); pb[c] = exp_neg(pb[c] - mb); ea += pa[c]; eb += pb[c]; }
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
template <int CMAX, bool EXACT = false>
__global__ void __launch_bounds__(THREADS, CMAX <= 16 ? 3 : 1) sums_kernel(const float *__restrict__ ta, const float *__restrict__ tb,
                                                                          double *__restrict__ sums, int C_, size_t V)
{
    const int C = EXACT ? CMAX : C_;
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
            const bool live = v < V && load_softmax<CMAX, EXACT>(a, bb, V, v, C, pa, pb);
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
            if (!load_softmax<CMAX, EXACT>(a, bb, V, v, C, pa, pb)) continue;
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

template <int CMAX, bool EXACT = false>
__global__ void __launch_bounds__(THREADS, CMAX <= 16 ? 4 : 1) grad_kernel(const float *__restrict__ ta, const float *__restrict__ tb,
                                                       const float *__restrict__ gsums, float *__restrict__ grad_a, int C_,
                                                       size_t V)
{
    const int C = EXACT ? CMAX : C_;
    __shared__ float g[2 * CMAX];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < 2 * C; i += THREADS) g[i] = gsums[(size_t)b * 2 * C + i];
    __syncthreads();
    const float *a = ta + (size_t)b * C * V, *bb = tb + (size_t)b * C * V;
    float *ga = grad_a + (size_t)b * C * V;
    for (size_t v = (size_t)blockIdx.x * THREADS + threadIdx.x; v < V; v += (size_t)gridDim.x * THREADS) {
        float pa[CMAX], pb[CMAX];
        if (!load_softmax<CMAX, EXACT>(a, bb, V, v, C, pa, pb)) {
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

// forward with the inverse warps fused in; C <= 16 (one accumulator per lane, as in sums_kernel).  Block = 32 x 8 output
// voxels of one d-plane, like the sampler.
template <int CMAX>
__global__ void __launch_bounds__(THREADS, 2) sums_warp_kernel(const __grid_constant__ WarpGeo G, double *__restrict__ sums, int C)
{
    __shared__ float red[THREADS / 32][2 * CMAX];
    const SampleParams &PA = G.A;
    const int ntw = (PA.Wo + 31) / 32;
    const int tw = blockIdx.x % ntw, th_ = blockIdx.x / ntw;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = tw * 32 + lane, h = th_ * 8 + warp, d = blockIdx.y, b = blockIdx.z;
    float pa[CMAX], pb[CMAX], x[32];
    bool live = w < PA.Wo && h < PA.Ho;
    if (live) {
        Corners q;
        gather_branch<CMAX>(G.A, b, w, h, d, C, pa, q);
        gather_branch<CMAX>(G.B, b, w, h, d, C, pb, q);
        live = softmax_pair<CMAX>(C, pa, pb);
    }
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
    const float acc = reduce_scatter32(x, lane);
    if (lane < 2 * CMAX) red[warp][lane] = acc;
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += THREADS) {
        double t = 0.0;
        for (int wv = 0; wv < THREADS / 32; ++wv) t += (double)red[wv][i];
        atomicAdd(&sums[(size_t)b * 2 * C + i], t);
    }
}

// backward with the inverse warp fused in: d loss / d logits_a (pre-zeroed) = adjoint of branch A's gather applied to the
// per-voxel gradient w.r.t. the warped logits, which is recomputed on the fly (both branches are gathered again).
// Neighbouring lanes merge the contributions to shared x-corners by shuffle before the global reduction, as in
// affine_sample_bwd_kernel.
template <int CMAX>
__global__ void __launch_bounds__(THREADS, 2) grad_warp_kernel(const __grid_constant__ WarpGeo G, const float *__restrict__ gsums,
                                                              float *__restrict__ grad_a, int C)
{
    __shared__ float g[2 * CMAX];
    const SampleParams &PA = G.A;
    const int ntw = (PA.Wo + 31) / 32;
    const int tw = blockIdx.x % ntw, th_ = blockIdx.x / ntw;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = tw * 32 + lane, h = th_ * 8 + warp, d = blockIdx.y, b = blockIdx.z;
    for (int i = threadIdx.x; i < 2 * C; i += THREADS) g[i] = gsums[(size_t)b * 2 * C + i];
    __syncthreads();
    const bool active = w < PA.Wo && h < PA.Ho;     // inactive lanes stay for the shuffles and contribute nothing
    float pa[CMAX], pb[CMAX];
    Corners qa, qb;
    bool live = active;
    gather_branch<CMAX>(G.A, b, min(w, PA.Wo - 1), min(h, PA.Ho - 1), d, C, pa, qa);
    gather_branch<CMAX>(G.B, b, min(w, PA.Wo - 1), min(h, PA.Ho - 1), d, C, pb, qb);
    live = softmax_pair<CMAX>(C, pa, pb) && active;
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
            pb[c] = 2.f * (g[2 * c] * pb[c] + g[2 * c + 1] * (pa[c] + pb[c]));   // q_c (pb is dead afterwards)
            dot = fmaf(pa[c], pb[c], dot);
        }
    }
    if (!active) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { qa.wgt[k] = 0.f; qa.off[k] = -1 - k; }     // distinct sentinels: never merged, never written
    }
    // pair p = (dy, dz): corners 2p (x0) and 2p+1 (x0+1); weight-0 corners (outside the volume) are skipped
    bool give[4], take[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int mine1 = qa.wgt[2 * p + 1] != 0.f ? qa.off[2 * p + 1] : -100;
        const int next_x0_raw = __shfl_down_sync(0xffffffffu, qa.off[2 * p], 1);
        const float next_w0 = __shfl_down_sync(0xffffffffu, qa.wgt[2 * p], 1);
        give[p] = lane < 31 && mine1 >= 0 && next_w0 != 0.f && next_x0_raw == mine1;
        take[p] = __shfl_up_sync(0xffffffffu, (int)give[p], 1) != 0 && lane > 0;
    }
    const size_t Vi = (size_t)PA.Di * PA.Hi * PA.Wi;
    float *gi = grad_a + (size_t)b * C * Vi;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
            const float gv = live ? pa[c] * (pb[c] - dot) : 0.f;       // d loss / d warped_a[c] at this voxel
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const float v1 = gv * qa.wgt[2 * p + 1];
                const float from_left = __shfl_up_sync(0xffffffffu, v1, 1);
                const float v0 = gv * qa.wgt[2 * p] + (take[p] ? from_left : 0.f);
                if (qa.wgt[2 * p] != 0.f && active) atomicAdd(gi + (size_t)c * Vi + qa.off[2 * p], v0);
                if (qa.wgt[2 * p + 1] != 0.f && active && !give[p]) atomicAdd(gi + (size_t)c * Vi + qa.off[2 * p + 1], v1);
            }
        }
    }
}

// channel counts up to 16 (every class subset the TTA plans use) get kernels specialised for the exact count
template <int C>
static void launch_sums_exact(int want, dim3 grid, cudaStream_t st, const float *a, const float *b, double *sums, size_t V)
{
    if (want == C) sums_kernel<C, true><<<grid, THREADS, 0, st>>>(a, b, sums, C, V);
    else if constexpr (C > 1) launch_sums_exact<C - 1>(want, grid, st, a, b, sums, V);
}
template <int C>
static void launch_grad_exact(int want, dim3 grid, cudaStream_t st, const float *a, const float *b, const float *g, float *ga, size_t V)
{
    if (want == C) grad_kernel<C, true><<<grid, THREADS, 0, st>>>(a, b, g, ga, C, V);
    else if constexpr (C > 1) launch_grad_exact<C - 1>(want, grid, st, a, b, g, ga, V);
}
template <int C>
static void touch_exact()
{
    DGTTA_TOUCH(sums_kernel<C, true>); DGTTA_TOUCH(grad_kernel<C, true>);
    if constexpr (C > 1) touch_exact<C - 1>();
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
    closs::touch_exact<16>();
    DGTTA_TOUCH(closs::sums_kernel<32>); DGTTA_TOUCH(closs::grad_kernel<32>);
    DGTTA_TOUCH(closs::sums_warp_kernel<8>); DGTTA_TOUCH(closs::sums_warp_kernel<16>);
    DGTTA_TOUCH(closs::grad_warp_kernel<8>); DGTTA_TOUCH(closs::grad_warp_kernel<16>);
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
    if (C <= 16) closs::launch_sums_exact<16>(C, grid, stream, target_a_dev, target_b_dev, sums_dev, (size_t)V);
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
    if (C <= 16) closs::launch_grad_exact<16>(C, grid, stream, target_a_dev, target_b_dev, grad_sums_dev, grad_a_dev, (size_t)V);
    else if (C <= 32) closs::grad_kernel<32><<<grid, closs::THREADS, 0, stream>>>(target_a_dev, target_b_dev, grad_sums_dev, grad_a_dev, C, (size_t)V);
    else closs::grad_kernel<128><<<grid, closs::THREADS, 0, stream>>>(target_a_dev, target_b_dev, grad_sums_dev, grad_a_dev, C, (size_t)V);
    return check_launch("consistency_grad_kernel");
}

static int warp_check(const void *a, const void *b, const void *ta, const void *tb, const void *o, int B, int C, int D, int H, int W)
{
    if (!a || !b || !ta || !tb || !o) { set_error("dgtta_consistency_warp: null pointer"); return DGTTA_ENULL; }
    if (B <= 0 || B > 65535 || C <= 0 || D <= 0 || D > 65535 || H <= 0 || W <= 0 || (size_t)D * H * W >= (size_t)1 << 31) {
        set_error("dgtta_consistency_warp: bad shape");
        return DGTTA_EINVAL;
    }
    if (C > 16) { set_error("dgtta_consistency_warp: more than 16 channels: warp first (dgtta_affine_sample_fwd), then dgtta_consistency_sums_*"); return DGTTA_EUNSUPPORTED; }
    return 0;
}

static closs::WarpGeo warp_geo(const float *a, const float *b, const float *ta, const float *tb, int B, int C, int D, int H, int W)
{
    closs::WarpGeo G;
    G.A = make_sample_params(a, ta, nullptr, nullptr, B, C, D, H, W, D, H, W);
    G.B = make_sample_params(b, tb, nullptr, nullptr, B, C, D, H, W, D, H, W);
    return G;
}

extern "C" int dgtta_consistency_warp_sums_fwd(const float *logits_a_dev, const float *logits_b_dev, const float *theta_a_dev,
                                               const float *theta_b_dev, double *sums_dev, int B, int C, int D, int H, int W,
                                               dgtta_stream_t stream_)
{
    int rc = warp_check(logits_a_dev, logits_b_dev, theta_a_dev, theta_b_dev, sums_dev, B, C, D, H, W);
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    cudaError_t e = cudaMemsetAsync(sums_dev, 0, (size_t)B * C * 2 * sizeof(double), stream);
    if (e != cudaSuccess) { set_error("dgtta_consistency_warp_sums_fwd: memset: %s", cudaGetErrorString(e)); return (int)e; }
    const closs::WarpGeo G = warp_geo(logits_a_dev, logits_b_dev, theta_a_dev, theta_b_dev, B, C, D, H, W);
    const dim3 grid((unsigned)(((W + 31) / 32) * ((H + 7) / 8)), (unsigned)D, (unsigned)B);
    if (C <= 8) closs::sums_warp_kernel<8><<<grid, closs::THREADS, 0, stream>>>(G, sums_dev, C);
    else closs::sums_warp_kernel<16><<<grid, closs::THREADS, 0, stream>>>(G, sums_dev, C);
    return check_launch("consistency_sums_warp_kernel");
}

extern "C" int dgtta_consistency_warp_sums_bwd(const float *logits_a_dev, const float *logits_b_dev, const float *theta_a_dev,
                                               const float *theta_b_dev, const float *grad_sums_dev, float *grad_logits_a_dev,
                                               int B, int C, int D, int H, int W, dgtta_stream_t stream_)
{
    int rc = warp_check(logits_a_dev, logits_b_dev, theta_a_dev, theta_b_dev, grad_logits_a_dev, B, C, D, H, W);
    if (rc) return rc;
    if (!grad_sums_dev) { set_error("dgtta_consistency_warp_sums_bwd: null pointer"); return DGTTA_ENULL; }
    cudaStream_t stream = (cudaStream_t)stream_;
    cudaError_t e = cudaMemsetAsync(grad_logits_a_dev, 0, (size_t)B * C * D * H * W * sizeof(float), stream);
    if (e != cudaSuccess) { set_error("dgtta_consistency_warp_sums_bwd: memset: %s", cudaGetErrorString(e)); return (int)e; }
    const closs::WarpGeo G = warp_geo(logits_a_dev, logits_b_dev, theta_a_dev, theta_b_dev, B, C, D, H, W);
    const dim3 grid((unsigned)(((W + 31) / 32) * ((H + 7) / 8)), (unsigned)D, (unsigned)B);
    if (C <= 8) closs::grad_warp_kernel<8><<<grid, closs::THREADS, 0, stream>>>(G, grad_sums_dev, grad_logits_a_dev, C);
    else closs::grad_warp_kernel<16><<<grid, closs::THREADS, 0, stream>>>(G, grad_sums_dev, grad_logits_a_dev, C);
    return check_launch("consistency_grad_warp_kernel");
}
