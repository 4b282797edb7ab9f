// Affine view warp: F.affine_grid + F.grid_sample fused into one gather (forward) / scatter (adjoint).
//
// Replaces the op pairs at dg_tta/tta/tta.py:523-532+548-551 (image warp, border padding),
// tta.py:571-575 (inverse warp of the prediction, zeros padding, autograd w.r.t. the input) and
// dg_tta/tta/torch_utils.py:55-73 (patch crop; nearest for labels).  align_corners=False throughout.
// Coordinates follow torch (third party; ATen/native/AffineGridGenerator.cpp, GridSampler.h):
//     base_i = linspace(-1, 1, N)[i] * (N-1) / N          (linspace evaluated from both ends)
//     (gx,gy,gz) = theta[b] . (base_w, base_h, base_d, 1)
//     ix = ((gx + 1) * W_in - 1) / 2 ; border: clamp to [0, W_in-1] ; zeros: corners outside contribute 0
// No grid tensor is ever materialised (the reference builds three 12 B/voxel grids per warp).
#include "common.cuh"

namespace dgtta {

struct SampleParams {
    const float *in;
    const float *theta;
    float *out;
    int B, C, Di, Hi, Wi, Do, Ho, Wo;
};

__device__ __forceinline__ float base_coord(int i, int n)
{
    if (n <= 1) return 0.f;
    const float step = __fdiv_rn(2.f, (float)(n - 1));
    const float v = (i < n / 2) ? __fadd_rn(-1.f, __fmul_rn(step, (float)i))
                                : __fsub_rn(1.f, __fmul_rn(step, (float)(n - 1 - i)));
    return __fdiv_rn(__fmul_rn(v, (float)(n - 1)), (float)n);
}

__device__ __forceinline__ float unnormalize(float g, int size)
{
    return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)size), 1.f), 0.5f);
}

__device__ __forceinline__ float clip_coord(float v, int size) { return fminf(fmaxf(v, 0.f), (float)(size - 1)); }

struct Coords {
    float ix, iy, iz;
};

// Block = 32 (w) x 8 (h) output voxels of one d-plane; the per-axis base coordinates (two IEEE divisions each)
// are tabulated once per block in shared memory instead of being recomputed per voxel.
constexpr int SBX = 32, SBY = 8;
constexpr int SAMPLE_THREADS = SBX * SBY;

struct BlockCoords {
    float th[12];
    float bx[SBX];
    float by[SBY];
    float bz;
};

__device__ __forceinline__ void block_setup(const SampleParams &P, BlockCoords &S, int b, int w0, int h0, int d)
{
    const int t = threadIdx.y * SBX + threadIdx.x;
    if (t < 12) S.th[t] = P.theta[b * 12 + t];
    if (t >= 32 && t < 32 + SBX) S.bx[t - 32] = base_coord(min(w0 + t - 32, P.Wo - 1), P.Wo);
    if (t >= 64 && t < 64 + SBY) S.by[t - 64] = base_coord(min(h0 + t - 64, P.Ho - 1), P.Ho);
    if (t == 96) S.bz = base_coord(d, P.Do);
    __syncthreads();
}

template <int PAD>
__device__ __forceinline__ Coords source_coords(const SampleParams &P, const BlockCoords &S)
{
    const float *th = S.th;
    const float xn = S.bx[threadIdx.x], yn = S.by[threadIdx.y], zn = S.bz;
    // row-times-column products summed left to right, like the reference's base_grid @ theta^T
    const float gx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(th[0], xn), __fmul_rn(th[1], yn)), __fmul_rn(th[2], zn)), th[3]);
    const float gy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(th[4], xn), __fmul_rn(th[5], yn)), __fmul_rn(th[6], zn)), th[7]);
    const float gz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(th[8], xn), __fmul_rn(th[9], yn)), __fmul_rn(th[10], zn)), th[11]);
    Coords c;
    c.ix = unnormalize(gx, P.Wi); c.iy = unnormalize(gy, P.Hi); c.iz = unnormalize(gz, P.Di);
    if (PAD == DGTTA_PAD_BORDER) { c.ix = clip_coord(c.ix, P.Wi); c.iy = clip_coord(c.iy, P.Hi); c.iz = clip_coord(c.iz, P.Di); }
    return c;
}

struct Corners {
    float wgt[8];
    int off[8];   // element offset inside one channel volume, -1 = outside (zeros padding)
};

__device__ __forceinline__ Corners trilinear_corners(const SampleParams &P, const Coords &c)
{
    const float fx = floorf(c.ix), fy = floorf(c.iy), fz = floorf(c.iz);
    const float tx = c.ix - fx, ty = c.iy - fy, tz = c.iz - fz;
    // clamp before the int conversion so that wild coordinates cannot overflow
    const int x0 = (int)fminf(fmaxf(fx, -2.f), (float)P.Wi), y0 = (int)fminf(fmaxf(fy, -2.f), (float)P.Hi),
              z0 = (int)fminf(fmaxf(fz, -2.f), (float)P.Di);
    Corners q;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
        const int xx = x0 + dx, yy = y0 + dy, zz = z0 + dz;
        const bool ok = xx >= 0 && xx < P.Wi && yy >= 0 && yy < P.Hi && zz >= 0 && zz < P.Di;
        q.wgt[k] = (dx ? tx : 1.f - tx) * (dy ? ty : 1.f - ty) * (dz ? tz : 1.f - tz);
        q.off[k] = ok ? (zz * P.Hi + yy) * P.Wi + xx : -1;
    }
    return q;
}

// grid: x = w-tiles * h-tiles, y = d, z = b.  Channels are looped inside (4 at a time: 32 independent gathers in
// flight) so that coordinates and weights are computed once per voxel.
template <int INTERP, int PAD>
__global__ void __launch_bounds__(SAMPLE_THREADS) affine_sample_fwd_kernel(const __grid_constant__ SampleParams P)
{
    __shared__ BlockCoords S;
    const int ntw = (P.Wo + SBX - 1) / SBX;
    const int tw = blockIdx.x % ntw, th_ = blockIdx.x / ntw;
    const int w0 = tw * SBX, h0 = th_ * SBY, d = blockIdx.y, b = blockIdx.z;
    block_setup(P, S, b, w0, h0, d);
    const int w = w0 + threadIdx.x, h = h0 + threadIdx.y;
    if (w >= P.Wo || h >= P.Ho) return;
    const size_t Vo = (size_t)P.Do * P.Ho * P.Wo, Vi = (size_t)P.Di * P.Hi * P.Wi;
    const Coords c = source_coords<PAD>(P, S);
    const float *src = P.in + (size_t)b * P.C * Vi;
    float *dst = P.out + (size_t)b * P.C * Vo + ((size_t)d * P.Ho + h) * P.Wo + w;
    if (INTERP == DGTTA_INTERP_NEAREST) {
        const float rx = nearbyintf(c.ix), ry = nearbyintf(c.iy), rz = nearbyintf(c.iz);
        const bool ok = rx >= 0.f && rx < (float)P.Wi && ry >= 0.f && ry < (float)P.Hi && rz >= 0.f && rz < (float)P.Di;
        const size_t off = ok ? ((size_t)(int)rz * P.Hi + (int)ry) * P.Wi + (int)rx : 0;
        for (int ch = 0; ch < P.C; ++ch) __stcs(dst + ch * Vo, ok ? __ldg(src + ch * Vi + off) : 0.f);
        return;
    }
    Corners q = trilinear_corners(P, c);
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (q.off[k] < 0) { q.off[k] = 0; q.wgt[k] = 0.f; }
    // channel loop: the eight corner offsets and weights stay in registers, only the two channel base pointers move
    // (one 64-bit add each per trip); four channels per trip keep 32 independent gathers in flight
    const float *cb = src;
    float *db = dst;
    int ch = 0;
    for (; ch + 4 <= P.C; ch += 4) {
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) v[u][k] = __ldg(cb + (size_t)u * Vi + q.off[k]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc = fmaf(v[u][k], q.wgt[k], acc);
            __stcs(db + (size_t)u * Vo, acc);
        }
        cb += 4 * Vi;
        db += 4 * Vo;
    }
    for (; ch + 1 < P.C; ++ch) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(__ldg(cb + q.off[k]), q.wgt[k], acc);
        __stcs(db, acc);
        cb += Vi;
        db += Vo;
    }
    if (ch < P.C) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(__ldg(cb + q.off[k]), q.wgt[k], acc);
        __stcs(db, acc);
    }
}

// adjoint of the trilinear forward: scatter grad_out into grad_in (pre-zeroed) with red.global.add.
// A warp is 32 consecutive output voxels along w; under the near-identity affines of the TTA loop lane i's x0+1 corner
// is lane i+1's x0 corner, so that half of each lane's contributions is handed to the neighbour by shuffle and added
// there first: ~4.1 instead of 8 global reductions per voxel and channel (the scatter is bound by L2 atomics).
template <int PAD>
__global__ void __launch_bounds__(SAMPLE_THREADS) affine_sample_bwd_kernel(const __grid_constant__ SampleParams P)
{
    // here P.in = grad_out [B,C,Do,Ho,Wo], P.out = grad_in [B,C,Di,Hi,Wi]
    __shared__ BlockCoords S;
    const int ntw = (P.Wo + SBX - 1) / SBX;
    const int tw = blockIdx.x % ntw, th_ = blockIdx.x / ntw;
    const int w0 = tw * SBX, h0 = th_ * SBY, d = blockIdx.y, b = blockIdx.z;
    block_setup(P, S, b, w0, h0, d);
    const int lane = threadIdx.x;   // blockDim.x == 32: a warp is one row of the block
    const int w = w0 + lane, h = h0 + threadIdx.y;
    const bool active = w < P.Wo && h < P.Ho;   // inactive lanes stay for the shuffles and contribute nothing
    const size_t Vo = (size_t)P.Do * P.Ho * P.Wo, Vi = (size_t)P.Di * P.Hi * P.Wi;
    const Coords c = source_coords<PAD>(P, S);
    Corners q = trilinear_corners(P, c);
    if (!active) {
#pragma unroll
        for (int k = 0; k < 8; ++k) q.off[k] = -1;
    }
    // pair p = (dy, dz): corners 2p (x0) and 2p+1 (x0+1)
    bool give[4], take[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int next_x0 = __shfl_down_sync(0xffffffffu, q.off[2 * p], 1);
        give[p] = lane < 31 && q.off[2 * p + 1] >= 0 && next_x0 == q.off[2 * p + 1];
        take[p] = __shfl_up_sync(0xffffffffu, (int)give[p], 1) != 0 && lane > 0;
    }
    const float *go = P.in + (size_t)b * P.C * Vo + ((size_t)d * P.Ho + min(h, P.Ho - 1)) * P.Wo + min(w, P.Wo - 1);
    float *gi = P.out + (size_t)b * P.C * Vi;
    auto scatter = [&](float g, float *base) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const float v1 = g * q.wgt[2 * p + 1];
            const float from_left = __shfl_up_sync(0xffffffffu, v1, 1);
            const float v0 = g * q.wgt[2 * p] + (take[p] ? from_left : 0.f);
            if (q.off[2 * p] >= 0) atomicAdd(base + q.off[2 * p], v0);
            if (q.off[2 * p + 1] >= 0 && !give[p]) atomicAdd(base + q.off[2 * p + 1], v1);
        }
    };
    int ch = 0;
    for (; ch + 4 <= P.C; ch += 4) {
        float g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) g[u] = active ? __ldg(go + (size_t)u * Vo) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) scatter(g[u], gi + (size_t)u * Vi);
        go += 4 * Vo;
        gi += 4 * Vi;
    }
    for (; ch < P.C; ++ch) {
        scatter(active ? __ldg(go) : 0.f, gi);
        go += Vo;
        gi += Vi;
    }
}

// Label crop of get_batch (dg_tta/tta/torch_utils.py:71-82): nearest-neighbour sampling of the one-hot label channels
// (zeros outside), background channel prepended where no label is set, argmax over channels — in one pass, writing the
// int64 label map directly instead of an L-channel float patch plus sum / cat / argmax passes.  Ties go to the lowest
// index like torch.argmax; the background channel is index 0.
__global__ void __launch_bounds__(SAMPLE_THREADS) affine_label_argmax_kernel(const __grid_constant__ SampleParams P, long long *out)
{
    __shared__ BlockCoords S;
    const int ntw = (P.Wo + SBX - 1) / SBX;
    const int tw = blockIdx.x % ntw, th_ = blockIdx.x / ntw;
    const int w0 = tw * SBX, h0 = th_ * SBY, d = blockIdx.y, b = blockIdx.z;
    block_setup(P, S, b, w0, h0, d);
    const int w = w0 + threadIdx.x, h = h0 + threadIdx.y;
    if (w >= P.Wo || h >= P.Ho) return;
    const size_t Vo = (size_t)P.Do * P.Ho * P.Wo, Vi = (size_t)P.Di * P.Hi * P.Wi;
    const Coords c = source_coords<DGTTA_PAD_ZEROS>(P, S);
    const float rx = nearbyintf(c.ix), ry = nearbyintf(c.iy), rz = nearbyintf(c.iz);
    const bool ok = rx >= 0.f && rx < (float)P.Wi && ry >= 0.f && ry < (float)P.Hi && rz >= 0.f && rz < (float)P.Di;
    long long label = 0;
    if (ok) {
        const float *src = P.in + (size_t)b * P.C * Vi + ((size_t)(int)rz * P.Hi + (int)ry) * P.Wi + (int)rx;
        float sum = 0.f, best = -__int_as_float(0x7f800000);
        int arg = 0;
        for (int ch = 0; ch < P.C; ++ch) {
            const float v = __ldg(src + (size_t)ch * Vi);
            sum += v;
            if (v > best) { best = v; arg = ch + 1; }
        }
        const float bg = sum < 1.0f ? 1.f : 0.f;     // get_argmaxed_segs: (segs.sum(1) < 1.0).float()
        label = bg >= best ? 0 : arg;
    }
    // outside the volume every channel samples 0: sum = 0 < 1 -> background
    out[(size_t)b * Vo + ((size_t)d * P.Ho + h) * P.Wo + w] = label;
}

void preload_sampler()
{
    DGTTA_TOUCH(affine_label_argmax_kernel);
    DGTTA_TOUCH(affine_sample_fwd_kernel<DGTTA_INTERP_TRILINEAR, DGTTA_PAD_ZEROS>);
    DGTTA_TOUCH(affine_sample_fwd_kernel<DGTTA_INTERP_TRILINEAR, DGTTA_PAD_BORDER>);
    DGTTA_TOUCH(affine_sample_fwd_kernel<DGTTA_INTERP_NEAREST, DGTTA_PAD_ZEROS>);
    DGTTA_TOUCH(affine_sample_fwd_kernel<DGTTA_INTERP_NEAREST, DGTTA_PAD_BORDER>);
    DGTTA_TOUCH(affine_sample_bwd_kernel<DGTTA_PAD_ZEROS>);
    DGTTA_TOUCH(affine_sample_bwd_kernel<DGTTA_PAD_BORDER>);
}

static int sample_check(const void *a, const void *t, const void *o, int B, int C, int Di, int Hi, int Wi, int Do, int Ho, int Wo)
{
    if (!a || !t || !o) { set_error("dgtta_affine_sample: null pointer"); return DGTTA_ENULL; }
    if (B <= 0 || C <= 0 || Di <= 0 || Hi <= 0 || Wi <= 0 || Do <= 0 || Ho <= 0 || Wo <= 0 || B > 65535 || Do > 65535) {
        set_error("dgtta_affine_sample: bad shape");
        return DGTTA_EINVAL;
    }
    if ((size_t)Di * Hi * Wi >= (size_t)1 << 31 || (size_t)Do * Ho * Wo >= (size_t)1 << 31) {
        set_error("dgtta_affine_sample: volume too large");
        return DGTTA_EINVAL;
    }
    return 0;
}

static dim3 sample_grid(int B, int Do, int Ho, int Wo)
{
    return dim3((unsigned)(((Wo + SBX - 1) / SBX) * ((Ho + SBY - 1) / SBY)), (unsigned)Do, (unsigned)B);
}

}  // namespace dgtta

using namespace dgtta;

extern "C" int dgtta_affine_sample_fwd(const float *in_dev, const float *theta_dev, float *out_dev, int B, int C,
                                       int Di, int Hi, int Wi, int Do, int Ho, int Wo, int interp, int padding,
                                       dgtta_stream_t stream_)
{
    int rc = sample_check(in_dev, theta_dev, out_dev, B, C, Di, Hi, Wi, Do, Ho, Wo);
    if (rc) return rc;
    if ((interp != DGTTA_INTERP_TRILINEAR && interp != DGTTA_INTERP_NEAREST) ||
        (padding != DGTTA_PAD_ZEROS && padding != DGTTA_PAD_BORDER)) {
        set_error("dgtta_affine_sample_fwd: bad interp/padding");
        return DGTTA_EINVAL;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    SampleParams P{in_dev, theta_dev, out_dev, B, C, Di, Hi, Wi, Do, Ho, Wo};
    const dim3 grid = sample_grid(B, Do, Ho, Wo);
    const dim3 block(SBX, SBY, 1);
    if (interp == DGTTA_INTERP_TRILINEAR) {
        if (padding == DGTTA_PAD_ZEROS) affine_sample_fwd_kernel<DGTTA_INTERP_TRILINEAR, DGTTA_PAD_ZEROS><<<grid, block, 0, stream>>>(P);
        else affine_sample_fwd_kernel<DGTTA_INTERP_TRILINEAR, DGTTA_PAD_BORDER><<<grid, block, 0, stream>>>(P);
    } else {
        if (padding == DGTTA_PAD_ZEROS) affine_sample_fwd_kernel<DGTTA_INTERP_NEAREST, DGTTA_PAD_ZEROS><<<grid, block, 0, stream>>>(P);
        else affine_sample_fwd_kernel<DGTTA_INTERP_NEAREST, DGTTA_PAD_BORDER><<<grid, block, 0, stream>>>(P);
    }
    return check_launch("affine_sample_fwd_kernel");
}

extern "C" int dgtta_affine_sample_bwd_input(const float *grad_out_dev, const float *theta_dev, float *grad_in_dev,
                                             int B, int C, int Di, int Hi, int Wi, int Do, int Ho, int Wo,
                                             int padding, dgtta_stream_t stream_)
{
    int rc = sample_check(grad_out_dev, theta_dev, grad_in_dev, B, C, Di, Hi, Wi, Do, Ho, Wo);
    if (rc) return rc;
    if (padding != DGTTA_PAD_ZEROS && padding != DGTTA_PAD_BORDER) { set_error("dgtta_affine_sample_bwd_input: bad padding"); return DGTTA_EINVAL; }
    cudaStream_t stream = (cudaStream_t)stream_;
    cudaError_t e = cudaMemsetAsync(grad_in_dev, 0, (size_t)B * C * Di * Hi * Wi * sizeof(float), stream);
    if (e != cudaSuccess) { set_error("dgtta_affine_sample_bwd_input: memset: %s", cudaGetErrorString(e)); return (int)e; }
    SampleParams P{grad_out_dev, theta_dev, grad_in_dev, B, C, Di, Hi, Wi, Do, Ho, Wo};
    const dim3 grid = sample_grid(B, Do, Ho, Wo);
    const dim3 block(SBX, SBY, 1);
    if (padding == DGTTA_PAD_ZEROS) affine_sample_bwd_kernel<DGTTA_PAD_ZEROS><<<grid, block, 0, stream>>>(P);
    else affine_sample_bwd_kernel<DGTTA_PAD_BORDER><<<grid, block, 0, stream>>>(P);
    return check_launch("affine_sample_bwd_kernel");
}

extern "C" int dgtta_affine_label_argmax(const float *onehot_dev, const float *theta_dev, long long *out_dev, int B, int L,
                                         int Di, int Hi, int Wi, int Do, int Ho, int Wo, dgtta_stream_t stream_)
{
    int rc = sample_check(onehot_dev, theta_dev, out_dev, B, L, Di, Hi, Wi, Do, Ho, Wo);
    if (rc) return rc;
    SampleParams P{onehot_dev, theta_dev, nullptr, B, L, Di, Hi, Wi, Do, Ho, Wo};
    affine_label_argmax_kernel<<<sample_grid(B, Do, Ho, Wo), dim3(SBX, SBY, 1), 0, (cudaStream_t)stream_>>>(P, out_dev);
    return check_launch("affine_label_argmax_kernel");
}
