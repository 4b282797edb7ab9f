// Affine view warp: F.affine_grid + F.grid_sample fused into one gather (forward) / scatter (adjoint).
//
// Replaces the op pairs at dg_tta/tta/tta.py:523-532+548-551 (image warp, border padding),
// tta.py:571-575 (inverse warp of the prediction, zeros padding, autograd w.r.t. the input) and
// dg_tta/tta/torch_utils.py:55-73 (patch crop; nearest for labels).  align_corners=False throughout.
// Coordinates follow torch (third party; ATen/native/AffineGridGenerator.cpp, GridSampler.h):
//     base_i = linspace(-1, 1, N)[i] * (N-1) / N          linspace: fma(step, i, -1) below N/2, fma(-step, N-1-i, 1) above
//     (gx,gy,gz) = base @ theta[b]^T                      K = 4 accumulated in order with FMAs: fma(z,t2, fma(y,t1, x*t0)) + t3
//   — the rounding sequence torch's kernels execute (RangeFactories linspace, tensor mul / true div, bmm), reproduced
//   operation by operation so that mode="nearest" picks the same source voxel on .5 ties: bit-exact index work
//   (tests/test_sampler_gpu.py compares the label crops with np.array_equal against the reference's outputs).
//     ix = ((gx + 1) * W_in - 1) / 2 ; border: clamp to [0, W_in-1] ; zeros: corners outside contribute 0
// No grid tensor is ever materialised (the reference builds three 12 B/voxel grids per warp).
#include "sampler.cuh"

namespace dgtta {

// grid: x = w-tiles * h-tiles, y = d, z = b.  Channels are looped inside (4 at a time: 32 independent gathers in
// flight) so that coordinates and weights are computed once per voxel.
template <int INTERP, int PAD>
__global__ void __launch_bounds__(SAMPLE_THREADS) affine_sample_fwd_kernel(const __grid_constant__ SampleParams P)
{
    const int ntw = (P.Wo + SBX - 1) / SBX;
    const int tw = blockIdx.x % ntw, th_ = blockIdx.x / ntw;
    const int w0 = tw * SBX, h0 = th_ * SBY, d = blockIdx.y, b = blockIdx.z;
    const int w = w0 + threadIdx.x, h = h0 + threadIdx.y;
    if (w >= P.Wo || h >= P.Ho) return;
    const size_t Vo = (size_t)P.Do * P.Ho * P.Wo, Vi = (size_t)P.Di * P.Hi * P.Wi;
    const Coords c = source_coords<PAD>(P, b, w, h, d);
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
    if (P.bias) {
        // get_batch's image crop (torch_utils.py:58-62): grid_sample(vol - min, zeros) + min in one gather — the shift is
        // applied to every in-bounds corner before the weighting, the out-of-bounds corners contribute 0 (their weight is 0)
        const float bias = __ldg(P.bias + b);
        for (int ch = 0; ch < P.C; ++ch) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc = fmaf(__fsub_rn(__ldg(src + (size_t)ch * Vi + q.off[k]), bias), q.wgt[k], acc);
            __stcs(dst + (size_t)ch * Vo, __fadd_rn(acc, bias));
        }
        return;
    }
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
    const int ntw = (P.Wo + SBX - 1) / SBX;
    const int tw = blockIdx.x % ntw, th_ = blockIdx.x / ntw;
    const int w0 = tw * SBX, h0 = th_ * SBY, d = blockIdx.y, b = blockIdx.z;
    const int lane = threadIdx.x;   // blockDim.x == 32: a warp is one row of the block
    const int w = w0 + lane, h = h0 + threadIdx.y;
    const bool active = w < P.Wo && h < P.Ho;   // inactive lanes stay for the shuffles and contribute nothing
    const size_t Vo = (size_t)P.Do * P.Ho * P.Wo, Vi = (size_t)P.Di * P.Hi * P.Wi;
    const Coords c = source_coords<PAD>(P, b, w, h, d);
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
    const int ntw = (P.Wo + SBX - 1) / SBX;
    const int tw = blockIdx.x % ntw, th_ = blockIdx.x / ntw;
    const int w0 = tw * SBX, h0 = th_ * SBY, d = blockIdx.y, b = blockIdx.z;
    const int w = w0 + threadIdx.x, h = h0 + threadIdx.y;
    if (w >= P.Wo || h >= P.Ho) return;
    const size_t Vo = (size_t)P.Do * P.Ho * P.Wo, Vi = (size_t)P.Di * P.Hi * P.Wi;
    const Coords c = source_coords<DGTTA_PAD_ZEROS>(P, b, w, h, d);
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

// Integer label path.  Nearest sampling picks ONE source voxel, and get_argmaxed_segs is a per-voxel
// function of that voxel's L channel values, so the two commute: argmaxed(nearest_sample(onehot)) ==
// nearest_sample(argmaxed(onehot)) with 0 (background) outside the volume.  label_map_kernel evaluates the per-voxel
// rule once per volume (the host side caches it on the resident tensor); label_gather_kernel then crops from a
// 2 B/voxel map instead of the L x 4 B/voxel one-hot volume (104 channels in the reference's TotalSegmentator tasks,
// dg_tta/tta/nnunet_utils.py:191-195).
__global__ void __launch_bounds__(256) label_map_kernel(const float *onehot, short *map, int L, long long V)
{
    const long long v = (long long)blockIdx.x * 256 + threadIdx.x;
    const int b = blockIdx.y;
    if (v >= V) return;
    const float *src = onehot + (size_t)b * L * V + v;
    float sum = 0.f, best = -__int_as_float(0x7f800000);
    int arg = 0;
    for (int ch = 0; ch < L; ++ch) {
        const float x = __ldg(src + (size_t)ch * V);
        sum += x;                                   // torch's sum over dim 1 runs in channel order too
        if (x > best) { best = x; arg = ch + 1; }
    }
    const float bg = sum < 1.0f ? 1.f : 0.f;
    map[(size_t)b * V + v] = (short)(bg >= best ? 0 : arg);
}

__global__ void __launch_bounds__(SAMPLE_THREADS) label_gather_kernel(const __grid_constant__ SampleParams P, const short *map,
                                                                      long long *out)
{
    const int ntw = (P.Wo + SBX - 1) / SBX;
    const int tw = blockIdx.x % ntw, th_ = blockIdx.x / ntw;
    const int w = tw * SBX + threadIdx.x, h = th_ * SBY + threadIdx.y, d = blockIdx.y, b = blockIdx.z;
    if (w >= P.Wo || h >= P.Ho) return;
    const size_t Vo = (size_t)P.Do * P.Ho * P.Wo, Vi = (size_t)P.Di * P.Hi * P.Wi;
    const Coords c = source_coords<DGTTA_PAD_ZEROS>(P, b, w, h, d);
    const float rx = nearbyintf(c.ix), ry = nearbyintf(c.iy), rz = nearbyintf(c.iz);
    const bool ok = rx >= 0.f && rx < (float)P.Wi && ry >= 0.f && ry < (float)P.Hi && rz >= 0.f && rz < (float)P.Di;
    long long label = 0;
    if (ok) label = map[(size_t)b * Vi + ((size_t)(int)rz * P.Hi + (int)ry) * P.Wi + (int)rx];
    out[(size_t)b * Vo + ((size_t)d * P.Ho + h) * P.Wo + w] = label;
}

// min over a volume (img.min() of torch_utils.py:58): warp-shuffle + block reduction to per-block partials, then one
// block over the partials.  min is exact in any order, so the result is bitwise torch's.
constexpr int MIN_BLOCKS = 592;   // 148 SMs x 4
__global__ void __launch_bounds__(256) volume_min_partial_kernel(const float *in, long long n, float *partial)
{
    __shared__ float red[8];
    float m = __int_as_float(0x7f800000);
    bool nan = false;
    const long long n4 = ((reinterpret_cast<uintptr_t>(in) & 15) == 0) ? n / 4 : 0;
    const float4 *in4 = reinterpret_cast<const float4 *>(in);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const float4 v = __ldg(in4 + i);
        m = fminf(fminf(m, v.x), fminf(fminf(v.y, v.z), v.w));
        nan |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float v = __ldg(in + i);
        m = fminf(m, v);
        nan |= v != v;
    }
    if (nan) m = __int_as_float(0x7fc00000);      // torch.min propagates NaN; fminf would drop it
    // NaN-propagating reduction: a NaN partial wins
    auto comb = [](float a, float b) { return (a != a) ? a : ((b != b) ? b : fminf(a, b)); };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = comb(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < 8 ? red[threadIdx.x] : __int_as_float(0x7f800000);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) m = comb(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) partial[blockIdx.x] = m;
    }
}
__global__ void __launch_bounds__(1024) volume_min_final_kernel(const float *partial, int n, float *out)
{
    __shared__ float red[32];
    auto comb = [](float a, float b) { return (a != a) ? a : ((b != b) ? b : fminf(a, b)); };
    float m = threadIdx.x < n ? partial[threadIdx.x] : __int_as_float(0x7f800000);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = comb(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = red[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = comb(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) out[0] = m;
    }
}

void preload_sampler()
{
    DGTTA_TOUCH(label_map_kernel);
    DGTTA_TOUCH(label_gather_kernel);
    DGTTA_TOUCH(volume_min_partial_kernel);
    DGTTA_TOUCH(volume_min_final_kernel);
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

static int sample_fwd(const float *in_dev, const float *theta_dev, const float *bias_dev, float *out_dev, int B, int C, int Di,
                      int Hi, int Wi, int Do, int Ho, int Wo, int interp, int padding, dgtta_stream_t stream_);

extern "C" int dgtta_affine_sample_fwd(const float *in_dev, const float *theta_dev, float *out_dev, int B, int C,
                                       int Di, int Hi, int Wi, int Do, int Ho, int Wo, int interp, int padding,
                                       dgtta_stream_t stream_)
{
    return sample_fwd(in_dev, theta_dev, nullptr, out_dev, B, C, Di, Hi, Wi, Do, Ho, Wo, interp, padding, stream_);
}

extern "C" int dgtta_affine_crop_shifted_fwd(const float *in_dev, const float *theta_dev, const float *shift_dev, float *out_dev,
                                             int B, int C, int Di, int Hi, int Wi, int Do, int Ho, int Wo,
                                             dgtta_stream_t stream_)
{
    if (!shift_dev) { set_error("dgtta_affine_crop_shifted_fwd: null shift"); return DGTTA_ENULL; }
    return sample_fwd(in_dev, theta_dev, shift_dev, out_dev, B, C, Di, Hi, Wi, Do, Ho, Wo, DGTTA_INTERP_TRILINEAR,
                      DGTTA_PAD_ZEROS, stream_);
}

static int sample_fwd(const float *in_dev, const float *theta_dev, const float *bias_dev, float *out_dev, int B, int C, int Di,
                      int Hi, int Wi, int Do, int Ho, int Wo, int interp, int padding, dgtta_stream_t stream_)
{
    int rc = sample_check(in_dev, theta_dev, out_dev, B, C, Di, Hi, Wi, Do, Ho, Wo);
    if (rc) return rc;
    if ((interp != DGTTA_INTERP_TRILINEAR && interp != DGTTA_INTERP_NEAREST) ||
        (padding != DGTTA_PAD_ZEROS && padding != DGTTA_PAD_BORDER)) {
        set_error("dgtta_affine_sample_fwd: bad interp/padding");
        return DGTTA_EINVAL;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    const SampleParams P = make_sample_params(in_dev, theta_dev, out_dev, bias_dev, B, C, Di, Hi, Wi, Do, Ho, Wo);
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
    const SampleParams P = make_sample_params(grad_out_dev, theta_dev, grad_in_dev, nullptr, B, C, Di, Hi, Wi, Do, Ho, Wo);
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
    const SampleParams P = make_sample_params(onehot_dev, theta_dev, nullptr, nullptr, B, L, Di, Hi, Wi, Do, Ho, Wo);
    affine_label_argmax_kernel<<<sample_grid(B, Do, Ho, Wo), dim3(SBX, SBY, 1), 0, (cudaStream_t)stream_>>>(P, out_dev);
    return check_launch("affine_label_argmax_kernel");
}

extern "C" int dgtta_label_map_from_onehot(const float *onehot_dev, short *map_dev, int B, int L, long long V,
                                           dgtta_stream_t stream_)
{
    if (!onehot_dev || !map_dev) { set_error("dgtta_label_map_from_onehot: null pointer"); return DGTTA_ENULL; }
    if (B <= 0 || B > 65535 || L <= 0 || L > 32766 || V <= 0 || V >= (1ll << 38)) { set_error("dgtta_label_map_from_onehot: bad shape"); return DGTTA_EINVAL; }
    label_map_kernel<<<dim3((unsigned)((V + 255) / 256), (unsigned)B), 256, 0, (cudaStream_t)stream_>>>(onehot_dev, map_dev, L, V);
    return check_launch("label_map_kernel");
}

extern "C" int dgtta_affine_label_gather(const short *map_dev, const float *theta_dev, long long *out_dev, int B, int Di, int Hi,
                                         int Wi, int Do, int Ho, int Wo, dgtta_stream_t stream_)
{
    int rc = sample_check(map_dev, theta_dev, out_dev, B, 1, Di, Hi, Wi, Do, Ho, Wo);
    if (rc) return rc;
    const SampleParams P = make_sample_params(nullptr, theta_dev, nullptr, nullptr, B, 1, Di, Hi, Wi, Do, Ho, Wo);
    label_gather_kernel<<<sample_grid(B, Do, Ho, Wo), dim3(SBX, SBY, 1), 0, (cudaStream_t)stream_>>>(P, map_dev, out_dev);
    return check_launch("label_gather_kernel");
}

extern "C" size_t dgtta_volume_min_workspace_bytes(void) { return MIN_BLOCKS * sizeof(float); }

extern "C" int dgtta_volume_min(const float *in_dev, long long numel, float *out_dev, void *workspace_dev, size_t workspace_bytes,
                                dgtta_stream_t stream_)
{
    if (!in_dev || !out_dev || !workspace_dev) { set_error("dgtta_volume_min: null pointer"); return DGTTA_ENULL; }
    if (numel <= 0) { set_error("dgtta_volume_min: empty input"); return DGTTA_EINVAL; }
    if (workspace_bytes < MIN_BLOCKS * sizeof(float)) { set_error("dgtta_volume_min: workspace too small"); return DGTTA_EWORKSPACE; }
    const long long want = (numel + 256 * 16 - 1) / (256 * 16);
    const int blocks = (int)(want < 1 ? 1 : (want > MIN_BLOCKS ? MIN_BLOCKS : want));
    volume_min_partial_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(in_dev, numel, (float *)workspace_dev);
    int rc = check_launch("volume_min_partial_kernel");
    if (rc) return rc;
    volume_min_final_kernel<<<1, 1024, 0, (cudaStream_t)stream_>>>((const float *)workspace_dev, blocks, out_dev);
    return check_launch("volume_min_final_kernel");
}
